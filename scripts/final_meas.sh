# Round-end measurement set on ONE GPU (run under gpurun): tests, both bench arms, launch list of the same command.
# usage: bash scripts/final_meas.sh [tag]   (outputs: gpurun_out/*_<tag>.*)
T=${1:-r2_last}
set -x
python -m pytest tests -q -m gpu 2>&1 | tail -3 > gpurun_out/tests_$T.log
python bench.py > gpurun_out/bench_$T.json 2> gpurun_out/bench_$T.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_${T}_reference_arm.json 2>/dev/null
python bench.py --steps 2 --warmup 3 --no-families --no-inference --no-cpu-baseline > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 9000 --csv --log-file gpurun_out/launches_$T.csv python bench.py --steps 2 --warmup 3 --no-families --no-inference --no-cpu-baseline > gpurun_out/ncu_final.log 2>&1
python bench.py --config cfg5 --steps 5 --warmup 3 --no-cpu-baseline --no-inference > gpurun_out/bench_${T}_cfg5_n1.json 2>/dev/null
