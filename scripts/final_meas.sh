# Round-end measurement set on ONE GPU (run under gpurun): tests, both bench arms, launch lists of the same commands.
set -x
python -m pytest tests -q -m gpu 2>&1 | tail -3 > gpurun_out/tests_r2_final.log
python bench.py > gpurun_out/bench_r2_final.json 2> gpurun_out/bench_r2_final.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r2_final_reference_arm.json 2>/dev/null
python bench.py --steps 2 --warmup 3 --no-families --no-inference --no-cpu-baseline > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 9000 --csv --log-file gpurun_out/launches_r2_final.csv python bench.py --steps 2 --warmup 3 --no-families --no-inference --no-cpu-baseline > gpurun_out/ncu_final.log 2>&1
python scripts/infer_once.py > gpurun_out/plain_inf.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_r2_inference.csv python scripts/infer_once.py > gpurun_out/ncu_inf.log 2>&1
python bench.py --config cfg5 --steps 5 --warmup 3 --no-cpu-baseline --no-inference > gpurun_out/bench_r2_final_cfg5_n1.json 2>/dev/null
python bench.py --scaling strong --steps 5 --warmup 3 --no-cpu-baseline --no-inference --no-families > gpurun_out/bench_r2_final_n1_strong.json 2>/dev/null
