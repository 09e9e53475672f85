# usage (under gpurun): bash scripts/ncu_capture.sh <target of scripts/ncu_targets.py> <kernel-name regex> [launch-skip]
# plain run first (must exit 0), then ONE `ncu --set full` capture of the matching kernel; report -> gpurun_out/ncu_<target>.ncu-rep
T=$1; K=$2; S=${3:-2}
python scripts/ncu_targets.py $T > gpurun_out/ncu_${T}_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/ncu_${T}_plain.log; exit 1; }
tail -2 gpurun_out/ncu_${T}_plain.log
ncu --set full --clock-control none --import-source on -k regex:$K --launch-skip $S --launch-count 1 -f -o gpurun_out/ncu_$T python scripts/ncu_targets.py $T > gpurun_out/ncu_${T}.log 2>&1
tail -3 gpurun_out/ncu_${T}.log
