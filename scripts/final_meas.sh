set -x
python bench.py > gpurun_out/bench_r2_final.json 2> gpurun_out/bench_r2_final.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r2_final_reference_arm.json 2>/dev/null
python bench.py --steps 2 --warmup 3 --no-families --no-inference --no-cpu-baseline > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 9000 --csv --log-file gpurun_out/launches_r2_final.csv python bench.py --steps 2 --warmup 3 --no-families --no-inference --no-cpu-baseline > gpurun_out/ncu_final.log 2>&1
for t in zs:zs_kernel pwadd:igemm_kernel dsloss:dsloss_bwd_kernel adamw:adamw_pack_multi_kernel gnbwd:gn_bwd_dual_apply_kernel; do
  tgt=${t%%:*}; kn=${t##*:}
  python scripts/ncu_targets.py $tgt > gpurun_out/plain_$tgt.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:$kn -s 2 -c 1 -f -o gpurun_out/prof_r2_$tgt python scripts/ncu_targets.py $tgt > gpurun_out/ncu_$tgt.log 2>&1
done
ls -la gpurun_out/*.ncu-rep | tail -8
