set -x
B="--steps 20 --warmup 5 --no-families --no-inference --no-cpu-baseline"
for r in 0 8 16 4; do
  B3D_DP_RESERVED_SMS=$r python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 $B > gpurun_out/dp_res_$r.json 2> gpurun_out/dp_res_$r.err
  python -c "import json;d=json.load(open('gpurun_out/dp_res_$r.json'));print('R=$r', d['ms_per_step'], d['e2e']['ms_per_step'])"
done
B3D_RESERVED_SMS=8 python bench.py $B > gpurun_out/n1_res8.json 2>/dev/null; python -c "import json;d=json.load(open('gpurun_out/n1_res8.json'));print('N1 R=8', d['ms_per_step'])"
python bench.py $B > gpurun_out/n1_res0.json 2>/dev/null; python -c "import json;d=json.load(open('gpurun_out/n1_res0.json'));print('N1 R=0', d['ms_per_step'])"
