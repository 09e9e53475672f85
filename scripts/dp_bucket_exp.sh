# data-parallel bucket-size sweep at N = 2 (cfg 3): B3D_BUCKET_MB in the list given on the command line
B="--steps 20 --warmup 5 --no-families --no-inference --no-cpu-baseline"
for mb in "$@"; do
  B3D_BUCKET_MB=$mb python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 $B > gpurun_out/dp_bucket_$mb.json 2> gpurun_out/dp_bucket_$mb.err
  python -c "import json;d=json.load(open('gpurun_out/dp_bucket_$mb.json'));print('bucket_mb=$mb', d['ms_per_step'], d['e2e']['ms_per_step'])"
done
