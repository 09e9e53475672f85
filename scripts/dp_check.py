"""Data-parallel correctness on NCCL (run under torchrun, 2+ ranks): the gradients DataParallel leaves in .grad equal
the mean over ranks of the gradients each rank computes alone on its own shard (same weights, different data)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import b3d  # noqa
import unet3d_b200 as U
from unet3d_b200.parallel import DataParallel
from unet3d_b200 import _lib
_lib.set_ordered_issue(2)   # bit-reproducible forward / input gradients: what is left is the fp32-atomic jitter of the wgrad flush
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
dist.init_process_group("nccl", device_id=dev)
feats = [16, 32, 64, 128, 256]
torch.manual_seed(0)
model = U.UNet3D(4, 4, features=feats, dropout_rate=0.0).to(dev).train()
crit = U.DeepSupervisionLoss3D()
g = torch.Generator().manual_seed(100 + rank)
S = int(os.environ.get('DP_CHECK_SIZE', '64'))
x = torch.randn(1, 4, S, S, S, generator=g).to(dev)
y = torch.randint(0, 4, (1, S, S, S), generator=g).to(dev)
# local gradients, no communication
crit(model(x), y).backward()
local_g = {k: p.grad.detach().clone() for k, p in model.named_parameters() if p.grad is not None}
model.zero_grad(set_to_none=True)
crit(model(x), y).backward()   # the same local pass again: its difference from the first is the run-to-run jitter floor
local_g2 = {k: p.grad.detach().clone() for k, p in model.named_parameters() if p.grad is not None}
model.zero_grad(set_to_none=True)
dp = DataParallel(model, bucket_mb=1.0)   # small buckets: several all-reduces in flight during backward
crit(dp(x), y).backward()                 # pass 1: records the emission order (gather / all-reduce / scatter path)
torch.cuda.synchronize()
first_pass = {k: p.grad.detach().clone() for k, p in model.named_parameters() if p.grad is not None}
model.zero_grad(set_to_none=True)
crit(dp(x), y).backward()                 # pass 2: persistent buckets, weight gradients written in place, in-place all-reduce
torch.cuda.synchronize()
flat_ranges = [(f.data_ptr(), f.data_ptr() + f.numel() * 4) for f in dp.buckets._flat]
inside = sum(any(a <= p.grad.data_ptr() < b for a, b in flat_ranges) for p in model.parameters() if p.grad is not None)
ngrad = sum(p.grad is not None for p in model.parameters())
print("rank %d: debug: after pass 2: %d grads, %d in buckets, first_pass %d entries, layout %s, views hook %s" % (
    rank, ngrad, inside, len(first_pass), dp.buckets._layout is not None, model._grad_views is not None), flush=True)
d12 = max([0.0] + [float((first_pass[k] - p.grad).norm()) / max(float(first_pass[k].norm()), 1e-12)
          for k, p in model.named_parameters() if p.grad is not None and float(first_pass[k].norm()) > 1e-6])
print("rank %d: pass 2 uses %d persistent buckets; %d of %d .grad tensors alias a bucket; pass 1 vs pass 2 worst rel diff %.3g" % (
    rank, len(dp.buckets._flat), inside, ngrad, d12), flush=True)
# Per-tensor deviation is measured against max(|ref|, 1e-3 * |all gradients|): the biases of convs that feed a GroupNorm
# have a mathematically zero gradient (the normalisation removes any per-channel shift), so what the kernels leave there is
# rounding noise that differs from run to run (fp32 atomics) and must not be divided by its own tiny norm.
refs = {}
for k, p in model.named_parameters():
    if k not in local_g:
        continue
    ref = local_g[k].clone()
    dist.all_reduce(ref, op=dist.ReduceOp.SUM)
    refs[k] = ref / world
total = float(torch.sqrt(sum((r.double() ** 2).sum() for r in refs.values())))
worst, worst_k, sq = 0.0, None, 0.0
for k, p in model.named_parameters():
    if k not in refs:
        continue
    num = float((p.grad - refs[k]).norm())
    sq += num * num
    dev = num / max(float(refs[k].norm()), 1e-3 * total)
    if dev > worst:
        worst, worst_k = dev, k
whole = sq ** 0.5 / total
jit, jit_k = 0.0, None
for k in refs:
    d = float((local_g2[k] - local_g[k]).norm()) / max(float(local_g[k].norm()), 1e-3 * total)
    if d > jit:
        jit, jit_k = d, k
print("rank %d: run-to-run jitter of the local backward: worst tensor %.3g (%s)" % (rank, jit, jit_k), flush=True)
print("rank %d: DP gradients vs mean of local gradients: whole-model rel-L2 %.3g, worst tensor %.3g (%s), %d buckets" % (
    rank, whole, worst, worst_k, dp.buckets.buckets_launched), flush=True)
ok = worst < max(1e-4, 10 * jit) and whole < 1e-4 and d12 < max(1e-4, 10 * jit) and len(dp.buckets._flat) >= 2 and inside == ngrad   # dX path is order-independent; only the fp32 weight-gradient flush and the all-reduce order differ
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
