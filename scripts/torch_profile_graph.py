"""Per-kernel GPU time INSIDE the replayed CUDA graph of the cfg-3 training step (fwd + DS loss + bwd + fused AdamW + weight
re-pack), via torch.profiler (CUPTI sees the kernels of a graph replay).  This is the step bench.py times."""
import os, sys, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
import b3d  # noqa
import unet3d_b200 as U
dev = torch.device("cuda:0")
torch.manual_seed(0)
model = U.UNet3D(4, 4, dropout_rate=0.2).to(dev).train()
crit = U.DeepSupervisionLoss3D()
opt = torch.optim.AdamW(model.parameters(), lr=1e-4, weight_decay=1e-4, fused=True, capturable=True)
x = torch.randn(2, 4, 128, 128, 128, device=dev)
y = torch.randint(0, 4, (2, 128, 128, 128), device=dev)
step = U.GraphedTrainStep(model, crit, opt, x, y, warmup=3)
for _ in range(3): step(x, y)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    step(x, y)
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
agg = collections.defaultdict(lambda: [0, 0.0])
t0 = min(e.time_range.start for e in ev); t1 = max(e.time_range.end for e in ev)
for e in ev:
    k = e.name.split("(")[0][:64]
    agg[k][0] += 1; agg[k][1] += e.device_time_total if hasattr(e, "device_time_total") else e.cuda_time_total
tot = sum(v[1] for v in agg.values())
print("GPU span %.3f ms, sum of kernel time %.3f ms, %d kernels" % ((t1 - t0) / 1e3, tot / 1e3, len(ev)))
for k, (c, us) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:int(sys.argv[1]) if len(sys.argv) > 1 else 70]:
    print("%-66s %4d %8.3f ms %5.1f%%" % (k, c, us / 1e3, 100 * us / tot))
