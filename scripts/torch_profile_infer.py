"""Per-kernel GPU time of one eval forward (batch 1, 4x128^3) via torch.profiler."""
import os, sys, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
import b3d  # noqa
import unet3d_b200 as U
dev = torch.device("cuda:0")
torch.manual_seed(0)
model = U.UNet3D(4, 4).to(dev).eval()
x = torch.randn(1, 4, 128, 128, 128, device=dev)
with torch.no_grad():
    for _ in range(3): model(x)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        model(x); torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
agg = collections.defaultdict(lambda: [0, 0.0])
for e in ev:
    k = e.name.split("(")[0][:64]; agg[k][0] += 1; agg[k][1] += e.device_time_total
tot = sum(v[1] for v in agg.values())
print("GPU span %.3f ms, kernel sum %.3f ms, %d kernels" % ((max(e.time_range.end for e in ev) - min(e.time_range.start for e in ev)) / 1e3, tot / 1e3, len(ev)))
for k, (c, us) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:25]:
    print("%-66s %4d %8.3f ms %5.1f%%" % (k, c, us / 1e3, 100 * us / tot))
