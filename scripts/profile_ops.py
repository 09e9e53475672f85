"""Per-call CUDA-event timing of the tensor-core ops (conv fprop/dgrad/convT/wgrad) of one cfg-3 training step."""
import os, sys, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import b3d  # noqa
import unet3d_b200 as U
from unet3d_b200 import ops
dev = torch.device("cuda:0")
torch.manual_seed(0)
model = U.UNet3D(4, 4).to(dev).train()
crit = U.DeepSupervisionLoss3D()
x = torch.randn(2, 4, 128, 128, 128, device=dev)
y = torch.randint(0, 4, (2, 128, 128, 128), device=dev)
def step():
    model.zero_grad(set_to_none=True)
    loss = crit(model(x), y)
    loss.backward()
for _ in range(2): step()
torch.cuda.synchronize()
ops.PROFILE = []
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); step(); e1.record()
torch.cuda.synchronize()
rows = [(fam, fl, a.elapsed_time(b), tag) for fam, fl, a, b, tag in ops.PROFILE]
ops.PROFILE = None
agg = collections.OrderedDict()
for fam, fl, ms, tag in rows:
    k = (fam, tag)
    c = agg.setdefault(k, [0, 0.0, 0.0]); c[0] += 1; c[1] += ms; c[2] += fl
tot = sum(r[2] for r in rows)
print("step %.2f ms (with profiling events), tensor-core ops %.2f ms in %d calls" % (e0.elapsed_time(e1), tot, len(rows)))
for (fam, tag), (c, ms, fl) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-6s %-34s x%-2d %8.3f ms  %7.1f TFLOP/s" % (fam, tag, c, ms, fl / ms / 1e9))
