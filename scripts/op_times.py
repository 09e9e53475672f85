"""CUDA-event time of EVERY library call of one eager cfg-3 training step (single stream: each kernel runs alone), grouped by
(op, tensor shapes).  Usage: python scripts/op_times.py [batch] [size] > gpurun_out/op_times.txt"""
import collections, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import b3d  # noqa
import unet3d_b200 as U
from unet3d_b200 import ops

N = int(sys.argv[1]) if len(sys.argv) > 1 else 2
S = int(sys.argv[2]) if len(sys.argv) > 2 else 128
dev = torch.device("cuda:0")
REC = []
SKIP = {"roundup", "ld", "new_act", "zeros_scratch", "reset_scratch", "loss_cfg", "wgrad_join", "conv_wgrad", "convT2_wgrad",
        "pack_weight", "side_branch", "check", "ptr", "stream_ptr"}


def shapes(a):
    out = []
    for t in a:
        if torch.is_tensor(t) and t.dim() >= 4:
            out.append("x".join(str(v) for v in t.shape))
        if len(out) >= 2:
            break
    return " ".join(out)


def wrap(name):
    fn = getattr(ops, name)

    def w(*a, **k):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn(*a, **k)
        e1.record()
        REC.append((name, shapes(a), e0, e1))
        return out
    setattr(ops, name, w)


for nme in dir(ops):
    f = getattr(ops, nme)
    if callable(f) and not nme.startswith("__") and nme not in SKIP and getattr(f, "__module__", "") == ops.__name__ and not isinstance(f, type):
        if nme.startswith("_") and nme not in ("_conv_wgrad", "_convT2_wgrad"):
            continue
        wrap(nme)
ops.WGRAD_SIDE = False
torch.manual_seed(0)
model = U.UNet3D(4, 4).to(dev).train()
crit = U.DeepSupervisionLoss3D()
opt = U.make_adamw(model, capturable=False)
x = torch.randn(N, 4, S, S, S, device=dev)
y = torch.randint(0, 4, (N, S, S, S), device=dev)


def step():
    opt.zero_grad(set_to_none=True)
    loss = crit(model(x), y)
    loss.backward()
    opt.step()


for _ in range(3):
    step()
torch.cuda.synchronize()
REC.clear()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda._sleep(int(0.04 * 1.9e9))   # keep the device behind the host: event pairs then bracket device time only
e0.record(); step(); e1.record()
torch.cuda.synchronize()
agg = collections.OrderedDict()
byname = collections.OrderedDict()
for name, sh, a, b in REC:
    ms = a.elapsed_time(b)
    c = agg.setdefault((name, sh), [0, 0.0]); c[0] += 1; c[1] += ms
    c = byname.setdefault(name, [0, 0.0]); c[0] += 1; c[1] += ms
tot = sum(v[1] for v in byname.values())
print("eager step %.2f ms; library calls %.2f ms in %d calls" % (e0.elapsed_time(e1), tot, len(REC)))
print("---- by op")
for name, (c, ms) in sorted(byname.items(), key=lambda kv: -kv[1][1]):
    print("%-22s x%-3d %8.3f ms  %5.1f%%" % (name, c, ms, 100 * ms / tot))
print("---- by op and shape")
for (name, sh), (c, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:90]:
    print("%-18s %-44s x%-2d %8.3f ms" % (name, sh, c, ms))
