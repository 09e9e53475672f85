"""Where do two identical forward passes first differ?  Walks the saved-for-backward state of unet_fwd (every conv output,
statistics block and activation, in execution order) of two runs and prints, per tensor, whether it is bit-equal and how
large the difference is.  A legitimate source is the order of the fp64 statistics atomics (last-ulp changes of a
normalisation coefficient -> isolated 1-ulp bf16 flips that then spread); a race would show up as a large first difference
in a conv output."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import b3d  # noqa
import unet3d_b200 as U
from unet3d_b200 import functional as Fn
dev = torch.device("cuda", 0)
S = int(os.environ.get("JIT_SIZE", "64"))
torch.manual_seed(0)
model = U.UNet3D(4, 4, features=[16, 32, 64, 128, 256], dropout_rate=0.0).to(dev).train()
g = torch.Generator().manual_seed(100)
x = torch.randn(1, 4, S, S, S, generator=g).to(dev)
p = dict(model.named_parameters())
bufs = {k: v.clone() for k, v in model.named_buffers()}


def flat(o, path, out):
    if torch.is_tensor(o):
        out.append((path, o))
    elif isinstance(o, (list, tuple)):
        for i, e in enumerate(o):
            flat(e, "%s.%d" % (path, i), out)
    elif isinstance(o, dict):
        for k in o:
            flat(o[k], "%s.%s" % (path, k), out)


def run():
    b = {k: v.clone() for k, v in bufs.items()}
    with torch.no_grad():
        logits, deep, saved = Fn.unet_fwd(x, p, b, list(model.features), True, None, True)
    torch.cuda.synchronize()
    out = []
    flat(saved, "S", out)
    out.append(("logits", logits))
    return [(k, t.detach().clone()) for k, t in out]


base = run()
for rep in range(int(os.environ.get("REPS", "1"))):
    other = run()
    print("=== run %d vs run 0" % (rep + 1))
    seen = set()
    shown = 0
    for (k, a), (_, b) in zip(base, other):
        key = (a.data_ptr(), tuple(a.shape))
        if torch.equal(a, b) or key in seen or (a.dtype == torch.float64 and os.environ.get('SKIP_STATS', '1') == '1'):
            continue
        seen.add(key)
        af, bf = a.double(), b.double()
        nd = int((af != bf).sum())
        print("  %-28s %-22s %-8s differs in %d / %d elements, max abs %.3g (max |a| %.3g)" % (
            k, tuple(a.shape), str(a.dtype).replace("torch.", ""), nd, a.numel(), float((af - bf).abs().max()), float(af.abs().max())))
        shown += 1
        if shown >= 80:
            break
    if shown == 0:
        print("  bit-identical")
