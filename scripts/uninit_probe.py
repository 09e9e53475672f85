"""Hunt reads of uninitialised memory / run-to-run differences: run the same forward (and optionally backward) several times,
poisoning the caching allocator's free blocks with NaN (or with finite garbage) in between, and report the first op whose
output differs from the first run.  Usage: python scripts/uninit_probe.py [eval|train] [N] [size] [nan|big|none]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import b3d  # noqa
import unet3d_b200 as U
from unet3d_b200 import ops
from oracle import unet3d_oracle as O

mode = sys.argv[1] if len(sys.argv) > 1 else "eval"
N = int(sys.argv[2]) if len(sys.argv) > 2 else 1
size = int(sys.argv[3]) if len(sys.argv) > 3 else 128
poison_kind = sys.argv[4] if len(sys.argv) > 4 else "nan"
feats = tuple(int(v) for v in os.environ.get("FEATS", "32,64,128,256,512").split(","))
dev = torch.device("cuda:0")
LOG = []
NAMES = ["conv_fprop", "convT2_fprop", "convT2_dgrad", "_conv_wgrad", "_convT2_wgrad", "gn_apply", "gn_bwd", "gn_bwd_dual", "add_bf16",
         "pool_fwd", "pool_bwd", "to_ndhwc_bf16", "channel_sum", "gate_psi_fwd", "gate_se_fwd", "gate_apply_fwd", "gate_apply_bwd",
         "gate_se_bwd", "gate_psi_bwd", "add_channel_const", "ds_head_fwd", "ds_head_bwd", "trilinear_up_fwd", "trilinear_up_bwd",
         "final_bn_prepare", "final_head_fwd", "final_head_bwd", "loss_fwd", "loss_bwd"]


def digest(t):
    if not torch.is_tensor(t) or t.numel() == 0:
        return None
    f = t.detach().double() if t.is_floating_point() else t.detach().double()
    return (tuple(t.shape), float(torch.nan_to_num(f, nan=1e30).sum()), float(torch.nan_to_num(f, nan=1e30).abs().sum()), bool(torch.isnan(f).any()))


def flat(o):
    if torch.is_tensor(o):
        return [o]
    if isinstance(o, (tuple, list)):
        r = []
        for e in o:
            r += flat(e)
        return r
    return []


def wrap(name):
    fn = getattr(ops, name)

    def w(*a, **k):
        out = fn(*a, **k)
        LOG.append((name, [digest(t) for t in flat(out)]))
        return out
    setattr(ops, name, w)


for nme in NAMES:
    wrap(nme)
ops.WGRAD_SIDE = False


def poison():
    torch.cuda.synchronize()
    val = float("nan") if poison_kind == "nan" else 3.0e4
    blocks = []
    for sz in (1 << 30, 1 << 28, 1 << 26, 1 << 24, 1 << 22, 1 << 20, 1 << 18, 1 << 16, 1 << 14, 1 << 12, 1 << 10):
        for _ in range(6 if sz >= (1 << 28) else 24):
            try:
                blocks.append(torch.full((sz // 2,), val, dtype=torch.bfloat16, device=dev))
            except RuntimeError:
                break
    del blocks
    torch.cuda.synchronize()


sd = O.make_state_dict(4, 4, feats, seed=32)
x, y = O.make_inputs(N, size, size, size, seed=32)
model = U.UNet3D(4, 4, features=list(feats), dropout_rate=0.0)
model.load_state_dict(sd)
model = model.to(dev)
xd, yd = x.to(dev), y.to(dev)
crit = U.DeepSupervisionLoss3D()
runs = []
for it in range(4):
    LOG.clear()
    if it >= 2 and poison_kind != "none":
        poison()
    if mode == "eval":
        model.eval()
        with torch.no_grad():
            out = model(xd)
        LOG.append(("OUT", [digest(out)]))
    else:
        model.train()
        model.zero_grad(set_to_none=True)
        loss = crit(model(xd), yd)
        loss.backward()
        LOG.append(("LOSS", [digest(loss)]))
        for k, p in model.named_parameters():
            if p.grad is not None:
                LOG.append(("grad " + k, [digest(p.grad)]))
    torch.cuda.synchronize()
    runs.append(list(LOG))
    print("run %d: %d records%s" % (it, len(LOG), " (after poison)" if it >= 2 and poison_kind != "none" else ""), flush=True)

base = runs[0]
for it in range(1, len(runs)):
    r = runs[it]
    ndiff = 0
    for i, (a, b) in enumerate(zip(base, r)):
        if a != b:
            ndiff += 1
            if ndiff <= 6:
                print("run %d differs at record %d: %s\n   base %s\n   this %s" % (it, i, a[0], a[1], b[1]))
    print("run %d: %d of %d records differ" % (it, ndiff, len(base)), flush=True)
