"""Repeat the default-architecture forward (+loss) many times on identical inputs and report run-to-run deviations."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import b3d  # noqa
import unet3d_b200 as U
from oracle import unet3d_oracle as O
DEV = "cuda:0"
size = int(sys.argv[1]) if len(sys.argv) > 1 else 32
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 100
feats = (32, 64, 128, 256, 512)
sd = O.make_state_dict(4, 4, feats, seed=0)
x, y = O.make_inputs(2, size, size, size, seed=0)
model = U.UNet3D(4, 4, features=list(feats), dropout_rate=0.0)
model.load_state_dict(sd); model = model.to(DEV).train()
crit = U.DeepSupervisionLoss3D()
xd, yd = x.to(DEV), y.to(DEV)
ref_loss = None; ref_main = None
worst = 0.0
for it in range(iters):
    main, deep = model(xd)
    loss = float(crit((main, deep), yd))
    if ref_loss is None:
        ref_loss, ref_main = loss, main.detach().clone()
        print("reference loss", loss)
        continue
    dl = abs(loss - ref_loss) / abs(ref_loss)
    dm = float((main.detach() - ref_main).abs().max())
    worst = max(worst, dl)
    if dl > 1e-3 or dm > 0.05:
        print("iter %d: loss %.6f (rel dev %.3g), max |dlogit| %.4g" % (it, loss, dl, dm), flush=True)
print("worst relative loss deviation %.3g over %d runs" % (worst, iters))
