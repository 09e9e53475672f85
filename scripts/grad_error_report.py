"""Per-tensor gradient error of the bf16 CUDA path vs the fp32 oracle, next to the oracle's OWN bf16-autocast error
(same model, same inputs) — the error bar that the reference's numerics allow (SURVEY hard part 5)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
import b3d  # noqa
import unet3d_b200 as U
from oracle import unet3d_oracle as O
DEV = "cuda:0"
feats = (16, 32, 64, 128, 256)
size = int(sys.argv[1]) if len(sys.argv) > 1 else 32
sd = O.make_state_dict(4, 4, feats, seed=3)
x, y = O.make_inputs(2, size, size, size, seed=3)

def rel(a, b):
    a, b = a.double().reshape(-1), b.double().reshape(-1)
    return float((a - b).norm() / (b.norm() + 1e-30))

def oracle(dev, autocast, masks):
    sdg = {k: v.clone().to(dev).requires_grad_(v.is_floating_point() and "running" not in k) for k, v in sd.items()}
    with torch.autocast(dev.split(":")[0], dtype=torch.bfloat16, enabled=autocast):
        main, deep, bn = O.unet_forward(x.to(dev), sdg, feats, training=True, dropout_masks=[m.to(dev) for m in masks])
        loss = O.deep_supervision_loss(main.float(), [d.float() for d in deep], y.to(dev))
    loss.backward()
    return {k: (v.grad.cpu() if v.grad is not None else None) for k, v in sdg.items()}, float(loss)

for rep in range(2):
    model = U.UNet3D(4, 4, features=list(feats), dropout_rate=0.2)
    model.load_state_dict(sd); model = model.to(DEV).train()
    torch.manual_seed(99)
    main, deep = model(x.to(DEV))
    masks = [m.cpu() for m in model._last_dropout_masks]
    loss = U.DeepSupervisionLoss3D()((main, deep), y.to(DEV)); loss.backward()
    if rep == 0:
        g32, l32 = oracle("cpu", False, masks)
        g16, l16 = oracle(DEV, True, masks)
        tot = sum(float(g.double().norm()) ** 2 for g in g32.values() if g is not None) ** 0.5
    rows = []
    for k, p in model.named_parameters():
        if g32[k] is None or float(g32[k].double().norm()) < 1e-3 * tot: continue
        rows.append((rel(p.grad.cpu(), g32[k]), rel(g16[k], g32[k]), k))
    rows.sort(reverse=True)
    print("run %d: loss ours %.5f fp32 %.5f autocast %.5f" % (rep, float(loss), l32, l16))
    for a, b, k in rows[:8]:
        print("   %-44s ours %.4f   oracle-bf16-autocast %.4f   ratio %.2f" % (k, a, b, a / (b + 1e-12)))
    import statistics
    print("   median ours %.4f  autocast %.4f ; tensors where ours > autocast: %d / %d" % (
        statistics.median(r[0] for r in rows), statistics.median(r[1] for r in rows), sum(r[0] > r[1] for r in rows), len(rows)))
