"""Run-to-run reproducibility of the bf16 path on one GPU: forward logits bit-compared over two runs, per-tensor gradient
jitter over two identical backward passes (B3D_WGRAD_STREAM=0/1 to include or exclude the weight-gradient side stream)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import b3d  # noqa
import unet3d_b200 as U
dev = torch.device("cuda", 0)
S = int(os.environ.get("JIT_SIZE", "64"))
torch.manual_seed(0)
model = U.UNet3D(4, 4, features=[16, 32, 64, 128, 256], dropout_rate=0.0).to(dev).train()
crit = U.DeepSupervisionLoss3D()
g = torch.Generator().manual_seed(100)
x = torch.randn(1, 4, S, S, S, generator=g).to(dev)
y = torch.randint(0, 4, (1, S, S, S), generator=g).to(dev)
runs = []
for i in range(3):
    model.zero_grad(set_to_none=True)
    out = model(x)
    loss = crit(out, y)
    loss.backward()
    torch.cuda.synchronize()
    runs.append((out[0].detach().clone(), [o.detach().clone() for o in out[1]], float(loss),
                 {k: p.grad.detach().clone() for k, p in model.named_parameters() if p.grad is not None}))
a = runs[0]
for i in (1, 2):
    b = runs[i]
    print("run %d vs 0: logits bit-equal %s (max abs diff %.3g), ds outputs bit-equal %s, loss %.9g vs %.9g" % (
        i, bool(torch.equal(a[0], b[0])), float((a[0] - b[0]).abs().max()),
        [bool(torch.equal(p, q)) for p, q in zip(a[1], b[1])], a[2], b[2]))
    total = float(torch.sqrt(sum((t.double() ** 2).sum() for t in a[3].values())))
    rows = []
    for k in a[3]:
        d = float((a[3][k] - b[3][k]).norm())
        rows.append((d / max(float(a[3][k].norm()), 1e-3 * total), d / total, k))
    rows.sort(reverse=True)
    print("  whole-model jitter %.3g; worst tensors:" % (sum(r[1] ** 2 for r in rows) ** 0.5))
    for r in rows[:6]:
        print("   %-40s rel %.3g" % (r[2], r[0]))
