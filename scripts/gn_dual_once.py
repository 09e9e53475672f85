"""A few launches of the dual GroupNorm backward at the level-0 shape (2 x 128^3 x 32 channels) for the ncu --set full capture."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import b3d  # noqa
from unet3d_b200 import ops
dev = torch.device("cuda:0")
torch.manual_seed(0)
n, s, c, groups = 2, 128, 32, 8
ya = torch.randn(n, s, s, s, c, device=dev).to(torch.bfloat16)
yb = torch.randn(n, s, s, s, c, device=dev).to(torch.bfloat16)
dy = torch.randn(n, s, s, s, c, device=dev).to(torch.bfloat16)
ga, ba, gb = torch.ones(c, device=dev), torch.zeros(c, device=dev), torch.ones(c, device=dev)


def stats(y):
    yf = y.float().reshape(n, -1, groups, c // groups)
    return torch.stack([yf.sum(dim=(1, 3)), (yf * yf).sum(dim=(1, 3))], dim=-1).double()


sta, stb = stats(ya), stats(yb)
for _ in range(3):
    out = ops.gn_bwd_dual(dy, ya, sta, ga, ba, yb, stb, gb, groups)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    ops.gn_bwd_dual(dy, ya, sta, ga, ba, yb, stb, gb, groups)
e1.record(); torch.cuda.synchronize()
T = n * s ** 3 * c * 2 / 1e9
print("gn_bwd_dual 2x128^3x32: %.3f ms per call (reduce 3T + apply 5T = %.2f GB -> %.0f GB/s)" % (
    e0.elapsed_time(e1) / 5, 8 * T, 8 * T / (e0.elapsed_time(e1) / 5) * 1e3))
