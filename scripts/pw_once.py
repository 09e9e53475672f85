"""A few launches of the 1x1x1 dgrad-with-fused-add conv (16 -> 32 @2x128^3, the attention-gate W_x dgrad) for ncu."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import b3d  # noqa
from unet3d_b200 import ops
dev = torch.device("cuda:0")
torch.manual_seed(0)
cin, cout = int(os.environ.get("PW_CIN", "16")), int(os.environ.get("PW_COUT", "32"))
x = torch.randn(2, 128, 128, 128, cin, device=dev).to(torch.bfloat16)
w = torch.randn(cout, cin, 1, 1, 1, device=dev) * 0.1
wp, kp, rows = ops.pack_weight(w, ops.PACK_FPROP)
dx = torch.randn(2, 128, 128, 128, cout, device=dev).to(torch.bfloat16)
add = os.environ.get("PW_ADD", "1") == "1"
for _ in range(4):
    ops.conv_fprop(x, wp, rows, cout, 1, out=dx, add=dx if add else None)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    ops.conv_fprop(x, wp, rows, cout, 1, out=dx, add=dx if add else None)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
gb = 2 * 128 ** 3 * (cin + cout * (2 if add else 1)) * 2 / 1e9
print("conv1 %d->%d add=%s: %.3f ms, %.0f GB/s" % (cin, cout, add, ms, gb / ms * 1e3))
