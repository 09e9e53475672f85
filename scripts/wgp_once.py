"""A few launches of the streaming pointwise weight-gradient kernel at the level-0 shape (for the ncu --set full capture)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import b3d  # noqa
from unet3d_b200 import ops
dev = torch.device("cuda:0")
torch.manual_seed(0)
x = torch.randn(2, 128, 128, 128, 64, device=dev).to(torch.bfloat16)
dy = torch.randn(2, 128, 128, 128, 32, device=dev).to(torch.bfloat16)
for _ in range(4):
    dw = ops.conv_wgrad(x, dy, 64, 32, 1)
torch.cuda.synchronize()
print("ok", float(dw.abs().sum()))
