"""Timing + check of the streaming pointwise / transposed-conv weight-gradient kernel (conv_wgp.cu) at the cfg-3 shapes.
B3D_NO_WGP=1 runs the generic kernel (conv_wgrad.cu) for comparison."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
import b3d  # noqa
from unet3d_b200 import ops
dev = torch.device("cuda:0")


def timeit(fn, iters=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def pw(n, s, cin, cout, check=True):
    torch.manual_seed(0)
    x = torch.randn(n, s, s, s, cin, device=dev).to(torch.bfloat16)
    dy = torch.randn(n, s, s, s, cout, device=dev).to(torch.bfloat16)
    dw = ops.conv_wgrad(x, dy, cin, cout, 1)
    ref = torch.einsum("vi,vo->oi", x.reshape(-1, cin).float(), dy.reshape(-1, cout).float()).reshape(cout, cin, 1, 1, 1)
    err = (dw - ref).abs().max().item(); sc = ref.abs().max().item()
    ms = timeit(lambda: ops.conv_wgrad(x, dy, cin, cout, 1))
    gb = n * s ** 3 * (cin + cout) * 2 / 1e9
    print("wgrad1 %dx%d^3 %d,%d: maxdiff %.3g (scale %.3g) %s  %.3f ms  %.0f GB/s" % (
        n, s, cin, cout, err, sc, "OK" if err <= 2e-3 * sc + 1e-3 else "FAIL", ms, gb / ms * 1e3), flush=True)


def ct(n, s, cin, cout):
    torch.manual_seed(0)
    x = torch.randn(n, s, s, s, cin, device=dev).to(torch.bfloat16)
    dy = torch.randn(n, 2 * s, 2 * s, 2 * s, cout, device=dev).to(torch.bfloat16)
    dw = ops.convT2_wgrad(x, dy, cin, cout)
    wt = torch.zeros(cin, cout, 2, 2, 2, device=dev, requires_grad=True)
    F.conv_transpose3d(x.float().permute(0, 4, 1, 2, 3), wt, None, stride=2).backward(dy.float().permute(0, 4, 1, 2, 3))
    err = (dw - wt.grad).abs().max().item(); sc = wt.grad.abs().max().item()
    ms = timeit(lambda: ops.convT2_wgrad(x, dy, cin, cout))
    gb = n * s ** 3 * (cin + 8 * cout) * 2 / 1e9
    print("wgradT %dx%d^3 %d,%d: maxdiff %.3g (scale %.3g) %s  %.3f ms  %.0f GB/s" % (
        n, s, cin, cout, err, sc, "OK" if err <= 2e-3 * sc + 1e-3 else "FAIL", ms, gb / ms * 1e3), flush=True)


for a in [(2, 128, 64, 32), (2, 128, 32, 16), (2, 128, 16, 32), (2, 64, 32, 64), (2, 64, 128, 64), (2, 64, 64, 32),
          (2, 32, 64, 128), (2, 32, 256, 128), (2, 32, 128, 64), (2, 16, 128, 256), (2, 16, 512, 256), (2, 16, 256, 128),
          (2, 8, 256, 512), (2, 8, 1024, 512), (2, 8, 512, 256), (2, 4, 512, 1024)]:
    pw(*a)
for a in [(2, 64, 64, 32), (2, 32, 128, 64), (2, 16, 256, 128), (2, 8, 512, 256), (2, 4, 1024, 512)]:
    ct(*a)
