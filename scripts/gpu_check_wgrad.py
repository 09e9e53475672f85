"""GPU check + timing of the conv weight-gradient kernels against torch autograd (fp32 math on bf16-rounded inputs)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
import b3d  # noqa
from unet3d_b200 import ops
dev = torch.device("cuda:0")

def case(n, d, h, w, cin, cout, ks=3, iters=0, cin_real=None):
    torch.manual_seed(0)
    cin_real = cin_real or cin
    x = torch.randn(n, d, h, w, cin, device=dev).to(torch.bfloat16)
    if cin_real < cin: x[..., cin_real:] = 0
    dy = torch.randn(n, d, h, w, cout, device=dev).to(torch.bfloat16)
    dw = ops.conv_wgrad(x, dy, cin_real, cout, ks)
    torch.cuda.synchronize()
    xr = x[..., :cin_real].float().permute(0, 4, 1, 2, 3).contiguous()
    wt = torch.zeros(cout, cin_real, ks, ks, ks, device=dev, requires_grad=True)
    y = F.conv3d(xr, wt, padding=ks // 2)
    y.backward(dy.float().permute(0, 4, 1, 2, 3))
    ref = wt.grad
    err = (dw - ref).abs().max().item(); sc = ref.abs().max().item()
    ok = err <= 2e-3 * sc + 1e-3
    msg = "wgrad%d N%d %dx%dx%d %d,%d: maxdiff %.4g (scale %.4g) %s" % (ks, n, d, h, w, cin_real, cout, err, sc, "OK" if ok else "FAIL")
    if iters:
        for _ in range(2): ops.conv_wgrad(x, dy, cin_real, cout, ks)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters): ops.conv_wgrad(x, dy, cin_real, cout, ks)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        msg += "  %.3f ms  %.1f TFLOP/s" % (ms, 2.0 * n * d * h * w * cin_real * cout * ks ** 3 / ms / 1e9)
    print(msg, flush=True)
    return ok

if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    ok = True
    if which in ("all", "small"):
        ok &= case(1, 4, 8, 32, 32, 32)
        ok &= case(2, 5, 7, 48, 32, 32)
        ok &= case(1, 3, 16, 16, 16, 32, cin_real=4)
        ok &= case(2, 6, 9, 64, 32, 16)
        ok &= case(1, 8, 16, 32, 64, 64)
        ok &= case(1, 2, 8, 16, 128, 96)
        ok &= case(2, 1, 4, 16, 32, 32)
        ok &= case(1, 9, 33, 128, 16, 16)
    if which == "one":
        ok &= case(2, 128, 128, 128, 32, 32, iters=3)
    if which in ("all", "perf"):
        ok &= case(2, 128, 128, 128, 32, 32, iters=3)
        ok &= case(2, 128, 128, 128, 64, 32, iters=3)
        ok &= case(2, 128, 128, 128, 16, 32, iters=3, cin_real=4)
        ok &= case(2, 128, 128, 128, 32, 16, iters=3)
        ok &= case(2, 64, 64, 64, 64, 64, iters=3)
        ok &= case(2, 64, 64, 64, 128, 64, iters=3)
        ok &= case(2, 32, 32, 32, 128, 128, iters=3)
        ok &= case(2, 32, 32, 32, 256, 128, iters=3)
        ok &= case(2, 16, 16, 16, 256, 256, iters=3)
        ok &= case(2, 16, 16, 16, 512, 256, iters=3)
        ok &= case(2, 8, 8, 8, 512, 512, iters=3)
        ok &= case(2, 4, 4, 4, 1024, 1024, iters=3)
    print("ALL OK" if ok else "SOME FAILED")
    sys.exit(0 if ok else 1)
