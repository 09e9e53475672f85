"""Race detector: the conv kernels accumulate in a fixed order, so repeated launches must give BIT-IDENTICAL outputs."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import b3d  # noqa
from unet3d_b200 import ops
dev = "cuda:0"
BF = torch.bfloat16
torch.manual_seed(0)
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 100
bad = 0
for (n, d, h, w, cin, cout, ks) in [(2, 64, 64, 64, 32, 32, 3), (2, 64, 64, 64, 64, 32, 3), (2, 32, 32, 32, 16, 32, 3), (2, 32, 32, 32, 32, 16, 3),
                                    (2, 32, 32, 32, 64, 64, 3), (2, 32, 32, 32, 64, 128, 3), (2, 16, 16, 16, 128, 128, 3), (2, 8, 8, 8, 256, 256, 3),
                                    (2, 64, 64, 64, 64, 32, 1), (2, 4, 4, 4, 512, 512, 3)]:
    x = torch.randn(n, d, h, w, cin, device=dev).to(BF)
    wt = (torch.randn(cout, cin, ks, ks, ks, device=dev) / (cin * ks ** 3) ** 0.5).to(BF).float()
    wp, kp, rows = ops.pack_weight(wt, ops.PACK_FPROP)
    y0, st0 = ops.conv_fprop(x, wp, rows, cout, ks, groups=8)
    y0 = y0.clone(); torch.cuda.synchronize()
    mism = 0
    for it in range(iters):
        y, st = ops.conv_fprop(x, wp, rows, cout, ks, groups=8)
        if not torch.equal(y, y0):
            mism += 1
            if mism == 1:
                dmax = (y.float() - y0.float()).abs().max().item()
                print("   first mismatch at iter %d: max |diff| %.4g, %d elements differ" % (it, dmax, int((y != y0).sum())))
        srel = ((st - st0).abs() / (st0.abs() + 1)).max().item()
        if srel > 1e-3:
            mism += 1
            print("   stats mismatch rel %.3g at iter %d" % (srel, it))
    print("conv%d %dx%dx%dx%d %d->%d: %d / %d launches differ" % (ks, n, d, h, w, cin, cout, mism, iters), flush=True)
    bad += mism
# weight gradients: fp32 atomics make the flush order-dependent, so compare with a tolerance
for (n, d, h, w, cin, cout) in [(2, 64, 64, 64, 32, 32), (2, 32, 32, 32, 64, 64), (2, 16, 16, 16, 128, 128)]:
    x = torch.randn(n, d, h, w, cin, device=dev).to(BF)
    dy = torch.randn(n, d, h, w, cout, device=dev).to(BF)
    d0 = ops.conv_wgrad(x, dy, cin, cout, 3).clone()
    mism = 0
    for it in range(iters):
        dw = ops.conv_wgrad(x, dy, cin, cout, 3)
        r = ((dw - d0).abs().max() / d0.abs().max()).item()
        if r > 1e-4:
            mism += 1
            if mism == 1: print("   wgrad mismatch rel %.3g at iter %d" % (r, it))
    print("wgrad3 %dx%dx%dx%d %d,%d: %d / %d launches differ" % (n, d, h, w, cin, cout, mism, iters), flush=True)
    bad += mism
print("TOTAL MISMATCHES", bad)
sys.exit(1 if bad else 0)
