"""GPU bring-up check for the tcgen05 implicit-GEMM conv (run on the B200 box under a timeout).

Compares b3d_conv_fprop / convT2 against torch (cuDNN) on bf16-rounded inputs with fp32 math.
"""
import ctypes
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F

import b3d  # noqa: F401  (registers unet3d_b200)
from unet3d_b200 import _lib

L = _lib.lib()
dev = torch.device("cuda:0")
_lib.require_device(dev)
c_int, c_ll, c_vp, c_sz = _lib.c_int, _lib.c_ll, _lib.c_vp, _lib.c_sz


def pack(w, mode, Kp, rows, ntaps):
    Cout, Cin = (w.shape[0], w.shape[1]) if mode < 2 else (w.shape[1], w.shape[0])
    out = torch.empty((Kp // 8) * (ntaps if mode < 2 else 1) * rows * 8, dtype=torch.bfloat16, device=dev)
    _lib.check(L.b3d_pack_weight(c_int(mode), _lib.ptr(w.contiguous()), c_int(Cout), c_int(Cin), c_int(ntaps), _lib.ptr(out),
                                 c_int(Kp), c_int(rows), _lib.stream_ptr()))
    return out


def ru(a, b):
    return (a + b - 1) // b * b


def conv_case(N, D, H, W, Cin, Cout, ks, groups=8, bias=False, iters=0):
    torch.manual_seed(0)
    x = torch.randn(N, D, H, W, Cin, device=dev).to(torch.bfloat16)
    w = (torch.randn(Cout, Cin, ks, ks, ks, device=dev) / (Cin * ks ** 3) ** 0.5).to(torch.bfloat16).float()
    b = torch.randn(Cout, device=dev) if bias else None
    rows = ru(Cout, 16)
    wp = pack(w.reshape(Cout, Cin, ks ** 3), 0, Cin, rows, ks ** 3)
    y = torch.zeros(N, D, H, W, Cout, device=dev, dtype=torch.bfloat16)
    stats = torch.zeros(N, groups, 2, device=dev, dtype=torch.float64)
    ws = torch.empty(min(16 * N * D * H * W * Cout, 1 << 24), device=dev, dtype=torch.float32)
    err = _lib.err_flag(dev)

    def run():
        _lib.check(L.b3d_conv_fprop(_lib.ptr(x), c_ll(Cin), _lib.ptr(wp), c_int(rows), _lib.ptr(b), _lib.ptr(y), c_ll(Cout),
                                    c_int(N), c_int(D), c_int(H), c_int(W), c_int(Cin), c_int(Cout), c_int(ks),
                                    _lib.ptr(stats), c_int(groups), c_int(0), _lib.ptr(ws), c_sz(ws.numel() * 4),
                                    _lib.ptr(err), _lib.stream_ptr()))

    run()
    torch.cuda.synchronize()
    ref = F.conv3d(x.float().permute(0, 4, 1, 2, 3), w, b, padding=ks // 2).permute(0, 2, 3, 4, 1)
    diff = (y.float() - ref).abs().max().item()
    scale = ref.abs().max().item()
    rs = torch.stack([ref.reshape(N, -1, groups, Cout // groups).sum(dim=(1, 3)),
                      (ref ** 2).reshape(N, -1, groups, Cout // groups).sum(dim=(1, 3))], dim=-1).double()
    sdiff = ((stats - rs).abs() / (rs.abs() + 1.0)).max().item()
    ok = diff <= 2e-2 * max(scale, 1.0) and sdiff < 1e-3
    msg = "conv N%d %dx%dx%d Cin%d Cout%d ks%d: maxdiff %.4g (scale %.3g) stats rel %.3g %s" % (
        N, D, H, W, Cin, Cout, ks, diff, scale, sdiff, "OK" if ok else "FAIL")
    if iters:
        for _ in range(3):
            run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            run()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        fl = 2.0 * N * D * H * W * Cin * Cout * ks ** 3
        msg += "  %.3f ms  %.1f TFLOP/s" % (ms, fl / ms / 1e9)
    print(msg, flush=True)
    return ok


def convT_case(N, D, H, W, Cin, Cout, iters=0):
    torch.manual_seed(1)
    x = torch.randn(N, D, H, W, Cin, device=dev).to(torch.bfloat16)
    w = (torch.randn(Cin, Cout, 2, 2, 2, device=dev) / Cin ** 0.5).to(torch.bfloat16).float()
    b = torch.randn(Cout, device=dev)
    wp = pack(w.reshape(Cin, Cout, 8), 2, Cin, 8 * Cout, 8)
    y = torch.zeros(N, 2 * D, 2 * H, 2 * W, Cout, device=dev, dtype=torch.bfloat16)
    err = _lib.err_flag(dev)
    _lib.check(L.b3d_convT2_fprop(_lib.ptr(x), c_ll(Cin), _lib.ptr(wp), _lib.ptr(b), _lib.ptr(y), c_ll(Cout), c_int(N), c_int(D),
                                  c_int(H), c_int(W), c_int(Cin), c_int(Cout), _lib.ptr(err), _lib.stream_ptr()))
    torch.cuda.synchronize()
    ref = F.conv_transpose3d(x.float().permute(0, 4, 1, 2, 3), w, b, stride=2).permute(0, 2, 3, 4, 1)
    diff = (y.float() - ref).abs().max().item()
    ok = diff <= 2e-2 * max(ref.abs().max().item(), 1.0)
    msg = "convT N%d %dx%dx%d Cin%d Cout%d: maxdiff %.4g %s" % (N, D, H, W, Cin, Cout, diff, "OK" if ok else "FAIL")
    if iters:
        def run():
            _lib.check(L.b3d_convT2_fprop(_lib.ptr(x), c_ll(Cin), _lib.ptr(wp), _lib.ptr(b), _lib.ptr(y), c_ll(Cout), c_int(N), c_int(D),
                                          c_int(H), c_int(W), c_int(Cin), c_int(Cout), _lib.ptr(err), _lib.stream_ptr()))
        for _ in range(3):
            run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            run()
        e1.record()
        torch.cuda.synchronize()
        msg += "  %.3f ms" % (e0.elapsed_time(e1) / iters)
    print(msg, flush=True)
    # dgrad: dx = conv_transpose^T(dy)
    dy = torch.randn(N, 2 * D, 2 * H, 2 * W, Cout, device=dev).to(torch.bfloat16)
    rows = ru(Cin, 16)
    wpd = pack(w.reshape(Cin, Cout, 8), 3, 8 * Cout, rows, 8)
    dx = torch.zeros(N, D, H, W, Cin, device=dev, dtype=torch.bfloat16)
    ws = torch.empty(N * D * H * W * Cin, device=dev, dtype=torch.float32)
    _lib.check(L.b3d_convT2_dgrad(_lib.ptr(dy), c_ll(Cout), _lib.ptr(wpd), c_int(rows), _lib.ptr(dx), c_ll(Cin), c_int(N), c_int(D),
                                  c_int(H), c_int(W), c_int(Cin), c_int(Cout), _lib.ptr(ws), c_sz(ws.numel() * 4),
                                  _lib.ptr(err), _lib.stream_ptr()))
    torch.cuda.synchronize()
    refdx = F.conv3d(dy.float().permute(0, 4, 1, 2, 3), w.permute(0, 1, 2, 3, 4), stride=2).permute(0, 2, 3, 4, 1)
    diff2 = (dx.float() - refdx).abs().max().item()
    ok2 = diff2 <= 2e-2 * max(refdx.abs().max().item(), 1.0)
    print("convT dgrad: maxdiff %.4g %s" % (diff2, "OK" if ok2 else "FAIL"), flush=True)
    return ok and ok2


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    allok = True
    t0 = time.time()
    if which in ("all", "small"):
        allok &= conv_case(1, 8, 8, 8, 16, 16, 3)
        allok &= conv_case(2, 8, 8, 16, 32, 32, 3)
        allok &= conv_case(1, 16, 16, 16, 64, 64, 3, bias=True)
        allok &= conv_case(2, 4, 4, 4, 128, 256, 3)
        allok &= conv_case(1, 8, 8, 8, 32, 16, 1, groups=4, bias=True)
        allok &= conv_case(2, 16, 16, 16, 64, 32, 1, groups=4, bias=True)
        allok &= conv_case(1, 8, 16, 128, 32, 32, 3)
        allok &= conv_case(1, 4, 8, 64, 16, 16, 3, groups=16)
        allok &= convT_case(1, 4, 4, 4, 32, 16)
        allok &= convT_case(2, 8, 8, 8, 64, 32)
    if which in ("all", "zs"):  # shapes that take the z-marching kd-stacked path (conv_zs.cu)
        allok &= conv_case(1, 8, 16, 128, 32, 32, 3)
        allok &= conv_case(2, 16, 8, 128, 64, 32, 3, bias=True)
        allok &= conv_case(1, 8, 12, 64, 16, 16, 3, groups=16)
        allok &= conv_case(1, 8, 16, 64, 32, 16, 3, groups=8, bias=True)
        allok &= conv_case(2, 5, 9, 40, 32, 64, 3)
        allok &= conv_case(1, 16, 16, 16, 64, 64, 3, bias=True)
        allok &= conv_case(1, 6, 16, 24, 128, 64, 3)
        allok &= conv_case(1, 7, 20, 16, 64, 128, 3)
        allok &= conv_case(2, 3, 64, 64, 16, 32, 3)
        allok &= conv_case(1, 40, 16, 16, 32, 32, 3)
        allok &= conv_case(2, 2, 16, 16, 16, 64, 3)
    if which == "one":  # a single launch of the dominant layer shape, for ncu
        allok &= conv_case(2, 128, 128, 128, 32, 32, 3)
    if which == "convT":
        allok &= convT_case(1, 4, 4, 4, 32, 16)
        allok &= convT_case(2, 8, 8, 8, 64, 32)
        allok &= convT_case(1, 8, 16, 16, 128, 64)
        allok &= convT_case(2, 64, 64, 64, 64, 32, iters=5)
        allok &= convT_case(2, 32, 32, 32, 128, 64, iters=5)
    if which == "pw":
        allok &= conv_case(2, 128, 128, 128, 16, 32, 1, iters=5)
    if which == "pw64":
        allok &= conv_case(2, 128, 128, 128, 64, 32, 1, iters=5)
    if which == "l2":
        allok &= conv_case(2, 32, 32, 32, 128, 128, 3, iters=5)
    if which == "l3":
        allok &= conv_case(2, 16, 16, 16, 256, 256, 3, iters=5)
    if which == "l4":
        allok &= conv_case(2, 8, 8, 8, 512, 512, 3, iters=5)
    if which == "one1x1":
        allok &= conv_case(2, 128, 128, 128, 64, 32, 1)
    if which == "one64":
        allok &= conv_case(2, 128, 128, 128, 64, 32, 3)
    if which in ("all", "perf"):
        allok &= conv_case(2, 128, 128, 128, 32, 32, 3, iters=5)
        allok &= conv_case(2, 128, 128, 128, 64, 32, 3, iters=5)
        allok &= conv_case(2, 128, 128, 128, 16, 32, 3, iters=5)
        allok &= conv_case(2, 128, 128, 128, 32, 16, 3, groups=16, iters=5)
        allok &= conv_case(2, 64, 64, 64, 64, 64, 3, iters=5)
        allok &= conv_case(2, 64, 64, 64, 128, 64, 3, iters=5)
        allok &= conv_case(2, 32, 32, 32, 128, 128, 3, iters=5)
        allok &= conv_case(2, 32, 32, 32, 256, 128, 3, iters=5)
        allok &= conv_case(2, 16, 16, 16, 256, 256, 3, iters=5)
        allok &= conv_case(2, 8, 8, 8, 512, 512, 3, iters=5)
        allok &= conv_case(2, 4, 4, 4, 1024, 1024, 3, iters=5)
        allok &= conv_case(2, 128, 128, 128, 64, 32, 1, iters=5)
    print("ALL OK" if allok else "SOME FAILED", "in %.1fs" % (time.time() - t0), flush=True)
    sys.exit(0 if allok else 1)
