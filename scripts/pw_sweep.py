"""Sweep of implicit-GEMM plans (b3d_set_plan_override) for the level-0 / level-1 POINTWISE launches of cfg 3: the 1x1x1
convolutions (residual, attention-gate projections) and the ConvTranspose3d(k2,s2) forward / input gradient.  Prints, per
case, the planner's choice and time, the HBM floor (compulsory bytes at the measured copy bandwidth), and the best forced plan.
    python scripts/pw_sweep.py > gpurun_out/pw_sweep.txt"""
import ctypes, itertools, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import b3d  # noqa
from unet3d_b200 import ops, _lib
dev = torch.device("cuda:0")
L = _lib.lib()
L.b3d_last_plan.restype = ctypes.c_char_p
torch.manual_seed(0)
bf = torch.bfloat16
PEAK = 6544.7e9


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def sweep(name, fn, nbytes, kcs, bns):
    L.b3d_set_plan_override(0, 0, 0, 0, 0)
    base = timeit(fn)
    base_plan = L.b3d_last_plan().decode()
    res = []
    for kc, bn in itertools.product(kcs, bns):
        for td, th, tw in ((0, 0, 0), (1, 16, 32), (1, 32, 32), (2, 16, 32), (2, 32, 32), (4, 16, 32), (1, 16, 64), (1, 8, 128), (1, 16, 128), (1, 32, 16)):
            L.b3d_set_plan_override(td, th, tw, kc, bn)
            try:
                t = timeit(fn, 5)
                res.append((t, L.b3d_last_plan().decode()))
            except Exception:
                pass
    L.b3d_set_plan_override(0, 0, 0, 0, 0)
    res.sort()
    floor = nbytes / PEAK * 1e6
    print("%-34s planner %6.1f us [%s]\n%-34s floor   %6.1f us | best %6.1f us [%s]" % (name, base * 1e3, base_plan, "", floor, res[0][0] * 1e3 if res else -1, res[0][1] if res else ""), flush=True)
    for t, plan in res[1:3]:
        print("%-34s next    %6.1f us [%s]" % ("", t * 1e3, plan))


def pw(n, s, cin, cout, groups, add=False):
    x = torch.randn(n, s, s, s, cin, device=dev).to(bf)
    w = torch.randn(cout, cin, 1, 1, 1, device=dev) * 0.1
    wp, kp, rows = ops.pack_weight(w, ops.PACK_FPROP)
    out = ops.new_act(n, s, s, s, cout, dev)
    nb = n * s ** 3 * (cin + cout) * 2
    sweep("1x1 %dx%d^3 %d->%d g%d" % (n, s, cin, cout, groups), lambda: ops.conv_fprop(x, wp, rows, cout, 1, groups=groups, out=out), nb,
          [k for k in (16, 32, 64) if cin % k == 0], [b for b in (16, 32, 64) if b <= cout])


def convT(n, s, cin, cout):
    x = torch.randn(n, s, s, s, cin, device=dev).to(bf)
    wt = torch.randn(cin, cout, 2, 2, 2, device=dev) * 0.1
    packs = ops.pack_weight_pair(wt, True)
    wpf, _, _ = packs[ops.PACK_CONVT_FPROP]
    wpd, kd, rowsd = packs[ops.PACK_CONVT_DGRAD]
    bias = torch.zeros(cout, device=dev)
    cat = ops.new_act(n, 2 * s, 2 * s, 2 * s, 2 * cout, dev)
    up = cat[..., cout:]
    nb = n * s ** 3 * cin * 2 + n * (2 * s) ** 3 * cout * 2
    sweep("convT fprop %dx%d^3 %d->%d" % (n, s, cin, cout), lambda: ops.convT2_fprop(x, wpf, bias, cout, out=up), nb,
          [k for k in (16, 32, 64) if cin % k == 0], (32, 64, 128, 256))
    du = torch.randn(n, 2 * s, 2 * s, 2 * s, 2 * cout, device=dev).to(bf)[..., cout:]
    sweep("convT dgrad %dx%d^3 %d<-%d" % (n, s, cin, cout), lambda: ops.convT2_dgrad(du, wpd, rowsd, cin), nb,
          [k for k in (16, 32) if cout % k == 0], [b for b in (16, 32, 64) if b <= cin])


which = sys.argv[1] if len(sys.argv) > 1 else "all"
if which in ("all", "pw"):
    pw(2, 128, 16, 32, 8)      # enc0 residual (4 -> 32, input padded to 16)
    pw(2, 128, 32, 16, 4)      # attention-gate projections at level 0
    pw(2, 128, 64, 32, 8)      # dec0 residual
    pw(2, 64, 64, 32, 4)       # level-1 gate projections
    pw(2, 64, 32, 64, 8)       # enc1 residual
if which in ("all", "convT"):
    convT(2, 64, 64, 32)       # ups.12: 64^3 -> 128^3
    convT(2, 32, 128, 64)      # ups.9
