"""In-process sweep of implicit-GEMM tile plans (b3d_set_plan_override) for the 3x3x3 convolutions of levels 2-5 (and their
dgrads = the same kernel with swapped channel counts) at cfg 3 (batch 2) and cfg 2 (batch 1).  Prints the planner's choice,
the best forced plan and the ratio."""
import ctypes, itertools, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import b3d  # noqa
from unet3d_b200 import ops, _lib
dev = torch.device("cuda:0")
L = _lib.lib()
L.b3d_last_plan.restype = ctypes.c_char_p
torch.manual_seed(0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
DEEP = len(sys.argv) > 1 and sys.argv[1] == "deep"


def timeit(fn, n=8):
    """ms per call; every call is timed alone with cold L2 (operands come from HBM, as inside a real step: the weights of one
    deep conv would otherwise stay L2-resident) and with the device kept behind the host (no host gap inside the event pair)."""
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    tot = 0.0
    pairs = []
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda._sleep(int(3e-4 * 1.9e9))
        flush.zero_()
        e0.record(); fn(); e1.record()
        pairs.append((e0, e1))
    torch.cuda.synchronize()
    return sum(a.elapsed_time(b) for a, b in pairs) / n


def run(n, s, cin, cout, ks=3, stats=True):
    x = torch.randn(n, s, s, s, cin, device=dev).to(torch.bfloat16)
    w = torch.randn(cout, cin, ks, ks, ks, device=dev) * 0.02
    wp, kp, rows = ops.pack_weight(w, ops.PACK_FPROP)
    groups = 8 if stats else 0
    fn = lambda: ops.conv_fprop(x, wp, rows, cout, ks, groups=groups)
    L.b3d_set_plan_override(0, 0, 0, 0, 0)
    base = timeit(fn)
    base_plan = L.b3d_last_plan().decode()
    gf = 2.0 * n * s ** 3 * cin * cout * ks ** 3 / 1e9
    results = []
    for kc, bn in itertools.product((16, 32, 64), (32, 64, 128, 256)):
        if bn > cout or cin % kc:
            continue
        L.b3d_set_plan_override(0, 0, 0, kc, bn)
        try:
            t = timeit(fn, 4)
            results.append((t, L.b3d_last_plan().decode(), (0, 0, 0, kc, bn)))
        except Exception:
            pass
    results.sort()
    # refine tile dims around the two best (KC, BN)
    for _, _, (_, _, _, kc, bn) in list(results[:2]):
        for td, th, tw in itertools.product((1, 2, 4, 8), (16, 32), (8, 16, 32)):
            L.b3d_set_plan_override(td, th, tw, kc, bn)
            try:
                t = timeit(fn, 4)
                results.append((t, L.b3d_last_plan().decode(), (td, th, tw, kc, bn)))
            except Exception:
                pass
    results.sort()
    L.b3d_set_plan_override(0, 0, 0, 0, 0)
    bt, bplan, bover = results[0]
    print("conv%d %dx%d^3 %4d->%4d: planner %.1f us (%4.0f TF/s) [%s] | best %.1f us (%4.0f TF/s) [%s] x%.2f" % (
        ks, n, s, cin, cout, base * 1e3, gf / base, base_plan, bt * 1e3, gf / bt, bplan, base / bt), flush=True)
    for t, plan, ov in results[1:4]:
        print("      next: %.1f us [%s]" % (t * 1e3, plan))


for n in (2, 1):
    for s, cin, cout in (((8, 512, 512), (8, 256, 512), (8, 1024, 512), (8, 512, 1024), (4, 512, 1024), (4, 1024, 1024), (4, 1024, 512)) if DEEP else ((32, 128, 128), (32, 256, 128), (32, 128, 256), (16, 256, 256), (16, 128, 256), (16, 512, 256), (16, 256, 512),
                         (8, 512, 512), (8, 256, 512), (8, 1024, 512), (8, 512, 1024), (4, 512, 1024), (4, 1024, 1024), (4, 1024, 512))):
        try:
            run(n, s, cin, cout)
        except Exception as e:
            print("FAILED", n, s, cin, cout, str(e)[:100])
