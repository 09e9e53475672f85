"""Does a tensor-bound weight-gradient kernel really overlap with a bandwidth-bound kernel of the main chain?  Times, at the
level-0 shapes of cfg 3 (2 x 128^3, 32 channels): the 3x3x3 weight gradient (wg2_kernel) alone, a GroupNorm backward /
gate_apply_bwd / pool_bwd alone, and the pair enqueued on two streams, with the weight gradient first or second.

    python scripts/overlap_probe.py > gpurun_out/overlap_probe.txt"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import b3d  # noqa
from unet3d_b200 import ops

dev = torch.device("cuda:0")
torch.manual_seed(0)
N, S, C, G = 2, 128, 32, 8
bf = torch.bfloat16


def act(c=C):
    return torch.randn(N, S, S, S, c, device=dev).to(bf)


def stats_of(y, groups):
    n = y.shape[0]
    yf = y.float().view(n, -1, groups, y.shape[-1] // groups)
    return torch.stack([yf.sum(dim=(1, 3), dtype=torch.float64), (yf.double() ** 2).sum(dim=(1, 3))], dim=-1).contiguous()


x, dy, y, dout, r = act(), act(), act(), act(), act()
st, st_r = stats_of(y, G), stats_of(r, G)
gam, bet = torch.rand(C, device=dev) + 0.5, torch.randn(C, device=dev)
dw = torch.empty(C, C, 3, 3, 3, device=dev)
dx = torch.empty_like(y)
side = torch.cuda.Stream()
ops.WGRAD_SIDE = False   # the probe places the kernels on streams itself


def wgrad():
    ops._conv_wgrad(x, dy, C, C, 3, dw=dw)


def gn_bwd():
    ops.gn_bwd(dout, y, st, gam, bet, G, True, dx=dx)


def gn_dual():
    ops.gn_bwd_dual(dout, y, st, gam, bet, r, st_r, gam, G)


def gn_fwd():
    ops.gn_apply(y, st, gam, bet, G, True, out=dx)


def conv():
    ops.conv_fprop(x, WP, ROWS, C, 3, groups=8)


w = torch.randn(C, C, 3, 3, 3, device=dev) * 0.05
WP, _, ROWS = ops.pack_weight(w, ops.PACK_FPROP)


def timed(fn, it=20):
    for _ in range(3):
        ops.reset_scratch(); fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    tot = 0.0
    for _ in range(it):
        ops.reset_scratch()
        torch.cuda._sleep(int(0.002 * 1.9e9))   # ~2 ms spin: the host enqueues everything before the device gets there
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / it * 1e3


def pair(a, b, a_first):
    """a on the side stream, b on the current stream; both wait for the same point, the current stream joins the side stream."""
    def run():
        cur = torch.cuda.current_stream()
        side.wait_stream(cur)
        if a_first:
            with torch.cuda.stream(side):
                a()
            b()
        else:
            b()
            with torch.cuda.stream(side):
                a()
        cur.wait_stream(side)
    return run


ta = timed(wgrad)
print("wg2 32,32 weight gradient alone                 %7.1f us" % ta)
for name, fn in (("gn_bwd (reduce + apply)", gn_bwd), ("gn_bwd_dual (reduce + apply)", gn_dual), ("gn_apply forward", gn_fwd),
                 ("zs conv 32->32 (tensor-bound, cannot co-reside)", conv)):
    tb = timed(fn)
    t1 = timed(pair(wgrad, fn, True))
    t2 = timed(pair(wgrad, fn, False))
    print("%-48s alone %7.1f us | sum %7.1f | concurrent, wgrad enqueued first %7.1f, second %7.1f  (hidden: %5.1f / %5.1f us)"
          % (name, tb, ta + tb, t1, t2, ta + tb - t1, ta + tb - t2))
