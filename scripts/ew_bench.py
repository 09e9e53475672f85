"""HBM efficiency of the level-0 bandwidth kernels, each timed alone with CUDA events on cfg-3 shapes (2 x 128^3 x 32 ch:
every tensor is 268 MB, larger than L2, so no flush is needed between iterations).

    python scripts/ew_bench.py [iters] > gpurun_out/ew_bench.txt

Prints, per op: us per call, the compulsory bytes (each input once + each output once) and the resulting GB/s against the
measured copy bandwidth in MEASURED_PEAKS.json."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import b3d  # noqa
from unet3d_b200 import ops

IT = int(sys.argv[1]) if len(sys.argv) > 1 else 20
ONLY = sys.argv[2] if len(sys.argv) > 2 else ""   # substring filter on the op name
dev = torch.device("cuda:0")
N, S, C, G = 2, 128, 32, 8
V = S * S * S
T = N * V * C * 2
try:
    PEAK = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    PEAK = 6544.7
torch.manual_seed(0)


def act(c=C, s=S):
    return (torch.randn(N, s, s, s, c, device=dev) * 1.0).to(torch.bfloat16)


def stats_of(y, groups):
    """[N][G][2] float64 (sum, sum of squares) like the conv epilogue writes them."""
    n = y.shape[0]
    yf = y.float().view(n, -1, groups, y.shape[-1] // groups)
    s = yf.sum(dim=(1, 3), dtype=torch.float64)
    q = (yf.double() ** 2).sum(dim=(1, 3))
    return torch.stack([s, q], dim=-1).contiguous()


def timeit(name, fn, nbytes):
    if ONLY not in name:
        return
    for _ in range(3):
        ops.reset_scratch(); fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    tot = 0.0
    for _ in range(IT):
        ops.reset_scratch()
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    us = tot / IT * 1e3
    gbs = nbytes / (us * 1e-6) / 1e9
    print("%-34s %8.1f us  %7.1f MB  %7.0f GB/s  %5.1f%% of copy peak" % (name, us, nbytes / 1e6, gbs, 100 * gbs / PEAK), flush=True)


ya, yb, dy = act(), act(), act()
st_a, st_b = stats_of(ya, G), stats_of(yb, G)
gam = torch.rand(C, device=dev) + 0.5
bet = torch.randn(C, device=dev) * 0.1
out = torch.empty_like(ya)
timeit("gn_apply relu", lambda: ops.gn_apply(ya, st_a, gam, bet, G, True, out=out), 2 * T)
timeit("gn_apply relu + GN(res)", lambda: ops.gn_apply(ya, st_a, gam, bet, G, True, res=yb, res_stats=st_b, res_gamma=gam,
                                                         res_beta=bet, res_groups=G, out=out), 3 * T)
dx = torch.empty_like(ya)
timeit("gn_bwd relu (reduce+apply)", lambda: ops.gn_bwd(dy, ya, st_a, gam, bet, G, True, dx=dx), 5 * T)
timeit("gn_bwd_dual (reduce+apply)", lambda: ops.gn_bwd_dual(dy, ya, st_a, gam, bet, yb, st_b, gam, G), 8 * T)
w = torch.randn(4, C, device=dev) * 0.1
dl = torch.randn(N, S, S, S, 4, device=dev)
timeit("ds_head_bwd_cl accumulate", lambda: ops.ds_head_bwd_cl(dl, ya, w, dx, True), 3 * T + dl.numel() * 4)
timeit("ds_head_bwd_cl overwrite", lambda: ops.ds_head_bwd_cl(dl, ya, w, dx, False), 2 * T + dl.numel() * 4)
psi = torch.randn(N, V, device=dev)
st_psi = torch.stack([psi.sum(1, dtype=torch.float64), (psi.double() ** 2).sum(1)], dim=-1).contiguous()
gpsi = torch.ones(1, device=dev); bpsi = torch.zeros(1, device=dev)
ca = torch.rand(N, C, device=dev)
timeit("gate_apply_fwd", lambda: ops.gate_apply_fwd(ya, psi, st_psi, gpsi, bpsi, ca, out), 2 * T + psi.numel() * 4)
timeit("gate_apply_bwd", lambda: ops.gate_apply_bwd(dy, ya, psi, st_psi, gpsi, bpsi, ca, dx), 3 * T + 2 * psi.numel() * 4)
dyp = act(C, S // 2)
timeit("pool_bwd accumulate", lambda: ops.pool_bwd(ya, None, dyp, dx=dx, accumulate=True), 3 * T + T // 8)
timeit("pool_bwd overwrite", lambda: ops.pool_bwd(ya, None, dyp), 2 * T + T // 8)
timeit("add_bf16", lambda: ops.add_bf16(ya, yb, out=out), 3 * T)
