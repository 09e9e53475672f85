"""Thin Python wrappers over the C ABI of libb3d.so (one function per exported entry point).

Conventions
-----------
* activations: torch.bfloat16 tensors of shape [N, D, H, W, C] (NDHWC).  A tensor may be a channel slice of a wider
  buffer (e.g. half of a concat buffer): only stride(-1) == 1 and "voxel pitch" stride(3) (= ld, in elements) matter.
* logits / loss tensors: torch.float32 [N, K, D, H, W] contiguous (the reference's NCDHW), targets int64 [N, D, H, W].
* statistics: torch.float64 [N or 1, G, 2] = (sum, sum of squares) accumulated by the producing kernel.
* everything is enqueued on the current CUDA stream; nothing here synchronises with the host.
"""
import ctypes

import torch

from . import _lib
from ._lib import c_int, c_ll, c_sz, c_vp, check, ptr, stream_ptr

c_float, c_double = ctypes.c_float, ctypes.c_double
EPS = 1e-5


def _os_environ_flag(name, default):
    import os
    v = os.environ.get(name)
    return default if v is None else v != "0"
PROFILE = None  # bench.py sets this to a list: (family, algorithmic flops, start event, end event) per tensor-core call


PROFILE_BW = None  # bench.py: list of (family, compulsory HBM bytes, start event, end event, tag) per bandwidth-bound call


class _prof_bw:
    """CUDA events around a bandwidth-bound call together with its COMPULSORY traffic (every input once + every output once,
    SURVEY §8d) — bench.py turns them into achieved GB/s against the measured HBM peak."""

    def __init__(self, family, nbytes, tag=""):
        self.family, self.nbytes, self.tag = family, nbytes, tag

    def __enter__(self):
        if PROFILE_BW is not None:
            self.a = torch.cuda.Event(enable_timing=True)
            self.b = torch.cuda.Event(enable_timing=True)
            self.a.record()
        return self

    def __exit__(self, *exc):
        if PROFILE_BW is not None:
            self.b.record()
            PROFILE_BW.append((self.family, self.nbytes, self.a, self.b, self.tag))
        return False


class _prof:
    def __init__(self, family, flops, tag=""):
        self.family, self.flops, self.tag = family, flops, tag

    def __enter__(self):
        if PROFILE is not None:
            self.a = torch.cuda.Event(enable_timing=True)
            self.b = torch.cuda.Event(enable_timing=True)
            self.a.record()
        return self

    def __exit__(self, *exc):
        if PROFILE is not None:
            self.b.record()
            PROFILE.append((self.family, self.flops, self.a, self.b, self.tag))
        return False


def _L():
    return _lib.lib()


def roundup(a, b):
    return (a + b - 1) // b * b


def ld(t):
    """voxel pitch (elements) of an NDHWC activation (possibly a channel slice)."""
    assert t.dim() == 5 and t.stride(4) == 1, "expected NDHWC activation"
    n, d, h, w, _ = t.shape
    p = t.stride(3)
    assert t.stride(2) == w * p and t.stride(1) == h * w * p and t.stride(0) == d * h * w * p, "non-dense voxel layout"
    return p


def new_act(n, d, h, w, c, device, zero=False):
    f = torch.zeros if zero else torch.empty
    return f((n, d, h, w, c), dtype=torch.bfloat16, device=device)


# ---------------------------------------------------------------------------------------------------------------
# zeroed scratch: the ~130 small statistics / reduction buffers of a step (fp64 sums the kernels accumulate into with
# atomics) are carved out of a few zero-filled arenas by a bump allocator instead of one fill kernel each.  A slice is handed
# out once and never reused, so it is zero when its kernel runs; an arena is freed when its last slice dies.  Arenas are
# per stream (the fill is ordered on the stream that was current when the arena was created).
# ---------------------------------------------------------------------------------------------------------------
_ARENA_BYTES = 1 << 18
_arenas = {}
_arena_lock = __import__("threading").Lock()   # the reference serves from threaded Flask workers (main.py:1059)


def zeros_scratch(shape, dtype, device):
    n = 1
    for s_ in shape:
        n *= int(s_)
    nbytes = roundup(max(n, 1) * torch.empty((), dtype=dtype).element_size(), 256)
    if device.type != "cuda" or nbytes > _ARENA_BYTES // 4:
        return torch.zeros(shape, dtype=dtype, device=device)
    # an arena filled during a graph capture belongs to that graph (and one filled eagerly is not re-zeroed by a replay)
    key = (device.index, torch.cuda.current_stream(device).cuda_stream, torch.cuda.is_current_stream_capturing())
    with _arena_lock:   # bump allocation must be atomic across host threads
        a = _arenas.get(key)
        if a is None or a[1] + nbytes > _ARENA_BYTES:
            a = [torch.zeros(_ARENA_BYTES, dtype=torch.uint8, device=device), 0]
            _arenas[key] = a
        off = a[1]
        a[1] = off + nbytes
    return a[0][off:off + n * torch.empty((), dtype=dtype).element_size()].view(dtype).view(shape)


def reset_scratch():
    """Forget the current arenas (a CUDA-graph capture must not slice an arena that was filled outside the capture)."""
    _arenas.clear()


# ---------------------------------------------------------------------------------------------------------------
# weights
# ---------------------------------------------------------------------------------------------------------------
PACK_FPROP, PACK_DGRAD, PACK_CONVT_FPROP, PACK_CONVT_DGRAD = 0, 1, 2, 3
PAIR_PACK = True   # functional.packed: produce the fprop and dgrad copies of a weight with one launch


def pack_weight(w, mode):
    """fp32 reference-layout weight -> bf16 [K/8][taps][rows][8] (see conv_igemm.cu).  Returns (packed, Kp, rows)."""
    w = w.detach()
    if not w.is_contiguous():
        w = w.contiguous()
    if mode in (PACK_FPROP, PACK_DGRAD):
        cout, cin = w.shape[0], w.shape[1]
        ntaps = w.shape[2] * w.shape[3] * w.shape[4]
        if mode == PACK_FPROP:
            kp, rows = roundup(cin, 16), roundup(cout, 16)
        else:
            kp, rows = roundup(cout, 16), roundup(cin, 16)
        ptaps = ntaps
    else:
        cin, cout = w.shape[0], w.shape[1]
        ntaps = 8
        if mode == PACK_CONVT_FPROP:
            kp, rows = roundup(cin, 16), 8 * cout
        else:
            kp, rows = 8 * cout, roundup(cin, 16)
        ptaps = 1
    out = torch.empty((kp // 8) * ptaps * rows * 8, dtype=torch.bfloat16, device=w.device)
    check(_L().b3d_pack_weight(c_int(mode), ptr(w), c_int(cout), c_int(cin), c_int(ntaps), ptr(out), c_int(kp), c_int(rows),
                               stream_ptr()))
    return out, kp, rows


def pack_weight_pair(w, conv_transpose):
    """Both packed copies of one weight in a single launch (one read of the fp32 weight).  Returns
    {mode: (packed, Kp, rows)} for modes (PACK_FPROP, PACK_DGRAD) or (PACK_CONVT_FPROP, PACK_CONVT_DGRAD)."""
    w = w.detach()
    if not w.is_contiguous():
        w = w.contiguous()
    dev = w.device
    if not conv_transpose:
        cout, cin = w.shape[0], w.shape[1]
        ntaps = w.shape[2] * w.shape[3] * w.shape[4]
        ri, ro = roundup(cin, 16), roundup(cout, 16)
        of = torch.empty(ntaps * ro * ri, dtype=torch.bfloat16, device=dev)
        od = torch.empty(ntaps * ri * ro, dtype=torch.bfloat16, device=dev)
        check(_L().b3d_pack_weight_pair(c_int(0), ptr(w), c_int(cout), c_int(cin), c_int(ntaps), ptr(of), ptr(od), stream_ptr()))
        return {PACK_FPROP: (of, ri, ro), PACK_DGRAD: (od, ro, ri)}
    cin, cout = w.shape[0], w.shape[1]
    ri = roundup(cin, 16)
    of = torch.empty(8 * cout * ri, dtype=torch.bfloat16, device=dev)
    od = torch.empty(ri * 8 * cout, dtype=torch.bfloat16, device=dev)
    check(_L().b3d_pack_weight_pair(c_int(1), ptr(w), c_int(cout), c_int(cin), c_int(8), ptr(of), ptr(od), stream_ptr()))
    return {PACK_CONVT_FPROP: (of, ri, 8 * cout), PACK_CONVT_DGRAD: (od, 8 * cout, ri)}


# ---------------------------------------------------------------------------------------------------------------
# tcgen05 convolutions
# ---------------------------------------------------------------------------------------------------------------
def conv_fprop(x, wpack, w_rows, cout, ks, bias=None, groups=0, stats_batch=False, out=None, cin=None, add=None,
               cin_real=None):
    """y = conv(x) (+bias) (+add); optional GroupNorm (per sample) or BatchNorm (stats_batch) partial sums of y.
    `cin`: number of leading channels of x to contract over (multiple of 16; defaults to all).
    `add`: NDHWC bf16 tensor summed into the result in the epilogue (1x1x1 only; may be `out` itself).
    `cin_real`: channels of x that are not zero padding — only used to count ALGORITHMIC FLOPs for bench.py's roofline."""
    n, d, h, w, c = x.shape
    cin = c if cin is None else cin
    dev = x.device
    if out is None:
        out = new_act(n, d, h, w, cout, dev)
    if (add is not None and ADD_ON_MMA and ks == 1 and bias is None and not groups and cin % 16 == 0 and cout % 16 == 0
            and w_rows == cout and _gcd_maps(cin, cout) <= 8):
        return _conv1_add_mma(x, wpack, w_rows, cin, cout, add, out), None
    stats = None
    if groups:
        stats = zeros_scratch((1 if stats_batch else n, groups, 2), torch.float64, dev)
    # split-K workspace (fp32 partial slices [split][V][Cout]): the library states how much it wants for this shape
    ws_bytes = int(_L().b3d_conv_fprop_workspace_bytes(c_int(n), c_int(d), c_int(h), c_int(w), c_int(cout))) if add is None else 0
    ws = torch.empty(ws_bytes // 4, dtype=torch.float32, device=dev) if ws_bytes else None
    with _prof("igemm", 2.0 * n * d * h * w * (cin if cin_real is None else cin_real) * cout * ks ** 3, "conv%d %dx%dx%dx%d %d->%d" % (ks, n, d, h, w, cin, cout)):
        check(_L().b3d_conv_fprop_add(ptr(x), c_ll(ld(x)), ptr(wpack), c_int(w_rows), ptr(bias), ptr(add),
                                  c_ll(ld(add) if add is not None else 0), ptr(out), c_ll(ld(out)),
                                  c_int(n), c_int(d), c_int(h), c_int(w), c_int(cin), c_int(cout), c_int(ks), ptr(stats),
                                  c_int(groups), c_int(1 if stats_batch else 0), ptr(ws),
                                  c_sz(ws_bytes if ws is not None else 0), ptr(_lib.err_flag(dev)), stream_ptr()))
    return out, stats


ADD_ON_MMA = _os_environ_flag("B3D_ADD_ON_MMA", True)
_EYE = {}


def _gcd_maps(cin, cout):
    import math
    g = math.gcd(cin, cout)
    return (cin + cout) // g if g % 16 == 0 else 99


def _conv1_add_mma(x, wpack, w_rows, cin, cout, add, out):
    """out = conv1x1(x) + add with the addend riding the tensor core's K dimension: B = [W ; I] (see b3d_conv1_add_mma)."""
    n, d, h, w, _ = x.shape
    dev = x.device
    eye = _EYE.get((cout, dev))
    if eye is None:
        eye = _EYE[(cout, dev)] = torch.eye(cout, dtype=torch.bfloat16, device=dev)
    aug = torch.cat([wpack.view(w_rows, cin), eye], dim=1)          # [Cout][Cin + Cout]: a few KB .. 1.5 MB per layer
    nbytes = n * d * h * w * (cin + 2 * cout) * 2
    with _prof("igemm", 2.0 * n * d * h * w * cin * cout, "conv1 %dx%dx%dx%d %d->%d +add" % (n, d, h, w, cin, cout)), \
            _prof_bw("pointwise_add", nbytes, "conv1+add"):
        check(_L().b3d_conv1_add_mma(ptr(x), c_ll(ld(x)), ptr(add), c_ll(ld(add)), ptr(aug), c_int(w_rows), ptr(out), c_ll(ld(out)),
                                     c_int(n), c_int(d), c_int(h), c_int(w), c_int(cin), c_int(cout), ptr(_lib.err_flag(dev)),
                                     stream_ptr()))
    return out


def convT2_fprop(x, wpack, bias, cout, out=None):
    n, d, h, w, cin = x.shape
    if out is None:
        out = new_act(n, 2 * d, 2 * h, 2 * w, cout, x.device)
    with _prof("igemm", 2.0 * n * d * h * w * cin * cout * 8, "convT %dx%dx%dx%d %d->%d" % (n, d, h, w, cin, cout)):
        check(_L().b3d_convT2_fprop(ptr(x), c_ll(ld(x)), ptr(wpack), ptr(bias), ptr(out), c_ll(ld(out)), c_int(n), c_int(d),
                                    c_int(h), c_int(w), c_int(cin), c_int(cout), ptr(_lib.err_flag(x.device)), stream_ptr()))
    return out


def convT2_dgrad(dy, wpack, w_rows, cin, out=None):
    n, d2, h2, w2, cout = dy.shape
    d, h, w = d2 // 2, h2 // 2, w2 // 2
    dev = dy.device
    if out is None:
        out = new_act(n, d, h, w, cin, dev)
    ws = torch.empty(int(_L().b3d_convT2_dgrad_workspace_bytes(c_int(n), c_int(d), c_int(h), c_int(w), c_int(cin))) // 4,
                     dtype=torch.float32, device=dev)
    with _prof("igemm", 2.0 * n * d * h * w * cin * cout * 8, "convT_dgrad %dx%dx%dx%d %d<-%d" % (n, d, h, w, cin, cout)):
        check(_L().b3d_convT2_dgrad(ptr(dy), c_ll(ld(dy)), ptr(wpack), c_int(w_rows), ptr(out), c_ll(ld(out)), c_int(n),
                                    c_int(d), c_int(h), c_int(w), c_int(cin), c_int(cout), ptr(ws), c_sz(ws.numel() * 4),
                                    ptr(_lib.err_flag(dev)), stream_ptr()))
    return out


# Weight gradients are leaves of the backward graph: nothing downstream waits for them until the optimizer / all-reduce.
# They are enqueued on a side stream (disable with B3D_WGRAD_STREAM=0) so that the bandwidth-bound kernels of the main
# chain (GroupNorm backward, gate, heads) overlap with these tensor-bound kernels: 21.4 -> 19.8 ms per cfg-3 step.
# modules.py joins the stream before the gradients leave backward; parallel.py makes the all-reduce buckets wait for it.
import os as _os

WGRAD_STREAM = None
WGRAD_SIDE = _os.environ.get("B3D_WGRAD_STREAM", "1") != "0"


class _wgrad_ctx:
    def __init__(self, *inputs):
        self.inputs = inputs

    def __enter__(self):
        global WGRAD_STREAM
        if WGRAD_SIDE and WGRAD_STREAM is None:
            WGRAD_STREAM = torch.cuda.Stream()
        self.s = WGRAD_STREAM
        if self.s is not None:
            self.s.wait_stream(torch.cuda.current_stream())   # inputs (and the previous optimizer step) are ordered before
            self.ctx = torch.cuda.stream(self.s)
            self.ctx.__enter__()
        return self

    def __exit__(self, *exc):
        if self.s is not None:
            self.ctx.__exit__(*exc)
            for t in self.inputs:
                t.record_stream(self.s)
        return False


# Independent HBM-bound branches (the residual 1x1x1 conv + its GroupNorm backward, the deep-supervision heads) can ride the
# same side stream so that they overlap with the tensor-bound 3x3x3 kernels of the main chain.  B3D_BRANCH_STREAM selects
# which: bit 0 = forward residual conv, bit 1 = forward deep-supervision heads, bit 2 = backward residual GroupNorm.
BRANCH_MASK = int(_os.environ.get("B3D_BRANCH_STREAM", "0"))


class side_branch:
    """with side_branch(enabled, inputs...) as br: <enqueue the branch> ; later br.join() before the main stream reads its
    results.  A no-op context when disabled (or when the side stream is switched off)."""

    def __init__(self, enabled, *inputs):
        self.on = bool(enabled) and WGRAD_SIDE
        self.inputs = inputs
        self.event = None

    def __enter__(self):
        global WGRAD_STREAM
        if self.on:
            if WGRAD_STREAM is None:
                WGRAD_STREAM = torch.cuda.Stream()
            WGRAD_STREAM.wait_stream(torch.cuda.current_stream())
            self.ctx = torch.cuda.stream(WGRAD_STREAM)
            self.ctx.__enter__()
        return self

    def __exit__(self, *exc):
        if self.on:
            self.event = WGRAD_STREAM.record_event()
            self.ctx.__exit__(*exc)
            for t in self.inputs:
                if t is not None:
                    t.record_stream(WGRAD_STREAM)
        return False

    def join(self):
        if self.event is not None:
            torch.cuda.current_stream().wait_event(self.event)


def wgrad_join():
    """Make the current stream wait for every weight gradient enqueued on the side stream."""
    if WGRAD_STREAM is not None:
        torch.cuda.current_stream().wait_stream(WGRAD_STREAM)


def _pad_w16(x, dy, up):
    """The weight-gradient kernels run K along rows of W positions in steps of 16 (whole small planes when W <= 16).  Rows
    whose width is not a multiple of 16 (the 160x192x160 volumes of BASELINE config 5: W = 40, 20 at levels 2-3) are
    zero-padded along W: padded dY columns contribute nothing and padded X columns equal the convolution's own zero
    padding, so the sum over voxels is unchanged.  `up`: dy is `up` times finer than x (2 for ConvTranspose)."""
    w = x.shape[3]
    if w % 16 == 0 or w <= 16:
        return x, dy
    pw = roundup(w, 16) - w
    xp = torch.nn.functional.pad(x, (0, 0, 0, pw))
    dyp = torch.nn.functional.pad(dy, (0, 0, 0, up * pw))
    return xp, dyp


def conv_wgrad(x, dy, cin_real, cout, ks, dw=None, accumulate=False):
    with _wgrad_ctx(x, dy):
        return _conv_wgrad(x, dy, cin_real, cout, ks, dw, accumulate)


def _conv_wgrad(x, dy, cin_real, cout, ks, dw=None, accumulate=False):
    """fp32 weight gradient in the reference layout [Cout, Cin, k, k, k].  x may carry zero-padded channels beyond
    cin_real (its whole channel extent is contracted, only the first cin_real rows are written)."""
    if ks == 1 and (x.shape[0] * x.shape[1] * x.shape[2] * x.shape[3]) % 16 != 0 and x.shape[0] * x.shape[1] * x.shape[2] * x.shape[3] > 16:
        # pointwise: every voxel is just a K row; flatten and zero-pad the voxel count to a multiple of 16
        v = x.shape[0] * x.shape[1] * x.shape[2] * x.shape[3]
        pv = roundup(v, 16) - v
        x = torch.nn.functional.pad(x.reshape(1, 1, 1, v, x.shape[-1]), (0, 0, 0, pv))
        dy = torch.nn.functional.pad(dy.reshape(1, 1, 1, v, dy.shape[-1]), (0, 0, 0, pv))
    else:
        x, dy = _pad_w16(x, dy, 1)
    n, d, h, w, cin = x.shape
    dev = x.device
    if dw is None:
        dw = torch.empty((cout, cin_real, ks, ks, ks), dtype=torch.float32, device=dev)
        accumulate = False
    cout_pad = roundup(cout, 16)
    assert dy.shape[-1] == cout and (cout_pad == cout or ld(dy) >= cout_pad), "dy must expose padded channels"
    ws = torch.empty(int(_L().b3d_conv_wgrad_workspace_bytes(c_int(cin), c_int(cout), c_int(ks))) // 4, dtype=torch.float32, device=dev)
    with _prof("wgrad", 2.0 * n * d * h * w * cin_real * cout * ks ** 3, "wgrad%d %dx%dx%dx%d %d,%d" % (ks, n, d, h, w, cin, cout)), \
            _prof_bw("wgrad_pw" if ks == 1 else "wgrad3_bytes", n * d * h * w * (cin + cout) * 2, "wgrad%d" % ks):
        check(_L().b3d_conv_wgrad(ptr(x), c_ll(ld(x)), ptr(dy), c_ll(ld(dy)), ptr(dw), c_int(1 if accumulate else 0), c_int(n),
                                  c_int(d), c_int(h), c_int(w), c_int(cin), c_int(cin_real), c_int(cout), c_int(ks), ptr(ws),
                                  c_sz(ws.numel() * 4), ptr(_lib.err_flag(dev)), stream_ptr()))
    return dw


def convT2_wgrad(x, dy, cin, cout, dw=None, accumulate=False):
    with _wgrad_ctx(x, dy):
        return _convT2_wgrad(x, dy, cin, cout, dw, accumulate)


def _convT2_wgrad(x, dy, cin, cout, dw=None, accumulate=False):
    """fp32 ConvTranspose3d(k2,s2) weight gradient [Cin, Cout, 2, 2, 2]; x coarse [N,D,H,W,Cin], dy fine [N,2D,2H,2W,Cout]."""
    x, dy = _pad_w16(x, dy, 2)
    n, d, h, w, _ = x.shape
    dev = x.device
    if dw is None:
        dw = torch.empty((cin, cout, 2, 2, 2), dtype=torch.float32, device=dev)
        accumulate = False
    ws = torch.empty(int(_L().b3d_convT2_wgrad_workspace_bytes(c_int(cin), c_int(cout))) // 4, dtype=torch.float32, device=dev)
    with _prof("wgrad", 2.0 * n * d * h * w * cin * cout * 8, "wgradT %dx%dx%dx%d %d,%d" % (n, d, h, w, cin, cout)):
        check(_L().b3d_convT2_wgrad(ptr(x), c_ll(ld(x)), ptr(dy), c_ll(ld(dy)), ptr(dw), c_int(1 if accumulate else 0), c_int(n),
                                    c_int(d), c_int(h), c_int(w), c_int(cin), c_int(cout), ptr(ws), c_sz(ws.numel() * 4),
                                    ptr(_lib.err_flag(dev)), stream_ptr()))
    return dw


# ---------------------------------------------------------------------------------------------------------------
# GroupNorm family
# ---------------------------------------------------------------------------------------------------------------
def _nvc(t):
    n, d, h, w, c = t.shape
    return n, d * h * w, c


def gn_apply(y, stats, gamma, beta, groups, relu, res=None, res_stats=None, res_gamma=None, res_beta=None, res_groups=0,
             out=None):
    """out = act(GN(y)) [+ GN'(res) if res_stats is given, + res otherwise]."""
    n, v, c = _nvc(y)
    if out is None:
        out = torch.empty_like(y, memory_format=torch.contiguous_format)
    res_mode = 0 if res is None else (1 if res_stats is not None else 2)
    tb = n * v * c * 2
    with _prof_bw("gn_fwd", tb * (2 if res is None else 3), "gn_apply C%d V%d" % (c, v)):
        check(_L().b3d_gn_apply(ptr(y), c_ll(ld(y)), ptr(stats), ptr(gamma), ptr(beta), c_int(groups), c_int(1 if relu else 0),
                                c_int(res_mode), ptr(res), c_ll(ld(res) if res is not None else 0), ptr(res_stats),
                                ptr(res_gamma), ptr(res_beta), c_int(res_groups if res_groups else 1), ptr(out), c_ll(ld(out)),
                                c_int(n), c_ll(v), c_int(c), c_float(EPS), stream_ptr()))
    return out


def gn_bwd(dy, y, stats, gamma, beta, groups, relu, dx=None, accumulate=False, sums=None, want_param_grads=True):
    """GroupNorm(+ReLU) backward.  Returns (dx, dgamma, dbeta).  If `sums` ([N][C][2] float64: Σdz, Σdz·x̂) is given the
    reduction phase is skipped (a producer already computed it)."""
    n, v, c = _nvc(y)
    dev = y.device
    with _prof_bw("gn_bwd", n * v * c * 2 * (5 if sums is None else 3), "gn_bwd C%d V%d" % (c, v)):
        if sums is None:
            sums = zeros_scratch((n, c, 2), torch.float64, dev)
            check(_L().b3d_gn_bwd_reduce(ptr(dy), c_ll(ld(dy)), ptr(y), c_ll(ld(y)), ptr(stats), ptr(gamma), ptr(beta),
                                         c_int(groups), c_int(1 if relu else 0), ptr(sums), c_int(n), c_ll(v), c_int(c),
                                         c_float(EPS), stream_ptr()))
        if dx is None:
            dx = torch.empty_like(y, memory_format=torch.contiguous_format)
            accumulate = False
        check(_L().b3d_gn_bwd_apply(ptr(dy), c_ll(ld(dy)), ptr(y), c_ll(ld(y)), ptr(stats), ptr(gamma), ptr(beta), c_int(groups),
                                    c_int(1 if relu else 0), ptr(sums), ptr(dx), c_ll(ld(dx)), c_int(1 if accumulate else 0),
                                    c_int(n), c_ll(v), c_int(c), c_float(EPS), stream_ptr()))
    dgamma = dbeta = None
    if want_param_grads:
        dgamma = torch.empty(c, dtype=torch.float32, device=dev)
        dbeta = torch.empty(c, dtype=torch.float32, device=dev)
        check(_L().b3d_gn_param_grad(ptr(sums), c_int(n), c_int(c), ptr(dgamma), ptr(dbeta), c_int(0), stream_ptr()))
    return dx, dgamma, dbeta


DUAL_GN_BWD = _os.environ.get("B3D_DUAL_GN_BWD", "1") != "0"


def gn_bwd_dual(dy, ya, stats_a, gamma_a, beta_a, yb, stats_b, gamma_b, groups):
    """Backward of  out = relu(GN_a(ya)) + GN_b(yb)  w.r.t. both branches with ONE pass over dy per phase (8 instead of 10
    tensor passes).  Returns (dxa, dga, dba, dxb, dgb, dbb), or None when the shape does not suit the dual kernels."""
    if not DUAL_GN_BWD:
        return None
    n, v, c = _nvc(ya)
    dev = ya.device
    if c > 512 or (256 % max(c // 8, 1)) != 0 or ya.shape != yb.shape or dy.shape != ya.shape:
        return None
    sums_a = zeros_scratch((n, c, 2), torch.float64, dev)
    sums_b = zeros_scratch((n, c, 2), torch.float64, dev)
    dxa = torch.empty_like(ya, memory_format=torch.contiguous_format)
    dxb = torch.empty_like(yb, memory_format=torch.contiguous_format)
    with _prof_bw("gn_bwd", n * v * c * 2 * 8, "gn_bwd_dual C%d V%d" % (c, v)):
        rc = _dual_call(dy, ya, stats_a, gamma_a, beta_a, yb, stats_b, gamma_b, groups, sums_a, sums_b, dxa, dxb, n, v, c)
    if rc == 1:
        return None
    check(rc)
    g = [torch.empty(c, dtype=torch.float32, device=dev) for _ in range(4)]
    check(_L().b3d_gn_param_grad(ptr(sums_a), c_int(n), c_int(c), ptr(g[0]), ptr(g[1]), c_int(0), stream_ptr()))
    check(_L().b3d_gn_param_grad(ptr(sums_b), c_int(n), c_int(c), ptr(g[2]), ptr(g[3]), c_int(0), stream_ptr()))
    return dxa, g[0], g[1], dxb, g[2], g[3]


def _dual_call(dy, ya, stats_a, gamma_a, beta_a, yb, stats_b, gamma_b, groups, sums_a, sums_b, dxa, dxb, n, v, c):
    return _L().b3d_gn_bwd_dual(ptr(dy), c_ll(ld(dy)), ptr(ya), c_ll(ld(ya)), ptr(stats_a), ptr(gamma_a), ptr(beta_a), ptr(yb),
                              c_ll(ld(yb)), ptr(stats_b), ptr(gamma_b), c_int(groups), ptr(sums_a), ptr(sums_b), ptr(dxa),
                              c_ll(ld(dxa)), ptr(dxb), c_ll(ld(dxb)), c_int(n), c_ll(v), c_int(c), c_float(EPS), stream_ptr())


def add_bf16(a, b, out=None):
    n, v, c = _nvc(a)
    if out is None:
        out = torch.empty_like(a, memory_format=torch.contiguous_format)
    check(_L().b3d_add_bf16(ptr(a), c_ll(ld(a)), ptr(b), c_ll(ld(b)), ptr(out), c_ll(ld(out)), c_ll(n * v), c_int(c),
                            stream_ptr()))
    return out


# ---------------------------------------------------------------------------------------------------------------
# pooling, layout, reductions
# ---------------------------------------------------------------------------------------------------------------
def pool_fwd(x, mask=None):
    n, d, h, w, c = x.shape
    out = new_act(n, d // 2, h // 2, w // 2, c, x.device)
    check(_L().b3d_pool_fwd(ptr(x), c_ll(ld(x)), ptr(mask), ptr(out), c_ll(ld(out)), c_int(n), c_int(d), c_int(h), c_int(w),
                            c_int(c), stream_ptr()))
    return out


def relu_pool_fwd(x):
    """MaxPool3d(2)(ReLU(x)) — BrainTumorClassifier.features (main.py:307-312)."""
    n, d, h, w, c = x.shape
    out = new_act(n, d // 2, h // 2, w // 2, c, x.device)
    check(_L().b3d_relu_pool_fwd(ptr(x), c_ll(ld(x)), ptr(out), c_ll(ld(out)), c_int(n), c_int(d), c_int(h), c_int(w), c_int(c),
                                 stream_ptr()))
    return out


def relu_adaptive_avgpool(x, osize, relu=True):
    """AdaptiveAvgPool3d(osize)(ReLU(x)) -> fp32 [N, C*od*oh*ow] in the reference's NCDHW flatten order (main.py:315,326)."""
    n, d, h, w, c = x.shape
    od, oh, ow = osize
    out = torch.empty((n, c * od * oh * ow), dtype=torch.float32, device=x.device)
    check(_L().b3d_relu_adaptive_avgpool(ptr(x), c_ll(ld(x)), ptr(out), c_int(n), c_int(d), c_int(h), c_int(w), c_int(c),
                                         c_int(od), c_int(oh), c_int(ow), c_int(1 if relu else 0), stream_ptr()))
    return out


def pool_bwd(x, mask, dy, dx=None, accumulate=False, cadd=None):
    """`cadd`: optional fp32 [N,C] constant added to dx in the same pass (accumulate mode; see b3d_pool_bwd_add)."""
    n, d, h, w, c = x.shape
    if dx is None:
        dx = new_act(n, d, h, w, c, x.device)
        accumulate = False
        assert cadd is None
    check(_L().b3d_pool_bwd_add(ptr(x), c_ll(ld(x)), ptr(mask), ptr(dy), c_ll(ld(dy)), ptr(dx), c_ll(ld(dx)),
                                c_int(1 if accumulate else 0), ptr(cadd), c_int(n), c_int(d), c_int(h), c_int(w), c_int(c),
                                stream_ptr()))
    return dx


def to_ndhwc_bf16(x, cpad):
    """fp32 NCDHW -> bf16 NDHWC with channels zero-padded to cpad."""
    n, cin, d, h, w = x.shape
    x = x.contiguous()
    if x.dtype != torch.float32:
        x = x.float()
    out = new_act(n, d, h, w, cpad, x.device)
    check(_L().b3d_to_ndhwc_bf16(ptr(x), ptr(out), c_ll(cpad), c_int(n), c_int(cin), c_ll(d * h * w), c_int(cpad), stream_ptr()))
    return out


def to_ncdhw_f32(x):
    n, d, h, w, c = x.shape
    out = torch.empty((n, c, d, h, w), dtype=torch.float32, device=x.device)
    check(_L().b3d_to_ncdhw_f32(ptr(x), c_ll(ld(x)), ptr(out), c_int(n), c_int(c), c_ll(d * h * w), stream_ptr()))
    return out


def channel_sum(x):
    """float64 [N][C] = Σ over voxels."""
    n, v, c = _nvc(x)
    sums = zeros_scratch((n, c), torch.float64, x.device)
    check(_L().b3d_channel_sum(ptr(x), c_ll(ld(x)), ptr(sums), c_int(n), c_ll(v), c_int(c), stream_ptr()))
    return sums


# ---------------------------------------------------------------------------------------------------------------
# attention gate pieces
# ---------------------------------------------------------------------------------------------------------------
def gate_psi_fwd(g1r, x1r, st_g, st_x, gam_g, bet_g, gam_x, bet_x, wpsi, bpsi):
    n, v, f = _nvc(g1r)
    dev = g1r.device
    psi_raw = torch.empty((n, v), dtype=torch.float32, device=dev)
    st_psi = zeros_scratch((n, 2), torch.float64, dev)
    check(_L().b3d_gate_psi_fwd(ptr(g1r), ptr(x1r), ptr(st_g), ptr(st_x), ptr(gam_g), ptr(bet_g), ptr(gam_x), ptr(bet_x),
                                ptr(wpsi), ptr(bpsi), ptr(psi_raw), ptr(st_psi), c_int(n), c_ll(v), c_int(f), c_float(EPS),
                                stream_ptr()))
    return psi_raw, st_psi


def gate_se_fwd(xsum, v, w1, b1, w2, b2):
    n, c = xsum.shape
    dev = xsum.device
    ca = torch.empty((n, c), dtype=torch.float32, device=dev)
    z = torch.empty((n, c // 8), dtype=torch.float32, device=dev)
    mean = torch.empty((n, c), dtype=torch.float32, device=dev)
    check(_L().b3d_gate_se_fwd(ptr(xsum), c_ll(v), ptr(w1), ptr(b1), ptr(w2), ptr(b2), ptr(ca), ptr(z), ptr(mean), c_int(n),
                               c_int(c), stream_ptr()))
    return ca, z, mean


def gate_apply_fwd(x, psi_raw, st_psi, gpsi, bpsi_n, ca, out):
    n, v, c = _nvc(x)
    check(_L().b3d_gate_apply_fwd(ptr(x), c_ll(ld(x)), ptr(psi_raw), ptr(st_psi), ptr(gpsi), ptr(bpsi_n), ptr(ca), ptr(out),
                                  c_ll(ld(out)), c_int(n), c_ll(v), c_int(c), c_float(EPS), stream_ptr()))
    return out


def gate_apply_bwd(dout, x, psi_raw, st_psi, gpsi, bpsi_n, ca, dx):
    n, v, c = _nvc(x)
    dev = x.device
    dpsin = torch.empty((n, v), dtype=torch.float32, device=dev)
    dca = zeros_scratch((n, c), torch.float64, dev)
    st_dpsi = zeros_scratch((n, 2), torch.float64, dev)
    check(_L().b3d_gate_apply_bwd(ptr(dout), c_ll(ld(dout)), ptr(x), c_ll(ld(x)), ptr(psi_raw), ptr(st_psi), ptr(gpsi),
                                  ptr(bpsi_n), ptr(ca), ptr(dx), c_ll(ld(dx)), ptr(dpsin), ptr(dca), ptr(st_dpsi), c_int(n),
                                  c_ll(v), c_int(c), c_float(EPS), stream_ptr()))
    return dpsin, dca, st_dpsi


def gate_se_bwd(dca, ca, z, mean, w1, w2, v):
    n, c = ca.shape
    dev = ca.device
    # the four small gradient accumulators come out of ONE zero-filled arena slice (no fill kernel each)
    r = c // 8
    buf = zeros_scratch((2 * r * c + r + c,), torch.float32, dev)
    dw1, dw2 = buf[:r * c].view(r, c), buf[r * c:2 * r * c].view(c, r)
    db1, db2 = buf[2 * r * c:2 * r * c + r], buf[2 * r * c + r:]
    xadd = torch.empty((n, c), dtype=torch.float32, device=dev)
    check(_L().b3d_gate_se_bwd(ptr(dca), ptr(ca), ptr(z), ptr(mean), ptr(w1), ptr(w2), c_ll(v), ptr(dw1), ptr(db1), ptr(dw2),
                               ptr(db2), ptr(xadd), c_int(n), c_int(c), stream_ptr()))
    return dw1, db1, dw2, db2, xadd


def gate_psi_bwd(dpsin, psi_raw, st_psi, st_dpsi, gpsi, g1r, x1r, st_g, st_x, gam_g, bet_g, gam_x, bet_x, wpsi):
    n, v, f = _nvc(g1r)
    dev = g1r.device
    dz = torch.empty_like(g1r)
    sums_g = zeros_scratch((n, f, 2), torch.float64, dev)
    sums_x = zeros_scratch((n, f, 2), torch.float64, dev)
    small = zeros_scratch((f + 3,), torch.float32, dev)  # dwpsi[f], dbpsi, dgpsi, dbpsi_n
    check(_L().b3d_gate_psi_bwd(ptr(dpsin), ptr(psi_raw), ptr(st_psi), ptr(st_dpsi), ptr(gpsi), ptr(g1r), ptr(x1r), ptr(st_g),
                                ptr(st_x), ptr(gam_g), ptr(bet_g), ptr(gam_x), ptr(bet_x), ptr(wpsi), ptr(dz), ptr(sums_g),
                                ptr(sums_x), c_vp(small.data_ptr()), c_vp(small.data_ptr() + 4 * f),
                                c_vp(small.data_ptr() + 4 * (f + 1)), c_vp(small.data_ptr() + 4 * (f + 2)), c_int(n), c_ll(v),
                                c_int(f), c_float(EPS), stream_ptr()))
    return dz, sums_g, sums_x, small[:f], small[f:f + 1], small[f + 1:f + 2], small[f + 2:f + 3]


def add_channel_const(dx, xadd):
    n, v, c = _nvc(dx)
    check(_L().b3d_add_channel_const(ptr(dx), c_ll(ld(dx)), ptr(xadd), c_int(n), c_ll(v), c_int(c), stream_ptr()))
    return dx


# ---------------------------------------------------------------------------------------------------------------
# heads
# ---------------------------------------------------------------------------------------------------------------
def ds_head_fwd(x, w, b):
    """x [N,D,H,W,C] -> low-res logits fp32 [N, D, H, W, 4] (float4 per voxel)."""
    n, d, h, wd, c = x.shape
    k = w.shape[0]
    out = torch.empty((n, d, h, wd, k), dtype=torch.float32, device=x.device)
    check(_L().b3d_ds_head_fwd(ptr(x), c_ll(ld(x)), ptr(w), ptr(b), ptr(out), c_ll(n * d * h * wd), c_int(c), c_int(k),
                               stream_ptr()))
    return out


def ds_head_bwd(dl_planar, x, w, dx, accumulate):
    """dl_planar fp32 [N,K,D,H,W] (low-res) ; accumulates into dx (bf16 NDHWC) ; returns (dW [K,C], db [K])."""
    n, d, h, wd, c = x.shape
    k = w.shape[0]
    buf = zeros_scratch((k * c + k,), torch.float32, x.device)
    dw, db = buf[:k * c].view(k, c), buf[k * c:]
    check(_L().b3d_ds_head_bwd(ptr(dl_planar), ptr(x), c_ll(ld(x)), ptr(w), ptr(dx), c_ll(ld(dx)), c_int(1 if accumulate else 0),
                               ptr(dw), ptr(db), c_int(n), c_ll(d * h * wd), c_int(c), c_int(k), stream_ptr()))
    return dw, db


def ds_head_bwd_cl(dl_cl, x, w, dx, accumulate):
    """As ds_head_bwd, with the logit gradient channel-last: dl_cl fp32 [N,D,H,W,K] (what the fused loss produces)."""
    n, d, h, wd, c = x.shape
    k = w.shape[0]
    dl_cl = dl_cl.contiguous()
    assert dl_cl.dtype == torch.float32 and tuple(dl_cl.shape) == (n, d, h, wd, k)
    buf = zeros_scratch((k * c + k,), torch.float32, x.device)
    dw, db = buf[:k * c].view(k, c), buf[k * c:]
    check(_L().b3d_ds_head_bwd_cl(ptr(dl_cl), ptr(x), c_ll(ld(x)), ptr(w), ptr(dx), c_ll(ld(dx)), c_int(1 if accumulate else 0),
                                  ptr(dw), ptr(db), c_int(n), c_ll(d * h * wd), c_int(c), c_int(k), stream_ptr()))
    return dw, db


def trilinear_up_fwd(lo, size):
    n, dl, hl, wl, k = lo.shape
    d, h, w = size
    out = torch.empty((n, k, d, h, w), dtype=torch.float32, device=lo.device)
    check(_L().b3d_trilinear_up_fwd(ptr(lo), ptr(out), c_int(n), c_int(dl), c_int(hl), c_int(wl), c_int(d), c_int(h), c_int(w),
                                    c_int(k), stream_ptr()))
    return out


def trilinear_up_bwd(dup, lo_size):
    n, k, d, h, w = dup.shape
    dl, hl, wl = lo_size
    dup = dup.contiguous()
    dlo = torch.empty((n, k, dl, hl, wl), dtype=torch.float32, device=dup.device)
    tmp = torch.empty(n * k * d * h * wl + n * k * d * hl * wl, dtype=torch.float32, device=dup.device)
    check(_L().b3d_trilinear_up_bwd(ptr(dup), ptr(dlo), ptr(tmp), c_int(n), c_int(dl), c_int(hl), c_int(wl), c_int(d), c_int(h),
                                    c_int(w), c_int(k), stream_ptr()))
    return dlo


def final_bn_prepare(stats, count, train, running_mean, running_var, num_batches, momentum, update_running):
    f2 = running_mean.numel()
    bn = torch.empty(2 * f2, dtype=torch.float32, device=running_mean.device)
    check(_L().b3d_final_bn_prepare(ptr(stats), c_double(float(count)), c_int(1 if train else 0), ptr(running_mean),
                                    ptr(running_var), ptr(num_batches), c_float(momentum), c_float(EPS), ptr(bn), c_int(f2),
                                    c_int(1 if update_running else 0), stream_ptr()))
    return bn


def final_head_fwd(h, bn, gamma, beta, w2, b2):
    n, d, hh, w, f2 = h.shape
    k = w2.shape[0]
    out = torch.empty((n, k, d, hh, w), dtype=torch.float32, device=h.device)
    check(_L().b3d_final_head_fwd(ptr(h), c_ll(ld(h)), ptr(bn), ptr(gamma), ptr(beta), ptr(w2), ptr(b2), ptr(out), c_int(n),
                                  c_ll(d * hh * w), c_int(f2), c_int(k), stream_ptr()))
    return out


def final_head_bwd(dl, h, bn, gamma, beta, w2, train):
    n, d, hh, w, f2 = h.shape
    k = w2.shape[0]
    dev = h.device
    red = zeros_scratch((2 * f2 + k * f2 + k,), torch.float64, dev)
    dh = torch.empty_like(h, memory_format=torch.contiguous_format)
    dgamma = torch.empty(f2, dtype=torch.float32, device=dev)
    dbeta = torch.empty(f2, dtype=torch.float32, device=dev)
    dw2 = torch.empty((k, f2), dtype=torch.float32, device=dev)
    db2 = torch.empty(k, dtype=torch.float32, device=dev)
    check(_L().b3d_final_head_bwd(ptr(dl), ptr(h), c_ll(ld(h)), ptr(bn), ptr(gamma), ptr(beta), ptr(w2), ptr(red),
                                  c_int(1 if train else 0), ptr(dh), c_ll(ld(dh)), ptr(dgamma), ptr(dbeta), ptr(dw2), ptr(db2),
                                  c_int(n), c_ll(d * hh * w), c_int(f2), c_int(k), stream_ptr()))
    return dh, dgamma, dbeta, dw2, db2


# ---------------------------------------------------------------------------------------------------------------
# loss and metrics
# ---------------------------------------------------------------------------------------------------------------
def loss_cfg(w_dice=0.0, smooth=1e-5, w_focal=0.0, f_alpha=0.25, f_gamma=2.0, w_ce=0.0, w_boundary=0.0, w_tv=0.0,
             tv_alpha=0.7, tv_beta=0.3, tv_smooth=1e-5):
    return (ctypes.c_float * 11)(w_dice, smooth, w_focal, f_alpha, f_gamma, w_ce, w_boundary, w_tv, tv_alpha, tv_beta,
                                 tv_smooth)


def _check_target(target, shape, device, who):
    """The kernels read labels as `const long long*`: anything else would be read out of bounds.  The reference
    (`F.one_hot` / `F.cross_entropy`, losses.py:20,31) raises for non-int64 labels too."""
    if target.dtype != torch.int64:
        raise _lib.B3DError("%s: target must be int64 class indices (got %s)" % (who, target.dtype))
    if target.device != device:
        raise _lib.B3DError("%s: target is on %s, logits on %s" % (who, target.device, device))
    if tuple(target.shape) != tuple(shape):
        raise _lib.B3DError("%s: target shape %s does not match logits' [N,D,H,W] = %s" % (who, tuple(target.shape), tuple(shape)))
    return target.contiguous()


def loss_fwd(logits, target, cfg):
    """Returns (values[6] = total,dice,focal,boundary,ce,tversky ; saved tuple for loss_bwd)."""
    n, k, d, h, w = logits.shape
    dev = logits.device
    logits = logits.contiguous()
    target = _check_target(target, (n, d, h, w), dev, "loss")
    prob = torch.empty_like(logits)
    need_e = cfg[6] != 0.0
    e = torch.empty_like(logits) if need_e else None
    acc = torch.empty((n, 16), dtype=torch.float64, device=dev)
    values = torch.empty(6, dtype=torch.float32, device=dev)
    check(_L().b3d_loss_fwd(ptr(logits), ptr(target), cfg, ptr(prob), ptr(e), ptr(acc), ptr(values), c_int(n), c_int(k),
                            c_int(d), c_int(h), c_int(w), stream_ptr()))
    return values, (prob, e, target, acc)


def target_u8(target):
    """int64 labels [N,D,H,W] -> uint8 (one 33 MB read per loss call instead of one per output and pass)."""
    n, d, h, w = target.shape
    target = _check_target(target, (n, d, h, w), target.device, "loss")
    out = torch.empty((n, d, h, w), dtype=torch.uint8, device=target.device)
    check(_L().b3d_target_u8(ptr(target), ptr(out), c_ll(target.numel()), stream_ptr()))
    return out


def dsloss_fwd(lo, tgt_u8, cfg, size):
    """Fused trilinear-upsample + CombinedLoss3D of ONE deep-supervision output from its low-res logits.
    lo: fp32 [N, D/s, H/s, W/s, 4] channel-last; tgt_u8: uint8 [N, D, H, W]; returns (values[6], acc)."""
    n, dl, hl, wl, k = lo.shape
    d, h, w = size
    s = d // dl
    if k != 4 or dl * s != d or hl * s != h or wl * s != w or s not in (1, 2, 4, 8):
        raise _lib.B3DError("dsloss: low-res logits %s do not up-sample to %s by 1/2/4/8" % (tuple(lo.shape), tuple(size)))
    if tuple(tgt_u8.shape) != (n, d, h, w) or tgt_u8.dtype != torch.uint8:
        raise _lib.B3DError("dsloss: target must be uint8 [N,D,H,W] = %s" % ((n, d, h, w),))
    lo = lo.contiguous()
    acc = torch.empty((n, 16), dtype=torch.float64, device=lo.device)
    values = torch.empty(6, dtype=torch.float32, device=lo.device)
    with _prof_bw("dsloss", n * d * h * w * 1 + lo.numel() * 4, "dsloss_fwd s%d" % s):
        check(_L().b3d_dsloss_fwd(ptr(lo), ptr(tgt_u8), cfg, ptr(acc), ptr(values), c_int(n), c_int(s), c_int(d), c_int(h), c_int(w),
                                  stream_ptr()))
    return values, acc


def dsloss_bwd(lo, tgt_u8, acc, cfg, gscale, wscale, size):
    """d(loss)/d(lo) fp32 [N, D/s, H/s, W/s, 4] (accumulated in fp64 for s > 1)."""
    n, dl, hl, wl, k = lo.shape
    d, h, w = size
    s = d // dl
    out = torch.empty(lo.shape, dtype=torch.float32 if s == 1 else torch.float64, device=lo.device)
    with _prof_bw("dsloss", n * d * h * w * 1 + lo.numel() * (8 if s == 1 else 12), "dsloss_bwd s%d" % s):
        check(_L().b3d_dsloss_bwd(ptr(lo), ptr(tgt_u8), ptr(acc), cfg, ptr(gscale), c_float(wscale), ptr(out), c_int(n), c_int(s),
                                  c_int(d), c_int(h), c_int(w), stream_ptr()))
    return out if s == 1 else out.float()


def loss_bwd(saved, cfg, gscale, wscale, shape):
    prob, e, target, acc = saved
    n, k, d, h, w = shape
    dlogits = torch.empty(shape, dtype=torch.float32, device=prob.device)
    check(_L().b3d_loss_bwd(ptr(prob), ptr(e), ptr(target), ptr(acc), cfg, ptr(gscale), c_float(wscale), ptr(dlogits), c_int(n),
                            c_int(k), c_int(d), c_int(h), c_int(w), stream_ptr()))
    return dlogits


def confusion(logits, target=None, want_mask=False):
    """(hist int64 [K,K] with H[pred,true], mask uint8 [N,D,H,W] or None) — exact integer counts."""
    n, k, d, h, w = logits.shape
    dev = logits.device
    logits = logits.contiguous()
    if logits.dtype != torch.float32:
        logits = logits.float()
    hist = torch.zeros((k, k), dtype=torch.int64, device=dev)
    mask = torch.empty((n, d, h, w), dtype=torch.uint8, device=dev) if want_mask else None
    tgt = _check_target(target, (n, d, h, w), dev, "metric") if target is not None else None
    check(_L().b3d_confusion(ptr(logits), ptr(tgt), ptr(mask), ptr(hist), c_int(n), c_int(k), c_ll(d * h * w), stream_ptr()))
    return hist, mask


def voxel_counts(mask):
    """mask uint8 [D,H,W] -> (per-class counts int64[4], per-slice (last axis) tumour counts int64[W])."""
    d, h, w = mask.shape
    mask = mask.contiguous()
    cls = torch.zeros(4, dtype=torch.int64, device=mask.device)
    sl = torch.zeros(w, dtype=torch.int64, device=mask.device)
    check(_L().b3d_voxel_counts(ptr(mask), c_ll(d * h * w), c_int(w), ptr(cls), ptr(sl), stream_ptr()))
    return cls, sl
