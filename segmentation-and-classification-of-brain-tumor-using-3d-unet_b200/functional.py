"""Explicit forward / backward schedules of the U-Net blocks over the sm_100a kernels (no autograd inside).

Every `*_fwd` returns (output, saved) and every `*_bwd(saved, dout)` returns the input gradient(s) plus a dict
{param_name: fp32 grad in the reference layout}.  Activations are NDHWC bf16 (see ops.py); parameters are the fp32
`nn.Parameter`s of the reference module tree (state_dict-compatible), looked up by their reference names.

Reference being reproduced: /root/reference/main.py:102-299 (UNet3D / DoubleConv3D / AttentionGate3D).
"""
import torch

from . import ops

def packed(param, mode):
    """bf16 tcgen05-ready copy of a conv weight, cached ON the parameter object until the parameter changes.  A cache entry is
    valid while (a) the tensor's version counter and storage address are unchanged (in-place updates, load_state_dict,
    .to(device)) and (b) no optimizer stepped since: torch's fused / capturable optimizers update parameters WITHOUT bumping
    the version counter, so a global optimizer post-step hook (registered at import) advances the cache epoch, as does a
    replayed CUDA graph (graph.py).  Anything else that writes parameter memory behind autograd's back must call
    `clear_pack_cache()`.  (A global cache keyed by id(param) is wrong: a new model can reuse the id, the version and —
    through the caching allocator — even the address of a freed parameter.)
    Both copies a layer needs (fprop + dgrad) are produced by one launch from one read of the fp32 weight."""
    pin = param.__dict__.get("_b3d_pack_pinned")
    if pin is not None:   # optim.FusedAdamW wrote these copies in the same pass that updated the parameter
        if pin[0] == param._version and pin[1] == param.data_ptr():
            return pin[2][mode]
        del param.__dict__["_b3d_pack_pinned"]   # the parameter was changed behind the optimizer (load_state_dict, .to, ...)
    cache = param.__dict__.setdefault("_b3d_pack", {})
    key = (param._version, param.data_ptr(), _PACK_EPOCH[0])
    hit = cache.get(mode)
    if hit is not None and hit[0] == key:
        return hit[1], hit[2], hit[3]
    if ops.PAIR_PACK and param.is_cuda:
        for m, (wp, kp, rows) in ops.pack_weight_pair(param, mode in (ops.PACK_CONVT_FPROP, ops.PACK_CONVT_DGRAD)).items():
            cache[m] = (key, wp, kp, rows)
    else:
        wp, kp, rows = ops.pack_weight(param, mode)
        cache[mode] = (key, wp, kp, rows)
    hit = cache[mode]
    return hit[1], hit[2], hit[3]


_PACK_EPOCH = [0]


def clear_pack_cache():
    """Invalidate every cached packed weight (see `packed`)."""
    _PACK_EPOCH[0] += 1


def _optimizer_stepped(optimizer, args, kwargs):
    _PACK_EPOCH[0] += 1


try:   # every torch optimizer step invalidates the packed weights (fused AdamW does not bump tensor version counters)
    from torch.optim.optimizer import register_optimizer_step_post_hook as _reg_hook
    _OPT_HOOK = _reg_hook(_optimizer_stepped)
except ImportError:   # pragma: no cover - very old torch: fall back to "never cache across calls that enable grad"
    _OPT_HOOK = None


def act_padded(n, d, h, w, c, device):
    """Activation whose channel count is not a multiple of 16 lives in a zero-padded 16-multiple buffer (tensor-core K /
    N granularity); the returned tensor is the [..., :c] slice, `.base_full` the padded view."""
    cp = ops.roundup(c, 16)
    if cp == c:
        t = ops.new_act(n, d, h, w, c, device)
        return t, t
    full = ops.new_act(n, d, h, w, cp, device, zero=True)
    return full[..., :c], full


GRAD_SINK = None   # data parallel: callable(name, shape) -> fp32 tensor the weight-gradient kernel writes into (or None)


def _sink(name, shape):
    return GRAD_SINK(name, tuple(shape)) if GRAD_SINK is not None else None


def bias_grad(dy):
    """Σ over all voxels and samples -> fp32 [C].  The batch is folded into the voxel axis (a view), so the per-sample sums of
    `channel_sum` need no second reduction kernel."""
    n, d, h, w, c = dy.shape
    if dy.stride(0) == d * dy.stride(1):
        return ops.channel_sum(dy.as_strided((1, n * d, h, w, c), (n * dy.stride(0), dy.stride(1), dy.stride(2), dy.stride(3), 1)))[0].float()
    return ops.channel_sum(dy).sum(dim=0).float()


# ----------------------------------------------------------------------------------------------------------------
# DoubleConv3D (main.py:205-242)
# ----------------------------------------------------------------------------------------------------------------
def double_conv_fwd(x, p, pre, cin_real, need_bwd):
    """x: [N,D,H,W,Cx] with Cx = roundup16(cin_real) (zero padded).  p: dict name -> Parameter."""
    w1, w2 = p[pre + "double_conv.0.weight"], p[pre + "double_conv.3.weight"]
    cout = w1.shape[0]
    w1p, _, rows1 = packed(w1, ops.PACK_FPROP)
    has_res_conv = (pre + "residual.0.weight") in p
    r = st_r = None
    if has_res_conv:   # the residual projection only depends on x: a side branch that overlaps with the 3x3x3 convs
        wrp, _, rowsr = packed(p[pre + "residual.0.weight"], ops.PACK_FPROP)
        with ops.side_branch(ops.BRANCH_MASK & 1, x) as br:
            r, st_r = ops.conv_fprop(x, wrp, rowsr, cout, 1, groups=8, cin_real=cin_real)
    y1, st1 = ops.conv_fprop(x, w1p, rows1, cout, 3, groups=8, cin_real=cin_real)
    a1 = ops.gn_apply(y1, st1, p[pre + "double_conv.1.weight"], p[pre + "double_conv.1.bias"], 8, True)
    w2p, _, rows2 = packed(w2, ops.PACK_FPROP)
    y2, st2 = ops.conv_fprop(a1, w2p, rows2, cout, 3, groups=8)
    if has_res_conv:
        br.join()
        out = ops.gn_apply(y2, st2, p[pre + "double_conv.4.weight"], p[pre + "double_conv.4.bias"], 8, True, res=r,
                           res_stats=st_r, res_gamma=p[pre + "residual.1.weight"], res_beta=p[pre + "residual.1.bias"],
                           res_groups=8)
    else:  # in == out: identity residual (main.py:225-226)
        out = ops.gn_apply(y2, st2, p[pre + "double_conv.4.weight"], p[pre + "double_conv.4.bias"], 8, True, res=x)
    saved = (x, y1, st1, a1, y2, st2, r, st_r, cin_real) if need_bwd else None
    return out, saved


def double_conv_bwd(saved, dout, p, pre, need_dx=True):
    x, y1, st1, a1, y2, st2, r, st_r, cin_real = saved
    grads = {}
    w1, w2 = p[pre + "double_conv.0.weight"], p[pre + "double_conv.3.weight"]
    cout = w1.shape[0]
    # tail: out = relu(GN(y2)) + GN_r(r).  Both branches see dout: one dual kernel pair reads it once per phase.
    br = None
    dual = ops.gn_bwd_dual(dout, y2, st2, p[pre + "double_conv.4.weight"], p[pre + "double_conv.4.bias"], r, st_r,
                           p[pre + "residual.1.weight"], 8) if r is not None else None
    if dual is not None:
        (dy2, grads[pre + "double_conv.4.weight"], grads[pre + "double_conv.4.bias"],
         dr, grads[pre + "residual.1.weight"], grads[pre + "residual.1.bias"]) = dual
    else:
        if r is not None:   # residual-branch GroupNorm backward: independent of the main chain until the final dx add
            with ops.side_branch(ops.BRANCH_MASK & 4, dout, r) as br:
                dr, grads[pre + "residual.1.weight"], grads[pre + "residual.1.bias"] = ops.gn_bwd(
                    dout, r, st_r, p[pre + "residual.1.weight"], p[pre + "residual.1.bias"], 8, False)
        dy2, grads[pre + "double_conv.4.weight"], grads[pre + "double_conv.4.bias"] = ops.gn_bwd(
            dout, y2, st2, p[pre + "double_conv.4.weight"], p[pre + "double_conv.4.bias"], 8, True)
    grads[pre + "double_conv.3.weight"] = ops.conv_wgrad(a1, dy2, cout, cout, 3, dw=_sink(pre + "double_conv.3.weight", w2.shape))
    w2d, _, rows2d = packed(w2, ops.PACK_DGRAD)
    da1, _ = ops.conv_fprop(dy2, w2d, rows2d, cout, 3)
    del dy2
    dy1, grads[pre + "double_conv.1.weight"], grads[pre + "double_conv.1.bias"] = ops.gn_bwd(
        da1, y1, st1, p[pre + "double_conv.1.weight"], p[pre + "double_conv.1.bias"], 8, True, dx=da1)
    grads[pre + "double_conv.0.weight"] = ops.conv_wgrad(x, dy1, cin_real, cout, 3, dw=_sink(pre + "double_conv.0.weight", w1.shape))
    dx = None
    if need_dx:
        w1d, _, rows1d = packed(w1, ops.PACK_DGRAD)
        dx, _ = ops.conv_fprop(dy1, w1d, rows1d, x.shape[-1], 3)
    del dy1
    if r is not None:
        if br is not None:
            br.join()
        grads[pre + "residual.0.weight"] = ops.conv_wgrad(x, dr, cin_real, cout, 1,
                                                          dw=_sink(pre + "residual.0.weight", p[pre + "residual.0.weight"].shape))
        if need_dx:
            wrd, _, rowsrd = packed(p[pre + "residual.0.weight"], ops.PACK_DGRAD)
            ops.conv_fprop(dr, wrd, rowsrd, x.shape[-1], 1, out=dx, add=dx)   # dx += residual-branch gradient, fused
    elif need_dx:
        ops.add_bf16(dx, dout, out=dx)
    return dx, grads


# ----------------------------------------------------------------------------------------------------------------
# AttentionGate3D (main.py:244-299)
# ----------------------------------------------------------------------------------------------------------------
def gate_fwd(g, x, p, pre, out, need_bwd):
    """g (gating, decoder) and x (skip) : [N,D,H,W,C]; writes x * psi * ca into `out` (may be a concat-buffer slice)."""
    n, d, h, w, c = x.shape
    f = p[pre + "W_g.0.weight"].shape[0]
    dev = x.device
    wgp, _, rowsg = packed(p[pre + "W_g.0.weight"], ops.PACK_FPROP)
    wxp, _, rowsx = packed(p[pre + "W_x.0.weight"], ops.PACK_FPROP)
    g1r, g1r_full = act_padded(n, d, h, w, f, dev)
    x1r, x1r_full = act_padded(n, d, h, w, f, dev)
    _, st_g = ops.conv_fprop(g, wgp, rowsg, f, 1, bias=p[pre + "W_g.0.bias"], groups=4, out=g1r)
    _, st_x = ops.conv_fprop(x, wxp, rowsx, f, 1, bias=p[pre + "W_x.0.bias"], groups=4, out=x1r)
    g1c = g1r if g1r.is_contiguous() else g1r.contiguous()
    x1c = x1r if x1r.is_contiguous() else x1r.contiguous()
    psi_raw, st_psi = ops.gate_psi_fwd(g1c, x1c, st_g, st_x, p[pre + "W_g.1.weight"], p[pre + "W_g.1.bias"],
                                       p[pre + "W_x.1.weight"], p[pre + "W_x.1.bias"], p[pre + "psi.0.weight"],
                                       p[pre + "psi.0.bias"])
    xsum = ops.channel_sum(x)
    ca, z, mean = ops.gate_se_fwd(xsum, d * h * w, p[pre + "channel_attention.1.weight"], p[pre + "channel_attention.1.bias"],
                                  p[pre + "channel_attention.3.weight"], p[pre + "channel_attention.3.bias"])
    ops.gate_apply_fwd(x, psi_raw, st_psi, p[pre + "psi.1.weight"], p[pre + "psi.1.bias"], ca, out)
    saved = (g, x, g1c, x1c, st_g, st_x, psi_raw, st_psi, ca, z, mean) if need_bwd else None
    return out, saved


def gate_bwd(saved, dout, p, pre, dg_add=None, defer_const=False):
    """Returns (dg [+ dg_add], dx, grads).  defer_const: do not add the channel-attention branch's per-(sample, channel) constant
    to dx here — return it as a 4th value so the caller's next pass over dx (pool_bwd) adds it for free."""
    g, x, g1r, x1r, st_g, st_x, psi_raw, st_psi, ca, z, mean = saved
    n, d, h, w, c = x.shape
    f = g1r.shape[-1]
    dev = x.device
    v = d * h * w
    grads = {}
    dx = ops.new_act(n, d, h, w, c, dev)
    dpsin, dca, st_dpsi = ops.gate_apply_bwd(dout, x, psi_raw, st_psi, p[pre + "psi.1.weight"], p[pre + "psi.1.bias"], ca, dx)
    w1, w2 = p[pre + "channel_attention.1.weight"], p[pre + "channel_attention.3.weight"]
    dw1, db1, dw2, db2, xadd = ops.gate_se_bwd(dca, ca, z, mean, w1, w2, v)
    grads[pre + "channel_attention.1.weight"] = dw1.reshape(w1.shape)
    grads[pre + "channel_attention.1.bias"] = db1
    grads[pre + "channel_attention.3.weight"] = dw2.reshape(w2.shape)
    grads[pre + "channel_attention.3.bias"] = db2
    dz, sums_g, sums_x, dwpsi, dbpsi, dgpsi, dbpsi_n = ops.gate_psi_bwd(
        dpsin, psi_raw, st_psi, st_dpsi, p[pre + "psi.1.weight"], g1r, x1r, st_g, st_x, p[pre + "W_g.1.weight"],
        p[pre + "W_g.1.bias"], p[pre + "W_x.1.weight"], p[pre + "W_x.1.bias"], p[pre + "psi.0.weight"])
    grads[pre + "psi.0.weight"] = dwpsi.reshape(p[pre + "psi.0.weight"].shape)
    grads[pre + "psi.0.bias"] = dbpsi
    grads[pre + "psi.1.weight"] = dgpsi
    grads[pre + "psi.1.bias"] = dbpsi_n
    # GN4 backward of both branches (the reductions were fused into gate_psi_bwd)
    dg1r, dg1r_full = act_padded(n, d, h, w, f, dev)
    dx1r, dx1r_full = act_padded(n, d, h, w, f, dev)
    _, grads[pre + "W_g.1.weight"], grads[pre + "W_g.1.bias"] = ops.gn_bwd(
        dz, g1r, st_g, p[pre + "W_g.1.weight"], p[pre + "W_g.1.bias"], 4, False, dx=dg1r, sums=sums_g)
    _, grads[pre + "W_x.1.weight"], grads[pre + "W_x.1.bias"] = ops.gn_bwd(
        dz, x1r, st_x, p[pre + "W_x.1.weight"], p[pre + "W_x.1.bias"], 4, False, dx=dx1r, sums=sums_x)
    del dz
    grads[pre + "W_g.0.weight"] = ops.conv_wgrad(g, dg1r, c, f, 1, dw=_sink(pre + "W_g.0.weight", p[pre + "W_g.0.weight"].shape))
    grads[pre + "W_g.0.bias"] = bias_grad(dg1r)
    grads[pre + "W_x.0.weight"] = ops.conv_wgrad(x, dx1r, c, f, 1, dw=_sink(pre + "W_x.0.weight", p[pre + "W_x.0.weight"].shape))
    grads[pre + "W_x.0.bias"] = bias_grad(dx1r)
    wgd, _, rowsgd = packed(p[pre + "W_g.0.weight"], ops.PACK_DGRAD)
    wxd, _, rowsxd = packed(p[pre + "W_x.0.weight"], ops.PACK_DGRAD)
    # dg is only ever added to the gradient of the up-sampled half of the concat buffer: fuse that add (dg_add)
    dg, _ = ops.conv_fprop(dg1r_full, wgd, rowsgd, c, 1, add=dg_add)
    ops.conv_fprop(dx1r_full, wxd, rowsxd, c, 1, out=dx, add=dx)                 # dx += W_x-branch gradient, fused
    if defer_const:
        return dg, dx, grads, xadd
    ops.add_channel_const(dx, xadd)
    return dg, dx, grads


# ----------------------------------------------------------------------------------------------------------------
# decoder step: ConvTranspose3d(k2,s2) -> attention gate -> concat   (main.py:181-195; cat order: attended skip first)
# ----------------------------------------------------------------------------------------------------------------
def up_gate_fwd(x_low, skip, p, idx, need_bwd):
    n, d, h, w, c = skip.shape
    dev = skip.device
    wt = p["ups.%d.weight" % idx]
    cin = wt.shape[0]
    cat = ops.new_act(n, d, h, w, 2 * c, dev)
    wtp, _, _ = packed(wt, ops.PACK_CONVT_FPROP)
    ops.convT2_fprop(x_low, wtp, p["ups.%d.bias" % idx], c, out=cat[..., c:])
    _, gsaved = gate_fwd(cat[..., c:], skip, p, "ups.%d." % (idx + 1), cat[..., :c], need_bwd)
    saved = (x_low, gsaved, cin, c) if need_bwd else None
    return cat, saved


def up_gate_bwd(saved, dcat, p, idx):
    """Returns (dx_low, dskip, grads)."""
    x_low, gsaved, cin, c = saved
    du, dskip, grads, xadd = gate_bwd(gsaved, dcat[..., :c], p, "ups.%d." % (idx + 1), dg_add=dcat[..., c:], defer_const=True)
    wt = p["ups.%d.weight" % idx]
    grads["ups.%d.weight" % idx] = ops.convT2_wgrad(x_low, du, cin, c, dw=_sink("ups.%d.weight" % idx, wt.shape))
    grads["ups.%d.bias" % idx] = bias_grad(du)
    wtd, _, rowsd = packed(wt, ops.PACK_CONVT_DGRAD)
    dx_low = ops.convT2_dgrad(du, wtd, rowsd, cin)
    return dx_low, (dskip, xadd), grads


# ----------------------------------------------------------------------------------------------------------------
# deep-supervision head (main.py:137-140,164-171) and final head (main.py:129-134)
# ----------------------------------------------------------------------------------------------------------------
def ds_head_fwd(skip, p, i):
    """Low-res logits of deep-supervision head i: fp32 channel-last [N,d,h,w,K].  The trilinear up-sampling to full resolution
    (main.py:165-170) is NOT done here: the fused loss interpolates in registers, anything else materialises lazily (lazy.py)."""
    k = p["deep_supervision.%d.weight" % i].shape[0]
    return ops.ds_head_fwd(skip, p["deep_supervision.%d.weight" % i].reshape(k, -1), p["deep_supervision.%d.bias" % i])


def ds_head_bwd(skip, dlo, p, i, dskip):
    """dlo: gradient of the LOW-RES logits, fp32 channel-last [N,d,h,w,K].  Accumulates into dskip; returns grads."""
    wgt = p["deep_supervision.%d.weight" % i]
    k = wgt.shape[0]
    dw, db = ops.ds_head_bwd_cl(dlo, skip, wgt.reshape(k, -1), dskip, accumulate=True)
    return {"deep_supervision.%d.weight" % i: dw.reshape(wgt.shape), "deep_supervision.%d.bias" % i: db}


def final_fwd(x, p, bufs, training, need_bwd):
    """final_conv = Conv3(f0 -> f0/2, bias) + BatchNorm3d + ReLU + Conv1(f0/2 -> K, bias); fp32 NCDHW logits out."""
    n, d, h, w, c = x.shape
    dev = x.device
    w0 = p["final_conv.0.weight"]
    f2 = w0.shape[0]
    w0p, _, rows0 = packed(w0, ops.PACK_FPROP)
    hbuf, hfull = act_padded(n, d, h, w, f2, dev)
    _, st = ops.conv_fprop(x, w0p, rows0, f2, 3, bias=p["final_conv.0.bias"], groups=f2 if training else 0,
                           stats_batch=True, out=hbuf)
    bn = ops.final_bn_prepare(st, n * d * h * w, training, bufs["final_conv.1.running_mean"],
                              bufs["final_conv.1.running_var"], bufs["final_conv.1.num_batches_tracked"], 0.1,
                              update_running=training)
    w3 = p["final_conv.3.weight"]
    k = w3.shape[0]
    logits = ops.final_head_fwd(hbuf, bn, p["final_conv.1.weight"], p["final_conv.1.bias"], w3.reshape(k, f2),
                                p["final_conv.3.bias"])
    saved = (x, hbuf, bn, training) if need_bwd else None
    return logits, saved


def final_bwd(saved, dlogits, p):
    x, hbuf, bn, training = saved
    n, d, h, w, c = x.shape
    w0, w3 = p["final_conv.0.weight"], p["final_conv.3.weight"]
    f2, k = w0.shape[0], w3.shape[0]
    grads = {}
    dh, dgam, dbet, dw3, db3 = ops.final_head_bwd(dlogits.contiguous(), hbuf, bn, p["final_conv.1.weight"],
                                                  p["final_conv.1.bias"], w3.reshape(k, f2), training)
    grads["final_conv.1.weight"], grads["final_conv.1.bias"] = dgam, dbet
    grads["final_conv.3.weight"], grads["final_conv.3.bias"] = dw3.reshape(w3.shape), db3
    if f2 % 16:
        dh_s, dh_full = act_padded(n, d, h, w, f2, x.device)
        dh_s.copy_(dh)
        dh, dh_k = dh_s, dh_full
    else:
        dh_k = dh
    grads["final_conv.0.weight"] = ops.conv_wgrad(x, dh, c, f2, 3, dw=_sink("final_conv.0.weight", w0.shape))
    grads["final_conv.0.bias"] = bias_grad(dh)
    w0d, _, rows0d = packed(w0, ops.PACK_DGRAD)
    dx, _ = ops.conv_fprop(dh_k, w0d, rows0d, c, 3)
    return dx, grads


# ----------------------------------------------------------------------------------------------------------------
# whole network (main.py:154-203)
# ----------------------------------------------------------------------------------------------------------------
def unet_fwd(x_ncdhw, p, bufs, features, training, dropout_masks, need_bwd):
    """x: fp32 NCDHW.  Returns (main_logits fp32 NCDHW, [low-res deep-supervision logits, channel-last] (train only), saved)."""
    n, cin, d, h, w = x_ncdhw.shape
    if d % 32 or h % 32 or w % 32:
        raise ValueError("UNet3D (b200 path): D,H,W must be multiples of 32, got %s" % ((d, h, w),))
    nl = len(features)
    x = ops.to_ndhwc_bf16(x_ncdhw, ops.roundup(cin, 16))
    S = {"enc": [], "pool": [], "up": [], "dec": [], "n": n, "size": (d, h, w)}
    skips, deep = [], []
    c_real = cin
    for i in range(nl):
        out, sv = double_conv_fwd(x, p, "downs.%d." % i, c_real, need_bwd)
        skips.append(out)
        S["enc"].append(sv)
        if training and i < nl - 1:
            with ops.side_branch(ops.BRANCH_MASK & 2, out):
                deep.append(ds_head_fwd(out, p, i))
        mask = dropout_masks[i] if (training and dropout_masks is not None) else None
        x = ops.pool_fwd(out, mask)
        S["pool"].append((out, mask))
        c_real = features[i]
    x, S["bott"] = double_conv_fwd(x, p, "bottleneck.", c_real, need_bwd)
    for j in range(nl):
        skip = skips[nl - 1 - j]
        cat, sv = up_gate_fwd(x, skip, p, 3 * j, need_bwd)
        S["up"].append(sv)
        x, sv2 = double_conv_fwd(cat, p, "ups.%d." % (3 * j + 2), cat.shape[-1], need_bwd)
        S["dec"].append(sv2)
    logits, S["final"] = final_fwd(x, p, bufs, training, need_bwd)
    if ops.BRANCH_MASK & 2:
        ops.wgrad_join()   # the deep-supervision maps were produced on the side stream
    S["skips"] = skips if need_bwd else None
    return logits, deep, (S if need_bwd else None)


def unet_bwd(S, dmain, ddeep, p, features, on_grads=None):
    """dmain: fp32 NCDHW gradient of the main logits; ddeep: gradients of the LOW-RES deep-supervision logits, channel-last
    (entries may be None).  Returns {name: grad}.
    `on_grads(dict)` is called as soon as a block's parameter gradients exist (reverse-topological order) so a
    data-parallel wrapper can start all-reducing them while the rest of backward runs."""
    nl = len(features)
    grads = {}

    def emit(g):
        grads.update(g)
        if on_grads is not None:
            on_grads(g)

    dx, g = final_bwd(S["final"], dmain, p)
    emit(g)
    dskips = [None] * nl
    for j in reversed(range(nl)):
        dcat, g = double_conv_bwd(S["dec"][j], dx, p, "ups.%d." % (3 * j + 2))
        emit(g)
        dx, dskip, g = up_gate_bwd(S["up"][j], dcat, p, 3 * j)
        emit(g)
        dskips[nl - 1 - j] = dskip
        del dcat
    dx, g = double_conv_bwd(S["bott"], dx, p, "bottleneck.")
    emit(g)
    for i in reversed(range(nl)):
        skip, mask = S["pool"][i]
        dskip, xadd = dskips[i]   # xadd: the gate's channel-attention constant still owed to dskip, added by this pass
        ops.pool_bwd(skip, mask, dx, dx=dskip, accumulate=True, cadd=xadd)
        if i < nl - 1 and ddeep is not None and i < len(ddeep) and ddeep[i] is not None:
            emit(ds_head_bwd(skip, ddeep[i], p, i, dskip))
        dx, g = double_conv_bwd(S["enc"][i], dskip, p, "downs.%d." % i, need_dx=(i > 0))
        emit(g)
    return grads
