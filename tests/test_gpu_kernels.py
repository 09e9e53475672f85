"""Per-kernel parity of the sm_100a library (through the C ABI) against plain PyTorch fp32 on the same bf16-rounded
inputs.  Tolerances: bf16 outputs are compared at 2^-7 relative to the tensor scale (one bf16 rounding of the result);
fp32 outputs / reductions at 1e-4 relative; integer outputs bit-exact."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    import b3d  # noqa: F401
    from unet3d_b200 import ops

DEV = "cuda:0"
BF = torch.bfloat16
# the PyTorch references must be true fp32 (cuDNN/cuBLAS default to TF32 for fp32 convs/matmuls on this GPU)
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False


def _bf(*shape, seed=0, scale=1.0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(DEV).to(BF)


def _close_bf16(got, ref, tol=1.0 / 64):
    scale = max(ref.abs().max().item(), 1e-6)
    err = (got.float() - ref).abs().max().item()
    assert err <= tol * scale, "max err %g vs scale %g" % (err, scale)


def _ncdhw(x):  # NDHWC -> NCDHW float
    return x.float().permute(0, 4, 1, 2, 3).contiguous()


def _ndhwc(x):
    return x.permute(0, 2, 3, 4, 1).contiguous()


def _stats(y, groups):
    n, c = y.shape[0], y.shape[-1]
    yf = y.float().reshape(n, -1, groups, c // groups)
    return torch.stack([yf.sum(dim=(1, 3)), (yf * yf).sum(dim=(1, 3))], dim=-1).double()


@pytest.mark.parametrize("n,s,cin,cout,ks,bias", [(2, 8, 16, 32, 3, False), (1, 16, 64, 64, 3, True), (2, 8, 32, 16, 1, True),
                                                  (1, 4, 256, 512, 3, False), (2, 8, 64, 8, 1, True)])
def test_conv_fprop(n, s, cin, cout, ks, bias):
    x = _bf(n, s, s, s, cin, seed=1)
    w = (_bf(cout, cin, ks, ks, ks, seed=2).float() / (cin * ks ** 3) ** 0.5).to(BF).float()
    b = torch.randn(cout, device=DEV) if bias else None
    wp, kp, rows = ops.pack_weight(w, ops.PACK_FPROP)
    groups = 4 if cout >= 16 else 2
    y, st = ops.conv_fprop(x, wp, rows, cout, ks, bias=b, groups=groups)
    ref = _ndhwc(F.conv3d(_ncdhw(x), w, b, padding=ks // 2))
    _close_bf16(y, ref)
    rs = _stats(ref, groups)
    assert ((st - rs).abs() / (rs.abs() + 1.0)).max().item() < 2e-3


def test_conv_dgrad_via_flipped_weights():
    n, s, cin, cout = 2, 8, 32, 64
    dy = _bf(n, s, s, s, cout, seed=3)
    w = (_bf(cout, cin, 3, 3, 3, seed=4).float() / (cin * 27) ** 0.5).to(BF).float()
    wp, kp, rows = ops.pack_weight(w, ops.PACK_DGRAD)
    dx, _ = ops.conv_fprop(dy, wp, rows, cin, 3)
    ref = _ndhwc(F.conv_transpose3d(_ncdhw(dy), w, padding=1))
    _close_bf16(dx, ref)


def test_conv_channel_slices_and_batch_stats():
    n, s = 2, 8
    buf = _bf(n, s, s, s, 64, seed=5)
    x = buf[..., 32:]  # second half of a concat buffer
    w = (_bf(16, 32, 3, 3, 3, seed=6).float() / (32 * 27) ** 0.5).to(BF).float()
    wp, kp, rows = ops.pack_weight(w, ops.PACK_FPROP)
    out_buf = torch.zeros(n, s, s, s, 48, device=DEV, dtype=BF)
    y, st = ops.conv_fprop(x, wp, rows, 16, 3, groups=16, stats_batch=True, out=out_buf[..., 16:32])
    ref = _ndhwc(F.conv3d(_ncdhw(x), w, None, padding=1))
    _close_bf16(out_buf[..., 16:32], ref)
    assert out_buf[..., :16].abs().max().item() == 0 and out_buf[..., 32:].abs().max().item() == 0
    rs = torch.stack([ref.sum(dim=(0, 1, 2, 3)), (ref * ref).sum(dim=(0, 1, 2, 3))], dim=-1).double()[None]
    assert ((st - rs).abs() / (rs.abs() + 1.0)).max().item() < 2e-3


def test_convT2_fprop_and_dgrad():
    n, s, cin, cout = 2, 4, 64, 32
    x = _bf(n, s, s, s, cin, seed=7)
    w = (_bf(cin, cout, 2, 2, 2, seed=8).float() / cin ** 0.5).to(BF).float()
    b = torch.randn(cout, device=DEV)
    wp, kp, rows = ops.pack_weight(w, ops.PACK_CONVT_FPROP)
    cat = torch.zeros(n, 2 * s, 2 * s, 2 * s, 2 * cout, device=DEV, dtype=BF)
    ops.convT2_fprop(x, wp, b, cout, out=cat[..., cout:])
    ref = _ndhwc(F.conv_transpose3d(_ncdhw(x), w, b, stride=2))
    _close_bf16(cat[..., cout:], ref)
    dy = _bf(n, 2 * s, 2 * s, 2 * s, cout, seed=9)
    wpd, kpd, rowsd = ops.pack_weight(w, ops.PACK_CONVT_DGRAD)
    dx = ops.convT2_dgrad(dy, wpd, rowsd, cin)
    refdx = _ndhwc(F.conv3d(_ncdhw(dy), w, stride=2))
    _close_bf16(dx, refdx)


@pytest.mark.parametrize("c,groups,relu,res", [(32, 8, True, 0), (64, 8, True, 1), (16, 4, False, 0), (256, 8, True, 1),
                                               (2048, 8, True, 0)])   # wide model's concat width: > 48 KB of dynamic smem
def test_groupnorm_fwd_bwd(c, groups, relu, res):
    n, s = 2, 8
    y = _bf(n, s, s, s, c, seed=10, scale=2.0)
    r = _bf(n, s, s, s, c, seed=11)
    gam = (1 + 0.2 * torch.randn(c, device=DEV))
    bet = 0.2 * torch.randn(c, device=DEV)
    gam_r = (1 + 0.2 * torch.randn(c, device=DEV))
    bet_r = 0.2 * torch.randn(c, device=DEV)
    st, st_r = _stats(y, groups), _stats(r, groups)
    yt = _ncdhw(y).requires_grad_(True)
    rt = _ncdhw(r).requires_grad_(True)
    gt, bt = gam.clone().requires_grad_(True), bet.clone().requires_grad_(True)
    ref = F.group_norm(yt, groups, gt, bt, 1e-5)
    if relu:
        ref = F.relu(ref)
    if res:
        ref = ref + F.group_norm(rt, groups, gam_r, bet_r, 1e-5)
    out = ops.gn_apply(y, st, gam, bet, groups, relu, res=r if res else None, res_stats=st_r if res else None,
                       res_gamma=gam_r, res_beta=bet_r, res_groups=groups)
    _close_bf16(out, _ndhwc(ref.detach()))
    dout = _bf(n, s, s, s, c, seed=12)
    ref.backward(_ncdhw(dout))
    dx, dg, db = ops.gn_bwd(dout, y, st, gam, bet, groups, relu)
    _close_bf16(dx, _ndhwc(yt.grad), tol=1.0 / 32)
    np.testing.assert_allclose(dg.cpu().numpy(), gt.grad.cpu().numpy(), rtol=2e-2, atol=2e-2 * gt.grad.abs().max().item())
    np.testing.assert_allclose(db.cpu().numpy(), bt.grad.cpu().numpy(), rtol=2e-2, atol=2e-2 * bt.grad.abs().max().item())


def test_pool_dropout_fwd_bwd():
    n, s, c = 2, 8, 32
    x = _bf(n, s, s, s, c, seed=13)
    x[0, 0, 0, 0:2, :] = 1.5  # ties inside one window: first maximum must win
    mask = ((torch.rand(n, c, device=DEV) > 0.2).float() / 0.8).contiguous()
    xt = _ncdhw(x).requires_grad_(True)
    ref = F.max_pool3d(xt, 2, 2) * mask[:, :, None, None, None]
    out = ops.pool_fwd(x, mask)
    _close_bf16(out, _ndhwc(ref.detach()), tol=1.0 / 128)
    dy = _bf(n, s // 2, s // 2, s // 2, c, seed=14)
    ref.backward(_ncdhw(dy))
    dx = ops.pool_bwd(x, mask, dy)
    _close_bf16(dx, _ndhwc(xt.grad), tol=1.0 / 128)
    out_eval = ops.pool_fwd(x, None)
    assert torch.equal(out_eval.float(), _ndhwc(F.max_pool3d(_ncdhw(x), 2, 2)))


def test_layout_and_channel_sum():
    x = torch.randn(2, 4, 8, 8, 8, device=DEV)
    y = ops.to_ndhwc_bf16(x, 16)
    assert y.shape == (2, 8, 8, 8, 16)
    assert torch.equal(y[..., :4].float(), _ndhwc(x).to(BF).float()) and y[..., 4:].abs().max().item() == 0
    back = ops.to_ncdhw_f32(y)
    assert torch.equal(back[:, :4], x.to(BF).float())
    a = _bf(2, 8, 8, 8, 64, seed=15)
    s = ops.channel_sum(a)
    np.testing.assert_allclose(s.cpu().numpy(), a.double().sum(dim=(1, 2, 3)).cpu().numpy(), rtol=1e-5, atol=1e-3)


def test_ds_head_and_trilinear():
    n, s, c, full = 2, 8, 64, 32
    x = _bf(n, s, s, s, c, seed=16)
    w = torch.randn(4, c, device=DEV) / c ** 0.5
    b = torch.randn(4, device=DEV)
    xt = _ncdhw(x).requires_grad_(True)
    wt, btt = w.clone().requires_grad_(True), b.clone().requires_grad_(True)
    lo_ref = F.conv3d(xt, wt.reshape(4, c, 1, 1, 1), btt)
    up_ref = F.interpolate(lo_ref, size=(full, full, full), mode="trilinear", align_corners=False)
    lo = ops.ds_head_fwd(x, w, b)
    np.testing.assert_allclose(lo.permute(0, 4, 1, 2, 3).cpu().numpy(), lo_ref.detach().cpu().numpy(), rtol=1e-4, atol=1e-4)
    up = ops.trilinear_up_fwd(lo, (full, full, full))
    np.testing.assert_allclose(up.cpu().numpy(), up_ref.detach().cpu().numpy(), rtol=1e-4, atol=1e-4)
    dup = torch.randn_like(up_ref)
    up_ref.backward(dup)
    dlo = ops.trilinear_up_bwd(dup, (s, s, s))
    dx = torch.zeros_like(x)
    dw, db = ops.ds_head_bwd(dlo, x, w, dx, accumulate=False)
    _close_bf16(dx, _ndhwc(xt.grad), tol=1.0 / 64)
    np.testing.assert_allclose(dw.cpu().numpy(), wt.grad.cpu().numpy(), rtol=1e-3, atol=1e-3 * wt.grad.abs().max().item())
    np.testing.assert_allclose(db.cpu().numpy(), btt.grad.cpu().numpy(), rtol=1e-3, atol=1e-3 * btt.grad.abs().max().item())
    # identity scale
    up1 = ops.trilinear_up_fwd(lo, (s, s, s))
    np.testing.assert_allclose(up1.cpu().numpy(), lo.permute(0, 4, 1, 2, 3).cpu().numpy(), rtol=0, atol=0)


@pytest.mark.parametrize("train", [True, False])
def test_final_head(train):
    n, s, f2 = 2, 8, 16
    h = _bf(n, s, s, s, f2, seed=17, scale=1.5)
    gam = 1 + 0.2 * torch.randn(f2, device=DEV)
    bet = 0.2 * torch.randn(f2, device=DEV)
    w2 = torch.randn(4, f2, device=DEV) / f2 ** 0.5
    b2 = torch.randn(4, device=DEV)
    rm, rv = 0.1 * torch.randn(f2, device=DEV), 1 + 0.2 * torch.rand(f2, device=DEV)
    nb = torch.zeros((), dtype=torch.int64, device=DEV)
    ht = _ncdhw(h).requires_grad_(True)
    gt, bt, wt, b2t = [t.clone().requires_grad_(True) for t in (gam, bet, w2, b2)]
    rm_ref, rv_ref = rm.clone(), rv.clone()
    ref = F.conv3d(F.relu(F.batch_norm(ht, rm_ref, rv_ref, gt, bt, train, 0.1, 1e-5)), wt.reshape(4, f2, 1, 1, 1), b2t)
    hf = h.float()
    st = torch.stack([hf.sum(dim=(0, 1, 2, 3)), (hf * hf).sum(dim=(0, 1, 2, 3))], dim=-1).double()[None]
    bn = ops.final_bn_prepare(st, n * s ** 3, train, rm, rv, nb, 0.1, update_running=train)
    out = ops.final_head_fwd(h, bn, gam, bet, w2, b2)
    np.testing.assert_allclose(out.cpu().numpy(), ref.detach().cpu().numpy(), rtol=2e-4, atol=2e-4)
    np.testing.assert_allclose(rm.cpu().numpy(), rm_ref.cpu().numpy(), rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(rv.cpu().numpy(), rv_ref.cpu().numpy(), rtol=1e-4, atol=1e-5)
    assert int(nb) == (1 if train else 0)
    dl = torch.randn_like(ref)
    ref.backward(dl)
    dh, dg, db, dw2, db2 = ops.final_head_bwd(dl, h, bn, gam, bet, w2, train)
    _close_bf16(dh, _ndhwc(ht.grad), tol=1.0 / 64)
    for got, want in ((dg, gt.grad), (db, bt.grad), (dw2, wt.grad), (db2, b2t.grad)):
        np.testing.assert_allclose(got.cpu().numpy(), want.cpu().numpy(), rtol=2e-3, atol=2e-3 * want.abs().max().item())


@pytest.mark.parametrize("cfgname", ["combined3d", "trainer", "tversky"])
def test_loss_fwd_bwd(cfgname):
    from oracle import unet3d_oracle as O
    n, s = 2, 16
    g = torch.Generator().manual_seed(21)
    logits = (torch.randn(n, 4, s, s + 2, s - 2, generator=g) * 2).to(DEV)
    target = torch.randint(0, 4, (n, s, s + 2, s - 2), generator=g).to(DEV)
    lt = logits.clone().requires_grad_(True)
    if cfgname == "combined3d":
        cfg = ops.loss_cfg(w_dice=0.5, smooth=1e-5, w_focal=0.3, f_alpha=0.25, f_gamma=2.0, w_boundary=0.2)
        ref, parts = O.combined_loss3d(lt, target)
    elif cfgname == "trainer":
        cfg = ops.loss_cfg(w_dice=0.5, smooth=1e-6, w_ce=0.3, w_focal=0.2, f_alpha=1.0, f_gamma=2.0)
        ref = O.trainer_combined_loss(lt, target)
    else:
        cfg = ops.loss_cfg(w_tv=1.0)
        ref = O.tversky_loss(lt, target)
    ref.backward()
    values, saved = ops.loss_fwd(logits, target, cfg)
    assert abs(values[0].item() - ref.item()) < 2e-5 * max(1.0, abs(ref.item()))
    if cfgname == "combined3d":
        assert abs(values[1].item() - parts["dice_loss"].item()) < 2e-5
        assert abs(values[2].item() - parts["focal_loss"].item()) < 2e-5
        assert abs(values[3].item() - parts["boundary_loss"].item()) < 2e-5
    gs = torch.full((1,), 0.5, device=DEV)
    dl = ops.loss_bwd(saved, cfg, gs, 2.0, tuple(logits.shape))  # 0.5 * 2.0 = 1
    gref = lt.grad
    assert (dl - gref).abs().max().item() <= 1e-4 * gref.abs().max().item() + 1e-9


@pytest.mark.parametrize("scale,size", [(1, (32, 32, 32)), (2, (32, 64, 32)), (4, (64, 32, 32)), (8, (32, 32, 64))])
def test_fused_deep_supervision_loss_vs_oracle_and_materialised_path(scale, size):
    """dsloss.cu: trilinear up-sampling (main.py:165-170) + CombinedLoss3D (losses.py:17-75) from LOW-RES logits, forward sums
    and d/d(low-res logits), against (a) the oracle on the F.interpolate'd map (fp32 CPU) and (b) this library's materialised
    path (trilinear_up_fwd + loss_fwd/loss_bwd + trilinear_up_bwd), which it must reproduce to fp32 rounding."""
    from oracle import unet3d_oracle as O
    n = 2
    d, h, w = size
    g = torch.Generator().manual_seed(60 + scale)
    lo = (torch.randn(n, d // scale, h // scale, w // scale, 4, generator=g) * 2).to(DEV)
    target = torch.randint(0, 4, (n, d, h, w), generator=g).to(DEV)
    cfg = ops.loss_cfg(w_dice=0.5, smooth=1e-5, w_focal=0.3, f_alpha=0.25, f_gamma=2.0, w_boundary=0.2)
    # (a) oracle
    lt = lo.cpu().clone().requires_grad_(True)
    up_ref = F.interpolate(lt.permute(0, 4, 1, 2, 3), size=size, mode="trilinear", align_corners=False)
    ref, parts = O.combined_loss3d(up_ref, target.cpu())
    (ref * 0.8).backward()
    # fused
    t8 = ops.target_u8(target)
    assert t8.dtype == torch.uint8 and torch.equal(t8.long(), target)
    values, acc = ops.dsloss_fwd(lo, t8, cfg, size)
    assert abs(values[0].item() - ref.item()) < 2e-5 * max(1.0, abs(ref.item())), (values.tolist(), ref.item())
    assert abs(values[1].item() - parts["dice_loss"].item()) < 2e-5
    assert abs(values[2].item() - parts["focal_loss"].item()) < 2e-5
    assert abs(values[3].item() - parts["boundary_loss"].item()) < 2e-5
    gs = torch.full((1,), 0.4, device=DEV)
    dlo = ops.dsloss_bwd(lo, t8, acc, cfg, gs, 2.0, size)            # 0.4 * 2.0 = 0.8
    assert dlo.shape == lo.shape and dlo.dtype == torch.float32
    gref = lt.grad
    assert (dlo.cpu() - gref).abs().max().item() <= 1e-4 * gref.abs().max().item() + 1e-9
    # (b) the materialised path of this library
    up = ops.trilinear_up_fwd(lo, size)
    v2, saved = ops.loss_fwd(up, target, cfg)
    assert abs(v2[0].item() - values[0].item()) < 1e-6 * max(1.0, abs(v2[0].item()))
    dup = ops.loss_bwd(saved, cfg, gs, 2.0, tuple(up.shape))
    if scale == 1:
        dlo2 = dup.permute(0, 2, 3, 4, 1)
    else:
        dlo2 = ops.trilinear_up_bwd(dup, (d // scale, h // scale, w // scale)).permute(0, 2, 3, 4, 1)
    assert (dlo - dlo2).abs().max().item() <= 2e-5 * dlo2.abs().max().item() + 1e-10
    # determinism: fixed-order partials + fp64 atomics
    v3, acc3 = ops.dsloss_fwd(lo, t8, cfg, size)
    assert torch.equal(v3, values)
    dlo3 = ops.dsloss_bwd(lo, t8, acc3, cfg, gs, 2.0, size)
    assert (dlo3 - dlo).abs().max().item() <= 1e-6 * dlo.abs().max().item()


def test_lazy_deep_output_fused_and_materialised_losses_agree():
    """lazy.LazyDeepOutput: DeepSupervisionLoss3D takes the fused path; the reference-style use (F.softmax / cross_entropy on the
    tensor itself) materialises it — both give the same value and the same gradient on the low-res logits."""
    import unet3d_b200 as U
    from unet3d_b200.lazy import LazyDeepOutput
    g = torch.Generator().manual_seed(70)
    size = (32, 32, 32)
    main = (torch.randn(1, 4, *size, generator=g)).to(DEV).requires_grad_(True)
    los = [(torch.randn(1, 32 // s, 32 // s, 32 // s, 4, generator=g)).to(DEV).requires_grad_(True) for s in (1, 2, 4, 8)]
    y = torch.randint(0, 4, (1,) + size, generator=g).to(DEV)
    crit = U.DeepSupervisionLoss3D()
    deep = [LazyDeepOutput(lo, size) for lo in los]
    assert all(tuple(q.shape) == (1, 4) + size and q.dtype == torch.float32 and q.is_cuda for q in deep)
    la = crit((main, deep), y)
    assert all(q._full is None for q in deep[:3]), "the fused loss must not materialise the up-sampled maps"
    la.backward()
    ga = [lo.grad.clone() for lo in los[:3]]
    assert los[3].grad is None                                   # 4th deep output unused (losses.py:118-124)
    for lo in los:
        lo.grad = None
    full = [q.materialize() for q in [LazyDeepOutput(lo, size) for lo in los]]
    assert all(type(f) is torch.Tensor and tuple(f.shape) == (1, 4) + size for f in full)
    lb = crit((main, full), y)
    lb.backward()
    assert abs(float(la) - float(lb)) <= 2e-6 * abs(float(lb))
    for a, lo in zip(ga, los[:3]):
        assert (a - lo.grad).abs().max().item() <= 2e-5 * lo.grad.abs().max().item() + 1e-10
    # reference-style consumers work on the lazy tensor directly
    q = LazyDeepOutput(los[1], size)
    ce = F.cross_entropy(q, y)
    assert abs(float(ce) - float(F.cross_entropy(full[1].detach(), y))) < 1e-6
    assert torch.equal(q.detach().argmax(1), full[1].argmax(1))


def test_confusion_and_voxel_counts_bit_exact():
    from oracle import unet3d_oracle as O
    n, s = 2, 16
    g = torch.Generator().manual_seed(22)
    logits = torch.randn(n, 4, s, s, s, generator=g)
    logits[0, :, 0, 0, :8] = 0.25  # exact ties -> lowest index, like torch.argmax
    logits[1, 2:, 1, 1, :8] = 3.0
    target = torch.randint(0, 4, (n, s, s, s), generator=g)
    hist, mask = ops.confusion(logits.to(DEV), target.to(DEV), want_mask=True)
    assert torch.equal(hist.cpu(), O.confusion_counts(logits, target))
    assert torch.equal(mask.cpu().long(), torch.argmax(logits, dim=1))
    cls, sl = ops.voxel_counts(mask[0])
    tumour, per_class, per_slice = O.voxel_counts(mask[0].cpu())
    assert cls.cpu().tolist() == per_class and sl.cpu().tolist() == per_slice and int(cls[1:].sum()) == tumour


@pytest.mark.parametrize("n,d,h,w,cin,cout,ks", [
    (2, 8, 8, 16, 32, 32, 3),      # variant 0 (kh stacked on M), row mode
    (1, 8, 16, 32, 64, 32, 3),     # variant 0 with two 32-channel blocks
    (2, 4, 8, 16, 16, 32, 3),      # variant 0, Cin=16 (G=8)
    (1, 8, 16, 16, 128, 64, 3),    # variant 1 row mode
    (2, 8, 8, 8, 256, 128, 3),     # variant 1 plane mode (W=8)
    (2, 4, 4, 4, 128, 256, 3),     # variant 1 plane mode (W=4)
    (2, 8, 8, 8, 32, 64, 3),       # small channels in plane mode
    (2, 8, 8, 16, 32, 16, 1),      # pointwise
    (1, 4, 4, 4, 256, 128, 1),     # pointwise, tiny volume
    (1, 2, 2, 2, 256, 512, 3),     # 2^3 level of a 32^3 input
    (2, 1, 1, 1, 512, 1024, 3),    # 1^3 bottleneck of a 32^3 input
    (1, 2, 2, 2, 256, 128, 1),     # pointwise with 8 voxels
    (1, 1, 1, 1, 512, 1024, 1),    # pointwise with a single voxel
    (1, 8, 32, 32, 512, 256, 1),   # wide pointwise: N block must shrink to fit shared memory
    (2, 5, 7, 24, 64, 32, 1),      # streaming pointwise kernel (conv_wgp.cu): ragged voxel count (TMA zero fill on the last tile)
    (1, 3, 16, 16, 16, 32, 1),     # conv_wgp: 16-channel blocks on X (32-byte swizzle)
    (2, 16, 16, 16, 64, 48, 1),    # conv_wgp: Cout 48 = three 16-channel blocks on M
    (1, 8, 8, 8, 1024, 512, 1),    # conv_wgp: 4 x 4 keys, several keys per CTA
    (2, 32, 32, 64, 64, 32, 1),    # conv_wgp: many K tiles per CTA (ring wrap)
])
def test_conv_wgrad(n, d, h, w, cin, cout, ks):
    x = _bf(n, d, h, w, cin, seed=31)
    dy = _bf(n, d, h, w, cout, seed=32)
    dw = ops.conv_wgrad(x, dy, cin, cout, ks)
    xt = _ncdhw(x)
    wt = torch.zeros(cout, cin, ks, ks, ks, device=DEV, requires_grad=True)
    F.conv3d(xt, wt, None, padding=ks // 2).backward(_ncdhw(dy))
    ref = wt.grad
    err = (dw - ref).abs().max().item()
    assert err <= 2e-3 * ref.abs().max().item() + 1e-3, "max err %g vs scale %g" % (err, ref.abs().max().item())


def test_conv_wgrad_padded_input_channels():
    # first encoder conv: 4 real input channels padded to 16 in the staged activation
    n, s = 2, 16
    x = _bf(n, s, s, s, 16, seed=33)
    x[..., 4:] = 0
    dy = _bf(n, s, s, s, 32, seed=34)
    dw = ops.conv_wgrad(x, dy, 4, 32, 3)
    wt = torch.zeros(32, 4, 3, 3, 3, device=DEV, requires_grad=True)
    F.conv3d(_ncdhw(x)[:, :4], wt, None, padding=1).backward(_ncdhw(dy))
    assert dw.shape == wt.shape
    assert (dw - wt.grad).abs().max().item() <= 2e-3 * wt.grad.abs().max().item() + 1e-3


@pytest.mark.parametrize("n,s,cin,cout", [(2, 4, 64, 32), (1, 8, 32, 16), (2, 16, 64, 32), (1, 1, 512, 256), (2, 2, 256, 128),
                                          (1, 32, 64, 32), (2, 8, 512, 256), (1, 4, 1024, 512)])
def test_convT2_wgrad(n, s, cin, cout):
    x = _bf(n, s, s, s, cin, seed=35)
    dy = _bf(n, 2 * s, 2 * s, 2 * s, cout, seed=36)
    dw = ops.convT2_wgrad(x, dy, cin, cout)
    wt = torch.zeros(cin, cout, 2, 2, 2, device=DEV, requires_grad=True)
    F.conv_transpose3d(_ncdhw(x), wt, None, stride=2).backward(_ncdhw(dy))
    assert (dw - wt.grad).abs().max().item() <= 2e-3 * wt.grad.abs().max().item() + 1e-3


# ------------------------------------------------------------------------------------------------------------------
# round-1 kernel set: shapes that select each conv code path (conv_zs.cu / conv_igemm.cu variants, conv_wg2.cu)
# ------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n,d,h,w,cin,cout,bias,groups", [
    (1, 8, 16, 128, 32, 32, False, 8),    # zs: COUT 32, full-width planes, several columns per CTA
    (2, 16, 8, 128, 64, 32, True, 8),     # zs: KC 64
    (1, 8, 12, 64, 16, 16, False, 16),    # zs: COUT 16 (N = 48), per-channel statistics, ragged H
    (2, 5, 9, 40, 32, 64, False, 8),      # zs: COUT 64 (ring of 8 slots), odd extents, ring wrap
    (1, 6, 16, 24, 128, 64, False, 8),    # zs: two K chunks, output-channel blocks of 16 (weights resident per block)
    (1, 7, 20, 16, 64, 128, False, 8),    # zs: 4 channel blocks of 32 -> weight reload between segments
    (2, 3, 64, 64, 16, 32, False, 8),     # zs: more CTAs than columns -> segments of 1-2 planes
    (2, 8, 16, 16, 128, 128, True, 8),    # igemm patch tiles, BN 128
    (2, 4, 4, 4, 256, 512, False, 8),     # igemm linear tiles + split-K slices
    (1, 2, 2, 2, 128, 256, False, 8),     # igemm, 8 voxels
])
def test_conv3_paths(n, d, h, w, cin, cout, bias, groups):
    x = _bf(n, d, h, w, cin, seed=41)
    wt = (_bf(cout, cin, 3, 3, 3, seed=42).float() / (cin * 27) ** 0.5).to(BF).float()
    b = torch.randn(cout, device=DEV) if bias else None
    wp, kp, rows = ops.pack_weight(wt, ops.PACK_FPROP)
    y, st = ops.conv_fprop(x, wp, rows, cout, 3, bias=b, groups=groups)
    ref = _ndhwc(F.conv3d(_ncdhw(x), wt, b, padding=1))
    _close_bf16(y, ref)
    rs = _stats(ref, groups)
    assert ((st - rs).abs() / (rs.abs() + 1.0)).max().item() < 2e-3
    # dgrad of the same layer = the same kernels on flipped / transposed packed weights
    dy = _bf(n, d, h, w, cout, seed=43)
    wd, _, rowsd = ops.pack_weight(wt, ops.PACK_DGRAD)
    dx, _ = ops.conv_fprop(dy, wd, rowsd, cin, 3)
    _close_bf16(dx, _ndhwc(F.conv_transpose3d(_ncdhw(dy), wt, padding=1)))


@pytest.mark.parametrize("n,s,cin,cout", [(1, 8, 128, 64), (2, 16, 64, 32), (1, 16, 32, 16)])
def test_convT2_fprop_multi_tap_blocks(n, s, cin, cout):
    """ConvTranspose n-blocks that span several 2x2x2 taps (pixel-shuffle target recomputed per 32-column chunk)."""
    x = _bf(n, s, s, s, cin, seed=44)
    w = (_bf(cin, cout, 2, 2, 2, seed=45).float() / cin ** 0.5).to(BF).float()
    b = torch.randn(cout, device=DEV)
    wp, kp, rows = ops.pack_weight(w, ops.PACK_CONVT_FPROP)
    y = ops.convT2_fprop(x, wp, b, cout)
    _close_bf16(y, _ndhwc(F.conv_transpose3d(_ncdhw(x), w, b, stride=2)))


@pytest.mark.parametrize("n,d,h,w,cin,cout", [
    (1, 4, 8, 32, 32, 32),      # wg2 <32,32>, one key
    (2, 5, 7, 48, 32, 32),      # ragged H, W = 3 x 16
    (2, 6, 9, 64, 32, 16),      # 16-channel dY blocks (8 descriptor groups on M)
    (1, 3, 16, 16, 16, 32),     # 16-channel X blocks (N = 48)
    (1, 8, 16, 32, 64, 64),     # 4 keys
    (1, 2, 8, 16, 128, 96),     # 12 keys, more keys than planes per CTA
    (2, 1, 4, 16, 32, 32),      # a single plane: the kd = 0 / 2 chains never start
    (1, 9, 33, 128, 16, 16),    # full-width rows, ragged H
])
def test_conv_wgrad_mn_major(n, d, h, w, cin, cout):
    x = _bf(n, d, h, w, cin, seed=46)
    dy = _bf(n, d, h, w, cout, seed=47)
    dw = ops.conv_wgrad(x, dy, cin, cout, 3)
    wt = torch.zeros(cout, cin, 3, 3, 3, device=DEV, requires_grad=True)
    F.conv3d(_ncdhw(x), wt, None, padding=1).backward(_ncdhw(dy))
    ref = wt.grad
    err = (dw - ref).abs().max().item()
    assert err <= 2e-3 * ref.abs().max().item() + 1e-3, "max err %g vs scale %g" % (err, ref.abs().max().item())


@pytest.mark.parametrize("shape,convt", [
    ((32, 32, 3, 3, 3), False), ((32, 4, 3, 3, 3), False), ((16, 32, 3, 3, 3), False), ((4, 16, 1, 1, 1), False),
    ((64, 32, 1, 1, 1), False), ((24, 40, 3, 3, 3), False), ((256, 128, 3, 3, 3), False), ((1, 8, 1, 1, 1), False),
    ((64, 32, 2, 2, 2), True), ((32, 16, 2, 2, 2), True), ((512, 256, 2, 2, 2), True), ((24, 8, 2, 2, 2), True),
])
def test_pack_weight_pair_matches_single_packs(shape, convt):
    """The one-read pair pack (what a training step runs for every conv weight) is bit-identical to the per-mode packs."""
    g = torch.Generator().manual_seed(41)
    w = torch.randn(*shape, generator=g).to(DEV)
    pair = ops.pack_weight_pair(w, convt)
    for mode, (wp, kp, rows) in pair.items():
        ref, kp_r, rows_r = ops.pack_weight(w, mode)
        assert (kp, rows) == (kp_r, rows_r)
        assert wp.shape == ref.shape and torch.equal(wp, ref), "mode %d differs" % mode


@pytest.mark.parametrize("c,s", [(32, 8), (16, 8), (256, 4), (512, 2)])
def test_groupnorm_dual_backward(c, s):
    """Dual backward of out = relu(GN_a(ya)) + GN_b(yb) (the tail of every residual block): against torch autograd and
    against the two single-branch kernels."""
    n, groups = 2, 8
    ya = _bf(n, s, s, s, c, seed=51, scale=2.0)
    yb = _bf(n, s, s, s, c, seed=52, scale=1.5)
    dout = _bf(n, s, s, s, c, seed=53)
    ga, ba = 1 + 0.2 * torch.randn(c, device=DEV), 0.2 * torch.randn(c, device=DEV)
    gb, bb = 1 + 0.2 * torch.randn(c, device=DEV), 0.2 * torch.randn(c, device=DEV)
    sta, stb = _stats(ya, groups), _stats(yb, groups)
    got = ops.gn_bwd_dual(dout, ya, sta, ga, ba, yb, stb, gb, groups)
    assert got is not None
    dxa, dga, dba, dxb, dgb, dbb = got
    yat, ybt = _ncdhw(ya).requires_grad_(True), _ncdhw(yb).requires_grad_(True)
    gat, bat, gbt, bbt = [t.clone().requires_grad_(True) for t in (ga, ba, gb, bb)]
    ref = F.relu(F.group_norm(yat, groups, gat, bat, 1e-5)) + F.group_norm(ybt, groups, gbt, bbt, 1e-5)
    ref.backward(_ncdhw(dout))
    _close_bf16(dxa, _ndhwc(yat.grad), tol=1.0 / 32)
    _close_bf16(dxb, _ndhwc(ybt.grad), tol=1.0 / 32)
    for a, b in ((dga, gat.grad), (dba, bat.grad), (dgb, gbt.grad), (dbb, bbt.grad)):
        np.testing.assert_allclose(a.cpu().numpy(), b.cpu().numpy(), rtol=2e-2, atol=2e-2 * b.abs().max().item())
    sxa, sga, sba = ops.gn_bwd(dout, ya, sta, ga, ba, groups, True)
    sxb, sgb, sbb = ops.gn_bwd(dout, yb, stb, gb, bb, groups, False)
    _close_bf16(dxa, sxa.float(), tol=1.0 / 64)
    _close_bf16(dxb, sxb.float(), tol=1.0 / 64)


@pytest.mark.parametrize("dims,cin,cout,alias", [
    ((1, 64, 32, 32), 16, 32, True),    # in-place dx += conv(dy) (attention-gate dgrad shape)
    ((1, 64, 32, 32), 32, 64, False),   # separate addend, output into a channel slice of a wider buffer
    ((2, 32, 32, 33), 32, 16, True),    # ragged voxel count
    ((1, 8, 16, 16), 16, 32, True),     # small volume
    ((1, 32, 64, 32), 64, 32, False),   # wider K
])
def test_conv1_fused_addend(dims, cin, cout, alias):
    """out = add + conv1x1(x): the branch-gradient accumulation of the backward pass (`dx += ...`)."""
    n, d, h, w = dims
    x = _bf(n, d, h, w, cin, seed=61)
    wt = (_bf(cout, cin, 1, 1, 1, seed=62).float() / cin ** 0.5).to(BF).float()
    wp, kp, rows = ops.pack_weight(wt, ops.PACK_FPROP)
    add = _bf(n, d, h, w, cout, seed=63)
    ref = _ndhwc(F.conv3d(_ncdhw(x), wt)) + add.float()
    if alias:
        out = add.clone()
        ops.conv_fprop(x, wp, rows, cout, 1, out=out, add=out)
    else:
        wide = torch.zeros(n, d, h, w, 2 * cout, dtype=BF, device=DEV)
        out = wide[..., cout:]
        ops.conv_fprop(x, wp, rows, cout, 1, out=out, add=add)
        assert float(wide[..., :cout].abs().max()) == 0.0
    _close_bf16(out, ref)


@pytest.mark.parametrize("reserved", [20, 100])
def test_results_do_not_depend_on_the_grid_width(reserved):
    """`b3d_set_reserved_sms` (data parallel: SMs left to NCCL's channel CTAs) only changes how the persistent grids split the
    work: conv fprop (z-marching and implicit-GEMM paths), weight gradients and GroupNorm agree with the full-width launch
    (same bf16 outputs up to the fp32 summation order of split-K / atomics) and with PyTorch."""
    from unet3d_b200 import _lib
    x = _bf(2, 16, 16, 32, 32, seed=71)
    w = (_bf(64, 32, 3, 3, 3, seed=72).float() / (32 * 27) ** 0.5).to(BF).float()
    wp, kp, rows = ops.pack_weight(w, ops.PACK_FPROP)
    xd = _bf(2, 4, 4, 4, 256, seed=73)
    wd = (_bf(256, 256, 3, 3, 3, seed=74).float() / (256 * 27) ** 0.5).to(BF).float()
    wdp, _, rowsd = ops.pack_weight(wd, ops.PACK_FPROP)
    dy = _bf(2, 16, 16, 32, 64, seed=75)
    gam, bet = torch.rand(64, device=DEV) + 0.5, torch.randn(64, device=DEV)

    def run():
        y, st = ops.conv_fprop(x, wp, rows, 64, 3, groups=8)
        yd, _ = ops.conv_fprop(xd, wdp, rowsd, 256, 3, groups=8)
        dw = ops.conv_wgrad(x, dy, 32, 64, 3)
        a = ops.gn_apply(y, st, gam, bet, 8, True)
        ops.wgrad_join()
        torch.cuda.synchronize()
        return y, st, yd, dw, a

    full = run()
    old = _lib.set_reserved_sms(reserved)
    try:
        part = run()
    finally:
        _lib.set_reserved_sms(old)
    ref = _ndhwc(F.conv3d(_ncdhw(x), w, None, padding=1))
    _close_bf16(part[0], ref)
    for a, b in zip(full, part):
        scale = max(a.float().abs().max().item(), 1e-6)
        assert (a.float() - b.float()).abs().max().item() <= scale / 128, "grid width changed a result"
