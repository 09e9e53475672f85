"""End-to-end parity of the B200 path (through the reference-shaped nn.Module API) against the oracle and the golden
vectors produced by the reference itself.

Stated tolerances (bf16 activations/weights, fp32 accumulation — north_star "bf16 relative tolerance"):
  * logits          : relative L2 error <= 2.5e-2 vs the fp32 reference logits (reference's own bf16-autocast error: 1.25e-2)
  * loss            : |delta| <= 1.5e-2 * |loss|
  * gradients       : per-tensor cosine similarity >= 0.9 and relative L2 <= 0.35 for tensors carrying >= 1e-3 of the total
                      gradient norm (the reference's own bf16-autocast gradients sit at median 0.145 / p90 0.29 rel-L2,
                      SURVEY hard part 5); whole-model gradient cosine >= 0.99
  * argmax / counts : bit-exact GIVEN IDENTICAL fp32 logits (kernel-level test); end-to-end agreement fraction >= 0.97
"""
import json
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    import b3d  # noqa: F401
    import unet3d_b200 as U
    from unet3d_b200 import ops

from oracle import unet3d_oracle as O

from parity_util import DEV, REPORT, rel_l2 as _rel_l2, cos as _cos, check_grads as _check_grads  # noqa: E402
from parity_util import oracle_autocast_grads as _oracle_autocast_grads, oracle_train as _oracle_train  # noqa: E402


def _load(model, sd):
    model.load_state_dict(sd)
    return model.to(DEV)


@pytest.mark.parametrize("case", ["model_small", "model_small_n2"])
def test_unet_eval_and_train_vs_golden_and_oracle(golden, golden_arrays, case):
    rec = golden[case]
    feats = tuple(rec["features"])
    sd = O.make_state_dict(4, 4, feats, seed=rec["seed"])
    x, y = O.make_inputs(rec["n"], rec["size"], rec["size"], rec["size"], seed=rec["seed"])
    model = _load(U.UNet3D(4, 4, features=list(feats), dropout_rate=0.0), sd)
    xd, yd = x.to(DEV), y.to(DEV)
    model.eval()
    with torch.no_grad():
        ev = model(xd)
    assert ev.shape == (rec["n"], 4) + (rec["size"],) * 3 and ev.dtype == torch.float32
    with torch.no_grad():
        ev_ref, _, _ = O.unet_forward(x, sd, feats, training=False)
    r = _rel_l2(ev.cpu(), ev_ref)
    REPORT[case + "_eval_logits_rel_l2"] = r
    assert r <= 2.5e-2, r
    arrays = golden_arrays(case)
    if "eval_logits" in arrays:  # the reference's own output
        assert _rel_l2(ev.cpu(), torch.from_numpy(arrays["eval_logits"])) <= 2.5e-2
    agree = float((ev.argmax(1).cpu() == ev_ref.argmax(1)).float().mean())
    REPORT[case + "_argmax_agree"] = agree
    assert agree >= 0.97, agree
    assert abs(U.calculate_dice_score(ev, yd) - O.dice_score(ev.cpu(), y)) < 1e-7  # metric exact on identical logits
    # ---- training step
    model.train()
    main, deep = model(xd)
    assert isinstance(deep, list) and len(deep) == 4 and all(d.shape == main.shape for d in deep)
    crit = U.DeepSupervisionLoss3D()
    loss = crit((main, deep), yd)
    loss.backward()
    rmain, rdeep, rloss, rgrads, rbn = _oracle_train(sd, x, y, feats)
    assert abs(rloss - rec["ds_loss"]) < 1e-4  # oracle == reference
    REPORT[case + "_loss"] = (float(loss), rloss)
    assert abs(float(loss) - rloss) <= 1.5e-2 * abs(rloss), (float(loss), rloss)
    assert _rel_l2(main.detach().cpu(), rmain) <= 2.5e-2
    for i in range(4):
        assert _rel_l2(deep[i].detach().cpu(), rdeep[i]) <= 2.5e-2, i
    _check_grads(model, rgrads, case, _oracle_autocast_grads(sd, x, y, feats))
    np.testing.assert_allclose(model.final_conv[1].running_mean.cpu().numpy(), rbn[0].detach().numpy(), rtol=2e-2, atol=2e-3)
    np.testing.assert_allclose(model.final_conv[1].running_var.cpu().numpy(), rbn[1].detach().numpy(), rtol=2e-2, atol=2e-3)
    assert int(model.final_conv[1].num_batches_tracked) == 1
    print("REPORT", json.dumps({k: v for k, v in REPORT.items()}, default=str))


def test_unet_dropout_uses_torch_rng_stream():
    """Dropout3d parity: (a) the masks the model draws are the ones F.dropout3d would draw from the same generator state
    (same op sequence on the torch CUDA generator), (b) with those masks the step matches the oracle."""
    import torch.nn.functional as F
    feats = (16, 32, 64, 128, 256)
    sd = O.make_state_dict(4, 4, feats, seed=3)
    x, y = O.make_inputs(2, 32, 32, 32, seed=3)
    model = _load(U.UNet3D(4, 4, features=list(feats), dropout_rate=0.2), sd)
    model.train()
    torch.manual_seed(99)
    ref_masks = [F.dropout3d(torch.ones(2, f, 2, 2, 2, device=DEV), 0.2, True)[:, :, 0, 0, 0].cpu() for f in feats]
    torch.manual_seed(99)
    main, deep = model(x.to(DEV))
    masks = [m.cpu() for m in model._last_dropout_masks]
    for a, b in zip(masks, ref_masks):
        assert torch.equal(a, b)
        assert set(a.unique().tolist()) <= {0.0, 1.25}
    loss = U.DeepSupervisionLoss3D()((main, deep), y.to(DEV))
    loss.backward()
    rmain, rdeep, rloss, rgrads, _ = _oracle_train(sd, x, y, feats, masks=masks)
    assert _rel_l2(main.detach().cpu(), rmain) <= 2.5e-2
    assert abs(float(loss) - rloss) <= 1.5e-2 * abs(rloss)
    _check_grads(model, rgrads, "dropout", _oracle_autocast_grads(sd, x, y, feats, masks=masks))


def test_default_architecture_train_step(golden):
    rec = golden["model_default"]
    feats = tuple(rec["features"])
    sd = O.make_state_dict(4, 4, feats, seed=rec["seed"])
    x, y = O.make_inputs(rec["n"], rec["size"], rec["size"], rec["size"], seed=rec["seed"])
    model = _load(U.UNet3D(4, 4, features=list(feats), dropout_rate=0.0), sd)
    model.train()
    main, deep = model(x.to(DEV))
    loss = U.DeepSupervisionLoss3D()((main, deep), y.to(DEV))
    loss.backward()
    assert abs(float(loss) - rec["ds_loss"]) <= 1.5e-2 * abs(rec["ds_loss"]), (float(loss), rec["ds_loss"])
    for k, p in model.named_parameters():
        g = rec["grads"][k]
        if g is None:
            assert p.grad is None or float(p.grad.abs().max()) == 0.0
        elif g["l2"] > 1e-3:
            got = float(p.grad.double().norm())
            assert abs(got - g["l2"]) <= 0.35 * g["l2"], (k, got, g["l2"])


def test_loss_modules_match_oracle():
    g = torch.Generator().manual_seed(5)
    logits = torch.randn(2, 4, 16, 16, 16, generator=g) * 2
    y = torch.randint(0, 4, (2, 16, 16, 16), generator=g)
    ld_, yd = logits.to(DEV).requires_grad_(True), y.to(DEV)
    tot, parts = U.CombinedLoss3D()(ld_, yd)
    rt, rp = O.combined_loss3d(logits, y)
    assert abs(float(tot) - float(rt)) < 2e-5
    assert set(parts) == {"dice_loss", "focal_loss", "boundary_loss", "total_loss"}
    for k in parts:
        assert isinstance(parts[k], float) and abs(parts[k] - float(rp[k])) < 2e-5
    assert abs(float(U.CombinedLoss()(ld_, yd)) - float(O.trainer_combined_loss(logits, y))) < 2e-5
    assert abs(float(U.TverskyLoss3D()(ld_, yd)) - float(O.tversky_loss(logits, y))) < 2e-5
    assert abs(float(U.DiceLoss()(ld_, yd)) - float(O.dice_loss(logits, y, 1e-6))) < 2e-5
    assert abs(float(U.FocalLoss()(ld_, yd)) - float(O.focal_loss(logits, y, 1.0, 2.0))) < 2e-5
    # tensor (eval-mode output) through the deep-supervision wrapper
    assert abs(float(U.DeepSupervisionLoss3D()(ld_, yd)) - float(rt)) < 2e-5
    lt = logits.clone().requires_grad_(True)
    (O.trainer_combined_loss(lt, y) * 3.0).backward()
    (U.CombinedLoss()(ld_, yd) * 3.0).backward()
    assert (ld_.grad.cpu() - lt.grad).abs().max() <= 1e-4 * lt.grad.abs().max()


def test_standalone_blocks_vs_reference_golden(golden_arrays):
    arrays = golden_arrays("blocks")
    sd = O.make_state_dict(16, 4, (32, 64, 128, 256, 512), seed=5)
    g = torch.Generator().manual_seed(9)
    x = torch.randn(2, 16, 8, 8, 8, generator=g)
    dc = U.DoubleConv3D(16, 32)
    dc.load_state_dict({k[len("downs.0."):]: v for k, v in sd.items() if k.startswith("downs.0.")})
    dc = dc.to(DEV)
    xd = x.to(DEV).requires_grad_(True)
    out = dc(xd)
    assert _rel_l2(out.detach().cpu(), torch.from_numpy(arrays["doubleconv_out"])) <= 2.5e-2
    out.square().mean().backward()
    xr = x.clone().requires_grad_(True)
    sdo = {k: v.clone().requires_grad_(True) for k, v in sd.items() if k.startswith("downs.0.")}
    O.double_conv(xr, sdo, "downs.0.").square().mean().backward()
    assert _cos(xd.grad.cpu(), xr.grad) >= 0.98
    for k, p in dc.named_parameters():
        assert _cos(p.grad.cpu(), sdo["downs.0." + k].grad) >= 0.95, k
    gg = torch.randn(2, 32, 8, 8, 8, generator=g)
    xx = torch.randn(2, 32, 8, 8, 8, generator=g)
    ag = U.AttentionGate3D(32, 32, 16)
    ag.load_state_dict({k[len("ups.13."):]: v for k, v in sd.items() if k.startswith("ups.13.")})
    ag = ag.to(DEV)
    gd, xd2 = gg.to(DEV).requires_grad_(True), xx.to(DEV).requires_grad_(True)
    out = ag(g=gd, x=xd2)
    assert _rel_l2(out.detach().cpu(), torch.from_numpy(arrays["gate_out"])) <= 2.5e-2
    out.square().mean().backward()
    gr, xr2 = gg.clone().requires_grad_(True), xx.clone().requires_grad_(True)
    sdo = {k: v.clone().requires_grad_(True) for k, v in sd.items() if k.startswith("ups.13.")}
    O.attention_gate(gr, xr2, sdo, "ups.13.").square().mean().backward()
    assert _cos(xd2.grad.cpu(), xr2.grad) >= 0.98 and _cos(gd.grad.cpu(), gr.grad) >= 0.95
    for k, p in ag.named_parameters():
        rg = sdo["ups.13." + k].grad
        if float(rg.abs().max()) < 1e-9:
            assert float(p.grad.abs().max()) < 1e-5, k  # psi.0.bias: analytically zero (cancels in GroupNorm(1,1))
        else:
            assert _cos(p.grad.cpu(), rg) >= 0.95, (k, _cos(p.grad.cpu(), rg))


def test_no_cpu_fallback():
    m = U.UNet3D(4, 4, features=[16, 32, 64, 128, 256])
    with pytest.raises(Exception):
        m(torch.randn(1, 4, 32, 32, 32))
    with pytest.raises(Exception):
        U.CombinedLoss3D()(torch.randn(1, 4, 8, 8, 8), torch.zeros(1, 8, 8, 8, dtype=torch.long))
    with pytest.raises(ValueError):
        m.to(DEV)(torch.randn(1, 4, 24, 32, 32, device=DEV))


def test_segment_and_volumes_bit_exact():
    feats = (16, 32, 64, 128, 256)
    sd = O.make_state_dict(4, 4, feats, seed=7)
    x, _ = O.make_inputs(1, 32, 32, 32, seed=7)
    model = _load(U.UNet3D(4, 4, features=list(feats)), sd).eval()
    mask, logits = U.segment(model, x.to(DEV), return_logits=True)
    assert mask.dtype == torch.uint8 and torch.equal(mask.long().cpu(), logits.argmax(1).cpu())
    vols = U.tumor_volumes(mask[0])
    tumour, per_class, per_slice = O.voxel_counts(mask[0].cpu())
    assert vols == {"tumor_voxels": tumour, "class_voxels": per_class, "slice_voxels": per_slice}


def test_graphed_train_step_matches_eager():
    """GraphedTrainStep (zero_grad + forward + DS loss + backward + AdamW as one CUDA graph) follows the eager step."""
    feats = (16, 32, 64, 128, 256)
    sd = O.make_state_dict(4, 4, feats, seed=5)
    x, y = O.make_inputs(2, 32, 32, 32, seed=5)
    xd, yd = x.to(DEV), y.to(DEV)
    losses = {}
    for mode in ("eager", "graph"):
        model = _load(U.UNet3D(4, 4, features=list(feats), dropout_rate=0.0), sd).train()
        crit = U.DeepSupervisionLoss3D()
        opt = torch.optim.AdamW(model.parameters(), lr=1e-3, weight_decay=1e-4, fused=True, capturable=True)
        out = []
        if mode == "graph":
            # the capture protocol runs 3 warm-up optimizer steps: undo them so that both runs start from the same weights
            step = U.GraphedTrainStep(model, crit, opt, xd, yd, warmup=1)
            model.load_state_dict(sd)
            for st in opt.state.values():      # the graph captured the state tensors by address: reset them in place
                for v in st.values():
                    if torch.is_tensor(v):
                        v.zero_()
            for _ in range(3):
                out.append(float(step(xd, yd)))
        else:
            for _ in range(3):
                opt.zero_grad(set_to_none=True)
                loss = crit(model(xd), yd)
                loss.backward()
                opt.step()
                out.append(float(loss))
        losses[mode] = out
    e, g = losses["eager"], losses["graph"]
    assert e[2] < e[0] and g[2] < g[0], (e, g)                    # both actually train
    for a, b in zip(e, g):                                        # same kernels in the same order: only fp32-atomic jitter
        assert abs(a - b) <= 2e-4 * abs(a), (e, g)


def test_optimizer_step_invalidates_packed_weights():
    """torch's fused AdamW updates parameters without bumping their version counters; the bf16 packed copies the kernels
    read must follow anyway (global optimizer post-step hook -> cache epoch).  After a step the model must compute exactly
    what a freshly constructed model with the updated state_dict computes."""
    feats = (16, 32, 64, 128, 256)
    sd = O.make_state_dict(4, 4, feats, seed=9)
    x, y = O.make_inputs(1, 32, 32, 32, seed=9)
    xd, yd = x.to(DEV), y.to(DEV)
    model = _load(U.UNet3D(4, 4, features=list(feats), dropout_rate=0.0), sd).train()
    crit = U.DeepSupervisionLoss3D()
    opt = torch.optim.AdamW(model.parameters(), lr=1e-2, weight_decay=1e-4, fused=True)
    before = model(xd)[0].detach().clone()
    opt.zero_grad(set_to_none=True)
    crit(model(xd), yd).backward()
    opt.step()
    after = model(xd)[0].detach().clone()
    fresh = _load(U.UNet3D(4, 4, features=list(feats), dropout_rate=0.0), {k: v.detach().clone() for k, v in model.state_dict().items()}).train()
    # BatchNorm running statistics were advanced by the forwards above; train mode normalises with batch statistics
    want = fresh(xd)[0].detach()
    assert not torch.equal(before, after), "the optimizer step did not change the forward pass"
    assert torch.equal(after, want), "forward after opt.step() differs from a fresh model with the same parameters (stale packed weights): max |d| = %g" % float((after - want).abs().max())


def test_packed_weight_cache_is_per_parameter():
    """Regression: the packed-weight cache used to be keyed by id(param); a second model that re-used the id, version and
    address of a freed parameter silently ran with the first model's weights."""
    import gc
    from unet3d_b200 import functional
    feats = (16, 32, 64, 128, 256)
    x, _ = O.make_inputs(1, 32, 32, 32, seed=7)
    xd = x.to(DEV)
    outs = []
    for seed in (1, 2, 1, 2):
        model = _load(U.UNet3D(4, 4, features=list(feats), dropout_rate=0.0), O.make_state_dict(4, 4, feats, seed=seed)).eval()
        with torch.no_grad():
            outs.append(model(xd).clone())
        del model
        gc.collect()
    assert _rel_l2(outs[2], outs[0]) < 3e-2 and _rel_l2(outs[3], outs[1]) < 3e-2     # same weights -> same logits
    assert _rel_l2(outs[1], outs[0]) > 0.3                                            # different weights -> different logits
    functional.clear_pack_cache()


def test_unet_width_not_multiple_of_16_vs_oracle():
    """W = 96 gives levels of width 48, 24, 12, 6, 3: the weight-gradient kernels' zero-padding path (ops._pad_w16) and the
    ragged tiles of every conv kernel (BASELINE config 5 runs 160x192x160 volumes)."""
    feats = (16, 32, 64, 128, 256)
    sd = O.make_state_dict(4, 4, feats, seed=9)
    g = torch.Generator().manual_seed(9)
    x = torch.randn(1, 4, 32, 64, 96, generator=g)
    y = torch.randint(0, 4, (1, 32, 64, 96), generator=g)
    model = _load(U.UNet3D(4, 4, features=list(feats), dropout_rate=0.0), sd).train()
    main, deep = model(x.to(DEV))
    loss = U.DeepSupervisionLoss3D()((main, deep), y.to(DEV))
    loss.backward()
    rmain, rdeep, rloss, rgrads, _ = _oracle_train(sd, x, y, feats)
    assert _rel_l2(main.detach().cpu(), rmain) <= 2.5e-2
    assert abs(float(loss) - rloss) <= 1.5e-2 * abs(rloss)
    _check_grads(model, rgrads, "ragged_w", _oracle_autocast_grads(sd, x, y, feats))


def test_run_to_run_reproducibility():
    """Two identical train passes IN ORDERED-ISSUE MODE (b3d_set_ordered_issue(1): one MMA issuer in the z-marching conv kernel;
    the default two ping-pong issuers accumulate in an order that jitters by an fp32 ulp, which flips ~1e-7 of the bf16 outputs
    — scripts/zs_jitter.py): the forward (logits, deep-supervision maps, loss) is bit-identical and every parameter
    gradient agrees to fp32 rounding.  The block-level reductions that feed activations or input gradients (conv-epilogue
    GroupNorm/BatchNorm sums, GroupNorm-backward sums, gate and loss sums) accumulate fixed-order fp32 partials in fp64, so
    the arrival order of warps/CTAs cannot flip a bf16 rounding downstream; only the weight-gradient flush (leaves of the
    backward graph) still uses fp32 atomics.  Before this, a single 1-ulp flip at a 4^3 level grew to 15 % run-to-run
    differences in deep-level weight gradients (scripts/diverge_check.py, scripts/jitter_check.py)."""
    feats = (16, 32, 64, 128, 256)
    sd = O.make_state_dict(4, 4, feats, seed=11)
    x, y = O.make_inputs(2, 32, 32, 32, seed=11)
    xd, yd = x.to(DEV), y.to(DEV)
    model = _load(U.UNet3D(4, 4, features=list(feats), dropout_rate=0.0), sd).train()
    crit = U.DeepSupervisionLoss3D()
    runs = []
    from unet3d_b200 import _lib
    prev = _lib.set_ordered_issue(2)
    try:
        for _ in range(3):
            model.zero_grad(set_to_none=True)
            main, deep = model(xd)
            loss = crit((main, deep), yd)
            loss.backward()
            torch.cuda.synchronize()
            runs.append((main.detach().clone(), [d.detach().clone() for d in deep], loss.detach().clone(),
                         {k: p.grad.detach().clone() for k, p in model.named_parameters() if p.grad is not None}))
        model.eval()
        with torch.no_grad():
            e1, e2 = model(xd), model(xd)
        assert torch.equal(e1, e2), "inference logits differ between two identical runs"
        model.train()
    finally:
        _lib.set_ordered_issue(prev)
    a = runs[0]
    total = float(torch.sqrt(sum((g.double() ** 2).sum() for g in a[3].values())))
    for b in runs[1:]:
        assert torch.equal(a[0], b[0]), "main logits differ between two identical runs"
        assert all(torch.equal(p, q) for p, q in zip(a[1], b[1])), "deep-supervision outputs differ between identical runs"
        assert torch.equal(a[2], b[2]), "loss differs between identical runs"
        for k in a[3]:
            d = float((a[3][k] - b[3][k]).norm()) / max(float(a[3][k].norm()), 1e-3 * total)
            assert d <= 2e-5, "gradient of %s differs by %.3g between identical runs" % (k, d)
    # default mode (two ping-pong issuers): same result up to isolated 1-ulp bf16 flips amplified by the deep levels
    main2, _ = model(xd)
    assert _rel_l2(main2.detach(), a[0]) <= 2.5e-2


@pytest.mark.parametrize("fused", [True, False])
def test_training_trajectory_follows_the_oracle(fused):
    """Five AdamW steps on one batch: the loss trajectory of the B200 path follows the fp32 CPU oracle's (same weights, same
    data, same optimizer).  This is the end-to-end guard for everything a single fwd/bwd parity test cannot see — above all
    that the kernels really read the UPDATED weights after every optimizer step (fused AdamW included)."""
    feats = (16, 32, 64, 128, 256)
    sd = O.make_state_dict(4, 4, feats, seed=13)
    x, y = O.make_inputs(2, 32, 32, 32, seed=13)
    steps, lr = 5, 2e-3
    # oracle: functional fp32 forward on CPU, torch autograd, torch AdamW
    params = {k: v.clone().requires_grad_(True) for k, v in sd.items() if v.is_floating_point() and "running" not in k}
    osd = {k: v.clone() for k, v in sd.items()}
    osd.update(params)
    oopt = torch.optim.AdamW(list(params.values()), lr=lr, weight_decay=1e-4)
    ref = []
    for _ in range(steps):
        oopt.zero_grad()
        main, deep, _ = O.unet_forward(x, osd, feats, training=True, dropout_masks=None)
        loss = O.deep_supervision_loss(main, deep, y)
        loss.backward()
        oopt.step()
        ref.append(float(loss))
    model = _load(U.UNet3D(4, 4, features=list(feats), dropout_rate=0.0), sd).train()
    crit = U.DeepSupervisionLoss3D()
    opt = torch.optim.AdamW(model.parameters(), lr=lr, weight_decay=1e-4, fused=fused)
    xd, yd = x.to(DEV), y.to(DEV)
    got = []
    for _ in range(steps):
        opt.zero_grad(set_to_none=True)
        loss = crit(model(xd), yd)
        loss.backward()
        opt.step()
        got.append(float(loss))
    assert ref[-1] < 0.97 * ref[0], ref                         # the problem does train at this learning rate
    for i, (a, b) in enumerate(zip(ref, got)):                  # bf16 activations: 1.5 % per step, like the single-step bar
        assert abs(a - b) <= 1.5e-2 * abs(a), "step %d: oracle %s vs b200 %s" % (i, ref, got)
    assert (ref[0] - got[-1]) >= 0.7 * (ref[0] - ref[-1]), "b200 path trains slower than the oracle: %s vs %s" % (got, ref)


def test_single_channel_input_vs_oracle():
    """The reference's own constructor default and inference call site build `UNet3D(in_channels=1)` (main.py:105,336,
    web_training.py:67): one real input channel padded to 16 in the staged activation, weight gradient of the first conv
    cut back to one channel."""
    feats = (16, 32, 64, 128, 256)
    sd = O.make_state_dict(1, 4, feats, seed=17)
    x, y = O.make_inputs(2, 32, 32, 32, in_channels=1, seed=17)
    model = _load(U.UNet3D(in_channels=1, out_channels=4, features=list(feats), dropout_rate=0.0), sd)
    model.eval()
    with torch.no_grad():
        ev = model(x.to(DEV))
    rev, _, _ = O.unet_forward(x, sd, feats, training=False)
    assert ev.shape == rev.shape and _rel_l2(ev.cpu(), rev) <= 2.5e-2
    model.train()
    main, deep = model(x.to(DEV))
    loss = U.DeepSupervisionLoss3D()((main, deep), y.to(DEV))
    loss.backward()
    rmain, rdeep, rloss, rgrads, _ = _oracle_train(sd, x, y, feats)
    assert _rel_l2(main.detach().cpu(), rmain) <= 2.5e-2
    assert abs(float(loss) - rloss) <= 1.5e-2 * abs(rloss)
    assert model.downs[0].double_conv[0].weight.grad.shape == (16, 1, 3, 3, 3)
    _check_grads(model, rgrads, "cin1", _oracle_autocast_grads(sd, x, y, feats))


@pytest.mark.parametrize("name,n,size,seed", [("a", 2, (32, 32, 32), 21), ("b", 1, (16, 32, 48), 22), ("c", 1, (64, 64, 64), 23)])
def test_classifier_vs_reference_golden_and_oracle(name, n, size, seed):
    """BrainTumorClassifier (main.py:301-328): eval forward against the reference's own outputs
    (tests/golden/classifier.npz) and the oracle; bf16 convolutions -> 2.5e-2 relative on the logits, same predicted class
    whenever the reference's top-2 margin exceeds that error."""
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "classifier.npz"))
    sd = O.make_classifier_state_dict(4, seed=seed)
    x, _ = O.make_inputs(n, *size, seed=seed)
    model = U.BrainTumorClassifier(4)
    assert [(k, tuple(v.shape)) for k, v in model.state_dict().items()] == O.classifier_param_shapes(4)
    model.load_state_dict(sd)
    model = model.to(DEV).eval()
    with torch.no_grad():
        got = model(x.to(DEV)).cpu()
    ref = torch.from_numpy(z["logits_" + name])
    orc = O.classifier_forward(x, sd)
    assert got.shape == ref.shape == orc.shape
    scale = float(ref.abs().max())
    assert float((got - ref).abs().max()) <= 2.5e-2 * scale, (got, ref)
    assert float((got - orc).abs().max()) <= 2.5e-2 * scale
    top2 = ref.topk(2, dim=1).values
    clear = (top2[:, 0] - top2[:, 1]) > 5e-2 * scale
    assert torch.equal(got.argmax(1)[clear], ref.argmax(1)[clear])
    with pytest.raises(NotImplementedError):
        model.train()(x.to(DEV))
