"""The reference trainer's LITERAL call sequence on the B200 path:

    /root/reference/training.py:290-309
        optimizer.zero_grad()
        with torch.cuda.amp.autocast():
            outputs = model(images); loss = criterion(outputs, masks)
        scaler.scale(loss).backward(); scaler.step(optimizer); scaler.update()
        dice = calculate_dice_score(outputs, masks)

plus the graph wrappers the bench's numbers come from: GraphedInference == eager bit-for-bit (and survives pack-cache churn),
GraphedTrainStep.prefetch/step_prefetched == __call__, a scheduler step changes the update under replay, and the NCCL
data-parallel equivalence check (2 GPUs, skipped on a 1-GPU box).
"""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    import b3d  # noqa: F401
    import unet3d_b200 as U
    from unet3d_b200 import functional

from oracle import unet3d_oracle as O
from parity_util import DEV, exact_fp32

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FEATS = (16, 32, 64, 128, 256)


def _model(sd, dropout=0.0, feats=FEATS):
    m = U.UNet3D(4, 4, features=list(feats), dropout_rate=dropout)
    m.load_state_dict(sd)
    return m.to(DEV)


def test_reference_trainer_loop_autocast_gradscaler_adamw_dice():
    """3 iterations of training.py:290-309 verbatim (autocast + GradScaler + AdamW(lr, wd=1e-4) + calculate_dice_score) against
    the fp32 oracle running the same loop without a scaler (loss scaling is a mathematical no-op when nothing overflows)."""
    sd = O.make_state_dict(4, 4, FEATS, seed=41)
    x, y = O.make_inputs(2, 32, 32, 32, seed=41)
    lr, steps = 2e-3, 3
    # ---- oracle: fp32, plain AdamW, Dice from its own logits
    params = {k: v.clone().requires_grad_(True) for k, v in sd.items() if v.is_floating_point() and "running" not in k}
    osd = {k: v.clone() for k, v in sd.items()}
    osd.update(params)
    oopt = torch.optim.AdamW(list(params.values()), lr=lr, weight_decay=1e-4, betas=(0.9, 0.999))
    ref_loss, ref_dice = [], []
    for _ in range(steps):
        oopt.zero_grad()
        main, deep, _ = O.unet_forward(x, osd, FEATS, training=True)
        loss = O.deep_supervision_loss(main, deep, y)
        loss.backward()
        oopt.step()
        ref_loss.append(float(loss.detach()))
        ref_dice.append(O.dice_score(main.detach(), y))
    # ---- B200 path: the trainer's literal sequence
    model = _model(sd).train()
    criterion = U.DeepSupervisionLoss3D()
    optimizer = torch.optim.AdamW(model.parameters(), lr=lr, weight_decay=1e-4, betas=(0.9, 0.999))   # training.py:186-191
    scaler = torch.cuda.amp.GradScaler()                                                             # training.py:199
    images, masks = x.to(DEV), y.to(DEV)
    got_loss, got_dice = [], []
    for _ in range(steps):
        optimizer.zero_grad()
        with torch.cuda.amp.autocast():
            outputs = model(images)
            loss = criterion(outputs, masks)
        scaler.scale(loss).backward()
        scaler.step(optimizer)
        scaler.update()
        dice = U.calculate_dice_score(outputs[0], masks)
        got_loss.append(loss.item())
        got_dice.append(dice)
        assert isinstance(dice, float) and 0.0 <= dice <= 1.0
        # the metric kernel is integer-exact on the logits it was given
        assert abs(dice - O.dice_score(outputs[0].detach().cpu(), y)) < 1e-7
    assert scaler.get_scale() == 65536.0, "GradScaler saw an inf/nan and backed off: %s" % scaler.get_scale()
    for i, (a, b) in enumerate(zip(ref_loss, got_loss)):
        assert abs(a - b) <= 1.5e-2 * abs(a), "step %d: oracle %s vs b200 %s" % (i, ref_loss, got_loss)
    assert got_loss[-1] < got_loss[0]
    for a, b in zip(ref_dice, got_dice):   # Dice of near-random logits: flips of near-ties move it by a few 1e-3
        assert abs(a - b) <= 2e-2, (ref_dice, got_dice)


def test_reference_loss_and_metric_classes_consume_our_model_outputs():
    """Drop-in both ways: the REFERENCE's own `DeepSupervisionLoss3D(CombinedLoss3D)` (losses.py:7-126) and
    `calculate_dice_score` (training.py:351-364), executed unmodified (oracle/_ref), take this model's train-mode output — the lazy
    deep-supervision tensors materialise through `__torch_function__` — and give the loss / gradients of our fused loss path."""
    from oracle import ref_slice
    if not ref_slice.available():
        pytest.skip("reference classes not materialised (oracle/make_ref.py needs /root/reference)")
    ns = ref_slice.load()
    sd = O.make_state_dict(4, 4, FEATS, seed=49)
    x, y = O.make_inputs(2, 32, 32, 32, seed=49)
    xd, yd = x.to(DEV), y.to(DEV)
    from unet3d_b200 import _lib
    prev = _lib.set_ordered_issue(2)      # both passes must see the same forward
    try:
        grads = {}
        for which in ("ours", "reference"):
            model = _model(sd).train()
            out = model(xd)
            crit = U.DeepSupervisionLoss3D() if which == "ours" else ns["DeepSupervisionLoss3D"]()
            loss = crit(out, yd)
            loss.backward()
            grads[which] = (float(loss.detach()), {k: p.grad.detach().clone() for k, p in model.named_parameters() if p.grad is not None})
            if which == "reference":
                assert all(q._full is not None for q in out[1][:3]), "the reference loss must have materialised the deep outputs"
                d_ref = ns["calculate_dice_score"](None, out[0].detach(), yd)
                assert abs(d_ref - U.calculate_dice_score(out[0].detach(), yd)) < 1e-6
    finally:
        _lib.set_ordered_issue(prev)
    (la, ga), (lb, gb) = grads["ours"], grads["reference"]
    assert abs(la - lb) <= 2e-5 * abs(lb), (la, lb)
    assert set(ga) == set(gb)
    tot = sum(float(v.double().norm()) ** 2 for v in gb.values()) ** 0.5
    worst = 0.0
    for k in gb:   # same forward, two implementations of the loss + its backward.  The logit gradients agree to fp32 rounding; the
        # bf16 backward amplifies that at the 1^3 / 2^3 levels (measured worst tensor: ups.0.weight 0.9 %), far below the 10-30 %
        # bf16-vs-fp32 error bar of those tensors
        d = float((ga[k] - gb[k]).norm()) / max(float(gb[k].norm()), 1e-3 * tot)
        worst = max(worst, d)
        assert d <= 3e-2, (k, d)
    dots = sum(float(ga[k].double().reshape(-1) @ gb[k].double().reshape(-1)) for k in gb)
    na = sum(float(ga[k].double().norm()) ** 2 for k in gb) ** 0.5
    assert dots / (na * tot) >= 0.9999, dots / (na * tot)


def test_trainer_combined_loss_on_eval_output_under_autocast():
    """validate_epoch (training.py:332-339): eval forward -> training.CombinedLoss -> calculate_dice_score, under no_grad."""
    sd = O.make_state_dict(4, 4, FEATS, seed=42)
    x, y = O.make_inputs(1, 32, 32, 32, seed=42)
    model = _model(sd).eval()
    with torch.no_grad():
        out = model(x.to(DEV))
        loss = U.CombinedLoss()(out, y.to(DEV))
        ref = O.trainer_combined_loss(out.cpu(), y)
    assert abs(loss.item() - float(ref)) < 2e-5
    assert abs(U.calculate_dice_score(out, y.to(DEV)) - O.dice_score(out.cpu(), y)) < 1e-7


def test_graphed_inference_is_bitwise_eager_and_owns_its_weights():
    """the captured kernels point at packed bf16 weights; a later cache miss must not free them."""
    sd = O.make_state_dict(4, 4, FEATS, seed=43)
    x, _ = O.make_inputs(1, 32, 32, 32, seed=43)
    x2, _ = O.make_inputs(1, 32, 32, 32, seed=44)
    model = _model(sd).eval()
    xd, x2d = x.to(DEV), x2.to(DEV)
    with torch.no_grad():
        want, want2 = model(xd).clone(), model(x2d).clone()
    infer = U.GraphedInference(model, xd)
    assert torch.equal(infer(xd), want)
    assert torch.equal(infer(x2d), want2)
    # churn: invalidate the pack cache, run eager forwards (re-pack -> the cache entries are REPLACED), a second graph on
    # another volume size, and enough allocations to recycle any freed block
    functional.clear_pack_cache()
    with torch.no_grad():
        model(x2d)
    big, _ = O.make_inputs(1, 64, 32, 32, seed=45)
    infer_b = U.GraphedInference(model, big.to(DEV))
    junk = [torch.randn(1 << 20, device=DEV) for _ in range(64)]
    junk = [torch.full((n,), 7.0, device=DEV, dtype=torch.bfloat16) for n in (1 << 12, 1 << 16, 1 << 20, 1 << 22) for _ in range(8)]
    torch.cuda.synchronize()
    assert torch.equal(infer(xd), want), "replay after pack-cache churn differs (graph read freed weights)"
    with torch.no_grad():
        assert torch.equal(infer_b(big.to(DEV)), model(big.to(DEV)))
    del junk
    # refresh() picks up changed parameters
    with torch.no_grad():
        for p in model.parameters():
            p.mul_(1.01)
        want3 = model(xd).clone()
    assert not torch.equal(want3, want)
    infer.refresh()
    assert torch.equal(infer(xd), want3)
    infer.close(); infer_b.close()


def _fresh_graph_step(sd, xd, yd, lr):
    model = _model(sd).train()
    crit = U.DeepSupervisionLoss3D()
    opt = U.make_adamw(model, lr=lr, weight_decay=1e-4, capturable=True)
    step = U.GraphedTrainStep(model, crit, opt, xd, yd, warmup=1)
    model.load_state_dict(sd)                      # undo the warm-up / capture optimizer steps
    for st in opt.state.values():
        for v in st.values():
            if torch.is_tensor(v):
                v.zero_()
    return model, opt, step


def test_prefetch_pipeline_equals_direct_call():
    """GraphedTrainStep.prefetch/step_prefetched (the e2e path of bench.py: pinned host -> staging on a copy stream ->
    static buffers -> replay) produces exactly the losses of __call__ on the same batches."""
    sd = O.make_state_dict(4, 4, FEATS, seed=46)
    batches = [O.make_inputs(2, 32, 32, 32, seed=50 + i) for i in range(3)]
    x0, y0 = batches[0]
    out = {}
    for mode in ("call", "prefetch"):
        model, opt, step = _fresh_graph_step(sd, x0.to(DEV), y0.to(DEV), 1e-3)
        losses = []
        if mode == "call":
            for xb, yb in batches:
                losses.append(float(step(xb.to(DEV), yb.to(DEV))))
        else:
            pinned = [(xb.pin_memory(), yb.pin_memory()) for xb, yb in batches]
            step.prefetch(*pinned[0])
            for i in range(len(pinned)):
                lt = step.step_prefetched()
                if i + 1 < len(pinned):
                    step.prefetch(*pinned[i + 1])
                losses.append(lt.item())
        out[mode] = losses
        step.close()
    for a, b in zip(out["call"], out["prefetch"]):
        assert abs(a - b) <= 2e-4 * abs(a), out   # same kernels, same order: only the fp32-atomic jitter of the wgrad flush


def test_scheduler_changes_the_update_under_graph_replay():
    """a python-float lr is baked into a captured optimizer; GraphedTrainStep keeps the lr in a device
    tensor so CosineAnnealingWarmRestarts (training.py:194-196,252) keeps working behind the replayed graph."""
    sd = O.make_state_dict(4, 4, FEATS, seed=47)
    x, y = O.make_inputs(1, 32, 32, 32, seed=47)
    xd, yd = x.to(DEV), y.to(DEV)
    model = _model(sd).train()
    opt = torch.optim.AdamW(model.parameters(), lr=1e-3, weight_decay=0.0, fused=True, capturable=True)   # FLOAT lr on purpose
    step = U.GraphedTrainStep(model, U.DeepSupervisionLoss3D(), opt, xd, yd, warmup=1)
    assert torch.is_tensor(opt.param_groups[0]["lr"]) and opt.param_groups[0]["lr"].is_cuda
    sched = U.make_scheduler(opt, T_0=2, T_mult=1, eta_min=1e-6)
    w = model.downs[0].double_conv[0].weight
    deltas, lrs = [], []
    for _ in range(4):
        before = w.detach().clone()
        step(xd, yd)
        torch.cuda.synchronize()
        deltas.append(float((w.detach() - before).abs().max()))
        lrs.append(float(opt.param_groups[0]["lr"]))
        sched.step()
    # T_0 = 2: lr alternates 1e-3, ~5e-4, 1e-3, ~5e-4; Adam's update magnitude ~ lr per element
    assert abs(lrs[0] - 1e-3) < 1e-9 and abs(lrs[1] - 5.005e-4) < 1e-6 and abs(lrs[2] - 1e-3) < 1e-9, lrs
    assert deltas[1] < 0.75 * deltas[0] and deltas[2] > 1.3 * deltas[1], (deltas, lrs)
    step.close()


def test_fused_adamw_matches_torch_adamw_and_keeps_packed_weights_current():
    """optim.FusedAdamW: 5 steps with the trainer's hyper-parameters (training.py:186-191) against
    torch.optim.AdamW on identical gradients — parameters and both moments agree to fp32 rounding; the bf16 packed copies the
    kernel wrote in the same pass are BIT-identical to a fresh pack of the updated parameter; a scheduler moves the lr; the
    state_dict is torch.optim.AdamW-compatible (loads into torch's optimizer and back)."""
    from unet3d_b200 import ops
    sd = O.make_state_dict(4, 4, FEATS, seed=48)
    ma, mb = _model(sd), _model(sd)
    lr0 = 3e-3
    oa = U.FusedAdamW(ma, lr=torch.tensor(lr0, device=DEV), weight_decay=1e-4, betas=(0.9, 0.999))
    ob = torch.optim.AdamW(mb.parameters(), lr=lr0, weight_decay=1e-4, betas=(0.9, 0.999))
    sa, sb = U.make_scheduler(oa, T_0=3, T_mult=1, eta_min=1e-6), U.make_scheduler(ob, T_0=3, T_mult=1, eta_min=1e-6)
    g = torch.Generator().manual_seed(5)
    skip = {"deep_supervision.3.weight", "deep_supervision.3.bias"}      # never receive a gradient (losses.py:118-124)
    for step in range(5):
        for (k, pa), (_, pb) in zip(ma.named_parameters(), mb.named_parameters()):
            if k in skip:
                pa.grad = pb.grad = None
                continue
            gr = (torch.randn(pa.shape, generator=g) * (0.1 + 0.05 * step)).to(DEV)
            pa.grad, pb.grad = gr.clone(), gr.clone()
        oa.step(); ob.step(); sa.step(); sb.step()
        assert abs(float(oa.param_groups[0]["lr"]) - ob.param_groups[0]["lr"]) < 1e-9
    packed = U.optim.packed_conv_params(ma)
    assert len(packed) == 49        # 11 blocks x 3 convs + 5 gates x 2 + 5 transposed convs + final_conv.0
    for (k, pa), (_, pb) in zip(ma.named_parameters(), mb.named_parameters()):
        if k in skip:
            assert torch.equal(pa, pb) and pa not in oa.state
            continue
        scale = float(pb.abs().max()) + 1e-12
        assert float((pa - pb).abs().max()) <= 2e-5 * scale + 1e-7, k
        for key in ("exp_avg", "exp_avg_sq"):
            ref = ob.state[pb][key]
            assert float((oa.state[pa][key] - ref).abs().max()) <= 1e-5 * float(ref.abs().max()) + 1e-12, (k, key)
        if id(pa) in packed:
            convt = packed[id(pa)][1]
            fresh = ops.pack_weight_pair(pa, convt)
            pin = pa.__dict__["_b3d_pack_pinned"][2]
            for mode, (buf, kp, rows) in fresh.items():
                assert pin[mode][1:] == (kp, rows)
                assert torch.equal(pin[mode][0], buf), "packed copy of %s (mode %d) is stale" % (k, mode)
    assert float(oa.state[next(iter(oa.state))]["step"]) == 5.0
    # the pinned copies are what the conv kernels use; a parameter changed behind the optimizer drops the pin
    from unet3d_b200 import functional
    w = ma.downs[0].double_conv[0].weight
    assert functional.packed(w, ops.PACK_FPROP)[0] is w.__dict__["_b3d_pack_pinned"][2][ops.PACK_FPROP][0]
    with torch.no_grad():
        w.mul_(2.0)
    got = functional.packed(w, ops.PACK_FPROP)[0]
    assert "_b3d_pack_pinned" not in w.__dict__ and torch.equal(got, ops.pack_weight_pair(w, False)[ops.PACK_FPROP][0])
    # checkpoint compatibility (training.py:398-404 saves optimizer.state_dict())
    sd_a = oa.state_dict()
    oc = torch.optim.AdamW(ma.parameters(), lr=lr0, weight_decay=1e-4)
    oc.load_state_dict(sd_a)
    od = U.FusedAdamW(ma, lr=torch.tensor(lr0, device=DEV), weight_decay=1e-4)
    od.load_state_dict(ob.state_dict())
    assert float(od._step_t) == 5.0


def test_fused_adamw_training_trajectory_follows_the_oracle():
    """Five steps of the real loop with FusedAdamW (eager): the loss trajectory follows the fp32 CPU oracle + torch.optim.AdamW."""
    sd = O.make_state_dict(4, 4, FEATS, seed=13)
    x, y = O.make_inputs(2, 32, 32, 32, seed=13)
    steps, lr = 5, 2e-3
    params = {k: v.clone().requires_grad_(True) for k, v in sd.items() if v.is_floating_point() and "running" not in k}
    osd = {k: v.clone() for k, v in sd.items()}
    osd.update(params)
    oopt = torch.optim.AdamW(list(params.values()), lr=lr, weight_decay=1e-4)
    ref = []
    for _ in range(steps):
        oopt.zero_grad()
        main, deep, _ = O.unet_forward(x, osd, FEATS, training=True, dropout_masks=None)
        loss = O.deep_supervision_loss(main, deep, y)
        loss.backward()
        oopt.step()
        ref.append(float(loss.detach()))
    model = _model(sd).train()
    crit = U.DeepSupervisionLoss3D()
    opt = U.make_adamw(model, lr=lr, weight_decay=1e-4)
    assert isinstance(opt, U.FusedAdamW)
    xd, yd = x.to(DEV), y.to(DEV)
    got = []
    for _ in range(steps):
        opt.zero_grad(set_to_none=True)
        loss = crit(model(xd), yd)
        loss.backward()
        opt.step()
        got.append(float(loss.detach()))
    for i, (a, b) in enumerate(zip(ref, got)):
        assert abs(a - b) <= 1.5e-2 * abs(a), "step %d: oracle %s vs b200 %s" % (i, ref, got)
    assert (ref[0] - got[-1]) >= 0.7 * (ref[0] - ref[-1]), "b200 path trains slower than the oracle: %s vs %s" % (got, ref)


def test_loss_and_metric_reject_non_int64_targets():
    logits = torch.randn(1, 4, 8, 8, 8, device=DEV)
    for bad in (torch.zeros(1, 8, 8, 8, dtype=torch.uint8, device=DEV), torch.zeros(1, 8, 8, 8, dtype=torch.int32, device=DEV),
                torch.zeros(1, 8, 8, 4, dtype=torch.int64, device=DEV), torch.zeros(1, 8, 8, 8, dtype=torch.int64)):
        with pytest.raises(Exception):
            U.CombinedLoss3D()(logits, bad)
        with pytest.raises(Exception):
            U.calculate_dice_score(logits, bad)


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs 2 GPUs (NCCL)")
def test_data_parallel_gradients_equal_mean_of_local_gradients_nccl():
    """scripts/dp_check.py under torchrun, 2 ranks over NCCL: the gradients DataParallel leaves in .grad equal the mean of the
    ranks' local gradients (whole model 1e-4, worst tensor max(1e-4, 10 x run-to-run jitter))."""
    env = dict(os.environ, DP_CHECK_SIZE="32", MASTER_ADDR="127.0.0.1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29731", os.path.join(ROOT, "scripts", "dp_check.py")]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    print(r.stdout[-2500:])
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
