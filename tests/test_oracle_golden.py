"""Pins oracle/unet3d_oracle.py against the golden vectors produced by executing the reference's own classes
(tests/golden/make_golden.py).  CPU only."""
import os

import numpy as np
import pytest
import torch

from conftest import sample_idx
from oracle import unet3d_oracle as O


def _check_summary(t, rec, rtol=2e-4, atol=2e-5):
    t = t.detach().double().reshape(-1)
    idx = torch.from_numpy(sample_idx(t.numel()))
    np.testing.assert_allclose(t[idx].numpy(), np.array(rec["samples"]), rtol=rtol, atol=atol)
    assert abs(float(t.norm()) - rec["l2"]) <= rtol * max(1.0, rec["l2"])


def test_state_dict_keys_match_reference(golden):
    for cfg, (cin, feats) in {"default4": (4, (32, 64, 128, 256, 512)), "default1": (1, (32, 64, 128, 256, 512)),
                              "light4": (4, (16, 32, 64, 128, 256))}.items():
        assert [[k, list(s)] for k, s in O.param_shapes(cin, 4, feats)] == golden["keys_" + cfg]
    n = sum(int(np.prod(s)) for k, s in O.param_shapes(4, 4) if "running" not in k and "num_batches" not in k)
    assert n == golden["nparams_default4"] == 92342623


@pytest.mark.parametrize("case", ["model_small", "model_small_n2", "model_small_dropout"])
def test_model_forward_backward_matches_reference(golden, golden_arrays, case):
    rec = golden[case]
    feats = tuple(rec["features"])
    sd = O.make_state_dict(4, 4, feats, seed=rec["seed"])
    x, y = O.make_inputs(rec["n"], rec["size"], rec["size"], rec["size"], seed=rec["seed"])
    with torch.no_grad():
        ev, _, _ = O.unet_forward(x, sd, feats, training=False)
    _check_summary(ev, rec["eval_logits"])
    assert O.confusion_counts(ev, y).tolist() == rec["eval_confusion"]
    assert abs(O.dice_score(ev, y) - rec["eval_dice_score"]) < 1e-6
    arrays = golden_arrays(case)
    masks = None
    if rec["dropout"] > 0:
        flat = torch.from_numpy(arrays["dropout_masks"])
        masks, o = [], 0
        for f in feats:
            masks.append(flat[o:o + rec["n"] * f].reshape(rec["n"], f))
            o += rec["n"] * f
    sdg = {k: v.clone().requires_grad_(v.is_floating_point() and "running" not in k) for k, v in sd.items()}
    main, deep, bn = O.unet_forward(x, sdg, feats, training=True, dropout_masks=masks)
    loss = O.deep_supervision_loss(main, deep, y)
    loss.backward()
    assert abs(float(loss) - rec["ds_loss"]) < 2e-5
    _check_summary(main, rec["train_main"])
    for d, r in zip(deep, rec["train_deep"]):
        _check_summary(d, r)
    np.testing.assert_allclose(bn[0].detach().numpy(), rec["bn_running_mean"], rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(bn[1].detach().numpy(), rec["bn_running_var"], rtol=1e-4, atol=1e-6)
    parts = O.combined_loss3d(main.detach(), y)[1]
    for k, v in rec["combined3d_main"].items():
        assert abs(float(parts[k]) - v) < 2e-5
    assert abs(float(O.trainer_combined_loss(main.detach(), y)) - rec["trainer_combined_main"]) < 2e-5
    for k, g in rec["grads"].items():
        if g is None:
            assert sdg[k].grad is None or float(sdg[k].grad.abs().max()) == 0.0
        else:
            got = float(sdg[k].grad.double().norm())
            assert abs(got - g["l2"]) <= 5e-3 * g["l2"] + 1e-6, (k, got, g["l2"])
    if "eval_logits" in arrays:
        np.testing.assert_allclose(ev.numpy(), arrays["eval_logits"], rtol=1e-4, atol=2e-5)
        for k, a in arrays.items():
            if k.startswith("grad."):
                gk = sdg[k[5:]].grad.numpy()
                np.testing.assert_allclose(gk, a, rtol=5e-3, atol=1e-5 + 5e-3 * np.abs(a).max())


def test_losses_match_reference(golden, golden_arrays):
    for rec in golden["loss"]:
        g = torch.Generator().manual_seed(rec["seed"])
        n, s = rec["n"], rec["size"]
        logits = (torch.randn(n, 4, s, s, s, generator=g) * 2.0).requires_grad_(True)
        deep = [(torch.randn(n, 4, s, s, s, generator=g) * 1.5).requires_grad_(True) for _ in range(4)]
        y = torch.randint(0, 4, (n, s, s, s), generator=g)
        arrays = golden_arrays("loss_seed%d" % rec["seed"])
        tot, parts = O.combined_loss3d(logits, y)
        for k, v in rec["combined3d"].items():
            assert abs(float(parts[k]) - v) < 1e-5
        tot.backward()
        np.testing.assert_allclose(logits.grad.numpy(), arrays["combined3d_grad"], rtol=1e-4, atol=1e-9)
        logits.grad = None
        assert abs(float(O.tversky_loss(logits, y)) - rec["tversky"]) < 1e-5
        tl = O.trainer_combined_loss(logits, y)
        assert abs(float(tl) - rec["trainer_combined"]) < 1e-5
        tl.backward()
        np.testing.assert_allclose(logits.grad.numpy(), arrays["trainer_grad"], rtol=1e-4, atol=1e-9)
        logits.grad = None
        ds = O.deep_supervision_loss(logits, deep, y)
        assert abs(float(ds) - rec["ds_loss"]) < 1e-5
        ds.backward()
        np.testing.assert_allclose(logits.grad.numpy(), arrays["ds_grad_main"], rtol=1e-4, atol=1e-9)
        for i in range(3):
            np.testing.assert_allclose(deep[i].grad.numpy(), arrays["ds_grad_deep%d" % i], rtol=1e-4, atol=1e-9)
        assert deep[3].grad is None and not rec["ds_deep3_has_grad"]
        assert abs(O.dice_score(logits.detach(), y) - rec["dice_score"]) < 1e-7
        assert O.confusion_counts(logits.detach(), y).tolist() == rec["confusion"]


def test_blocks_match_reference(golden_arrays):
    arrays = golden_arrays("blocks")
    sd = O.make_state_dict(16, 4, (32, 64, 128, 256, 512), seed=5)
    g = torch.Generator().manual_seed(9)
    x = torch.randn(2, 16, 8, 8, 8, generator=g)
    with torch.no_grad():
        out = O.double_conv(x, sd, "downs.0.")
    np.testing.assert_allclose(out.numpy(), arrays["doubleconv_out"], rtol=1e-4, atol=1e-5)
    gg = torch.randn(2, 32, 8, 8, 8, generator=g)
    xx = torch.randn(2, 32, 8, 8, 8, generator=g)
    with torch.no_grad():
        out = O.attention_gate(gg, xx, sd, "ups.13.")
    np.testing.assert_allclose(out.numpy(), arrays["gate_out"], rtol=1e-4, atol=1e-5)


def test_voxel_counts_semantics():
    g = torch.Generator().manual_seed(3)
    mask = torch.randint(0, 4, (6, 5, 7), generator=g)
    tumour, per_class, per_slice = O.voxel_counts(mask)
    assert tumour == sum(per_class[1:]) == sum(per_slice)
    assert sum(per_class) == mask.numel() and len(per_slice) == 7


def test_classifier_oracle_matches_reference_golden():
    """oracle.classifier_forward against the outputs of the reference's own BrainTumorClassifier (main.py:301-328), stored by
    tests/golden/make_golden_classifier.py."""
    import os
    import numpy as np
    import torch
    from oracle import unet3d_oracle as O
    CASES = [("a", 2, (32, 32, 32), 21), ("b", 1, (16, 32, 48), 22), ("c", 1, (64, 64, 64), 23)]   # = make_golden_classifier.CASES
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "classifier.npz"))
    assert [s.split(" ")[0] for s in z["keys"]] == [k for k, _ in O.classifier_param_shapes(4)]
    for name, n, (d, h, w), seed in CASES:
        sd = O.make_classifier_state_dict(4, seed=seed)
        x, _ = O.make_inputs(n, d, h, w, seed=seed)
        got = O.classifier_forward(x, sd)
        ref = torch.from_numpy(z["logits_" + name])
        assert got.shape == ref.shape
        assert float((got - ref).abs().max()) <= 2e-5 * max(1.0, float(ref.abs().max()))


# ---------------------------------------------------------------------------------------------------------------------
# input pipeline: oracle/preprocess_oracle.py vs the outputs of the reference's own BraTSDataset methods
# ---------------------------------------------------------------------------------------------------------------------
def _pp_golden():
    import json
    with open(os.path.join(os.path.dirname(__file__), "golden", "preprocess.json")) as fh:
        return json.load(fh)


def _pp_idx(numel, k=512):
    g = np.random.RandomState(12345)
    return np.sort(g.choice(numel, size=min(k, numel), replace=False))


def _pp_volume(seed, shape):
    rng = np.random.RandomState(seed)
    vol = np.round(rng.gamma(2.0, 300.0, size=shape)).astype(np.float64)
    vol[rng.rand(*shape) < 0.3] = 0
    seg = rng.choice([0, 1, 2, 4], size=shape, p=[0.7, 0.1, 0.1, 0.1]).astype(np.float64)
    return vol, seg


def test_preprocess_oracle_matches_reference_golden():
    from oracle import preprocess_oracle as P
    g = _pp_golden()
    for case in g["cases"][:2]:   # the 128^3 identity-size case is covered on the GPU side (keeps the CPU suite short)
        vol, seg = _pp_volume(case["seed"], tuple(case["shape"]))
        img, st = P.preprocess_image(vol)
        lab = P.preprocess_segmentation(seg)
        for k in ("p1", "p99", "mean", "std"):
            assert abs(st[k] - case["stats"][k]) <= 1e-9 * max(1.0, abs(case["stats"][k])), k
        flat = img.astype(np.float64).reshape(-1)
        np.testing.assert_allclose(flat[_pp_idx(flat.size)], case["image"]["samples"], rtol=0, atol=1e-6)
        assert abs(flat.sum() - case["image"]["sum"]) <= 1e-3 and abs(np.abs(flat).sum() - case["image"]["abs_sum"]) <= 1e-2
        assert [int(v) for v in np.bincount(lab.reshape(-1), minlength=4)] == case["label_counts"]
        assert [int(v) for v in lab.reshape(-1)[_pp_idx(lab.size)]] == case["label_samples"]


def test_augmentation_oracle_matches_reference_golden():
    from oracle import preprocess_oracle as P
    g = _pp_golden()
    vol, seg = _pp_volume(3, (30, 36, 28))
    img, _ = P.preprocess_image(vol)
    lab = P.preprocess_segmentation(seg)
    image4 = np.stack([img[:16, :16, :16] * (1 + 0.1 * c) for c in range(4)], 0).astype(np.float64)
    lab16 = lab[:16, :16, :16].copy()
    for rec in g["augment"]:
        np.random.seed(rec["seed"])
        a, s, prm = P.apply_augmentations(image4.copy(), lab16.copy())
        assert prm["k"] == rec["params"]["k"] and prm["flips"] == rec["params"]["flips"]
        assert abs(prm["noise_std"] - rec["params"]["noise_std"]) < 1e-15 and abs(prm["scale"] - rec["params"]["scale"]) < 1e-15
        flat = a.reshape(-1)
        np.testing.assert_allclose(flat[_pp_idx(flat.size)], rec["image"]["samples"], rtol=0, atol=1e-12)
        assert [int(v) for v in s.reshape(-1)[_pp_idx(s.size)]] == rec["label_samples"]
