"""GPU input pipeline (/root/reference/training.py:117-172) against the oracle
(oracle/preprocess_oracle.py, pinned to the reference's own BraTSDataset methods by tests/golden/preprocess.json).

Tolerances: percentiles are exact order statistics (bit-exact against numpy on fp32-representable inputs); mean / std 1e-9
relative (fp64 accumulation); the resized z-scored image 2e-6 absolute (double arithmetic, one float32 rounding at the end);
label maps bit-exact; rot90 / flips bit-exact, intensity scale to one float32 rounding; the noise field is checked statistically
(a device counter RNG cannot reproduce numpy's MT19937 stream)."""
import json
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    import b3d  # noqa: F401
    import unet3d_b200 as U

from oracle import preprocess_oracle as P

DEV = "cuda:0"
HERE = os.path.dirname(os.path.abspath(__file__))


def _volume(seed, shape):
    rng = np.random.RandomState(seed)
    vol = np.round(rng.gamma(2.0, 300.0, size=shape)).astype(np.float64)
    vol[rng.rand(*shape) < 0.3] = 0
    seg = rng.choice([0, 1, 2, 4], size=shape, p=[0.7, 0.1, 0.1, 0.1]).astype(np.float64)
    return vol, seg


def _idx(numel, k=512):
    g = np.random.RandomState(12345)
    return np.sort(g.choice(numel, size=min(k, numel), replace=False))


@pytest.mark.parametrize("seed,shape", [(3, (30, 36, 28)), (4, (155, 48, 40)), (5, (128, 128, 128)), (6, (17, 250, 33))])
def test_preprocess_image_and_labels_vs_oracle(seed, shape):
    vol, seg = _volume(seed, shape)
    ref_img, st = P.preprocess_image(vol)
    ref_lab = P.preprocess_segmentation(seg)
    img, stats = U.preprocess_image(vol, device=DEV, return_stats=True)
    stats = stats.cpu().numpy()
    assert stats[0] == st["p1"] and abs(stats[1] - st["p99"]) <= 1e-12 * max(1.0, abs(st["p99"])), (stats, st)
    assert abs(stats[2] - st["mean"]) <= 1e-9 * abs(st["mean"]) and abs(stats[3] - st["std"]) <= 1e-9 * st["std"]
    assert img.shape == (128, 128, 128) and img.dtype == torch.float32
    assert float((img.cpu() - torch.from_numpy(ref_img)).abs().max()) <= 2e-6
    for dtype in (torch.int64, torch.uint8):
        lab = U.preprocess_segmentation(seg, dtype=dtype, device=DEV)
        assert lab.dtype == dtype and torch.equal(lab.cpu().to(torch.uint8), torch.from_numpy(ref_lab))
    assert set(np.unique(ref_lab).tolist()) <= {0, 1, 2, 3}


def test_preprocess_matches_the_reference_golden():
    """Straight against the numbers the reference's own methods produced (tests/golden/make_golden_preprocess.py)."""
    g = json.load(open(os.path.join(HERE, "golden", "preprocess.json")))
    for case in g["cases"]:
        vol, seg = _volume(case["seed"], tuple(case["shape"]))
        img = U.preprocess_image(vol, device=DEV).cpu().double().reshape(-1).numpy()
        np.testing.assert_allclose(img[_idx(img.size)], case["image"]["samples"], rtol=0, atol=2e-6)
        assert abs(img.sum() - case["image"]["sum"]) <= 2e-6 * img.size
        lab = U.preprocess_segmentation(seg, dtype=torch.uint8, device=DEV).cpu().numpy()
        assert [int(v) for v in np.bincount(lab.reshape(-1), minlength=4)] == case["label_counts"]
        assert [int(v) for v in lab.reshape(-1)[_idx(lab.size)]] == case["label_samples"]


def test_percentiles_are_exact_order_statistics_with_ties_and_negatives():
    rng = np.random.RandomState(0)
    for n, maker in ((1, lambda: rng.randn(1)), (2, lambda: rng.randn(2)), (1000, lambda: np.round(rng.randn(1000) * 3)),
                     (100003, lambda: rng.randn(100003).astype(np.float32).astype(np.float64) * 1e3),
                     (50000, lambda: np.zeros(50000)), (70001, lambda: -np.abs(rng.randn(70001)).astype(np.float32).astype(np.float64))):
        x = np.asarray(maker()).astype(np.float32).astype(np.float64)   # the device path works on fp32-representable inputs
        p = P.percentiles(x)
        xd = torch.from_numpy(x.astype(np.float32)).to(DEV)
        st = U.preprocess.clip_stats(xd).cpu().numpy()
        assert abs(st[0] - p[0]) <= 1e-12 * max(1.0, abs(p[0])) and abs(st[1] - p[1]) <= 1e-12 * max(1.0, abs(p[1])), (n, st, p)
        c = np.clip(x, p[0], p[1])
        assert abs(st[2] - c.mean()) <= 1e-9 * max(1.0, abs(c.mean())) and abs(st[3] - c.std()) <= 1e-7 * max(1.0, c.std())


def test_preprocess_case_returns_trainer_tensors():
    vols = [_volume(10 + c, (40, 44, 36))[0] for c in range(4)]
    seg = _volume(10, (40, 44, 36))[1]
    image, mask = U.preprocess_case(vols, seg, device=DEV)
    assert image.shape == (4, 128, 128, 128) and image.dtype == torch.float32 and mask.shape == (128, 128, 128) and mask.dtype == torch.int64
    for c in range(4):
        ref, _ = P.preprocess_image(vols[c])
        assert float((image[c].cpu() - torch.from_numpy(ref)).abs().max()) <= 2e-6
    # the tensors plug straight into the model / loss / metric
    model = U.UNet3D(4, 4, features=[16, 32, 64, 128, 256]).to(DEV).eval()
    with torch.no_grad():
        out = model(image[None])
    assert out.shape == (1, 4, 128, 128, 128)
    assert 0.0 <= U.calculate_dice_score(out, mask[None]) <= 1.0


@pytest.mark.parametrize("k,flips", [(0, (False, False, False)), (1, (False, False, False)), (2, (True, False, True)),
                                     (3, (False, True, False)), (1, (True, True, True)), (0, (False, False, True))])
def test_augment_spatial_part_is_exact(k, flips):
    rng = np.random.RandomState(7)
    image = rng.randn(4, 16, 16, 24).astype(np.float32)
    seg = rng.randint(0, 4, size=(16, 16, 24)).astype(np.uint8)
    ref_img, ref_seg = P.spatial_augment(image, seg, k, list(flips))
    scale = 1.0625    # exactly representable: the product is exact in float32
    for dtype in (torch.uint8, torch.int64):
        out, oseg = U.apply_augmentations(torch.from_numpy(image).to(DEV), torch.from_numpy(seg).to(DEV).to(dtype), k=k, flips=flips,
                                          noise_std=0.0, scale=scale)
        assert torch.equal(out.cpu(), torch.from_numpy(np.ascontiguousarray(ref_img) * np.float32(scale)))
        assert torch.equal(oseg.cpu().to(torch.uint8), torch.from_numpy(np.ascontiguousarray(ref_seg)))


def test_augment_noise_statistics_and_reference_golden():
    g = json.load(open(os.path.join(HERE, "golden", "preprocess.json")))
    vol, seg = _volume(3, (30, 36, 28))
    img, _ = P.preprocess_image(vol)
    lab = P.preprocess_segmentation(seg)
    image4 = np.stack([img[:16, :16, :16] * (1 + 0.1 * c) for c in range(4)], 0).astype(np.float32)
    lab16 = np.ascontiguousarray(lab[:16, :16, :16])
    for rec in g["augment"]:   # the reference's own outputs: same decisions, noise removed -> equal up to the noise amplitude
        prm = rec["params"]
        out, oseg = U.apply_augmentations(torch.from_numpy(image4).to(DEV), torch.from_numpy(lab16).to(DEV), k=prm["k"], flips=prm["flips"],
                                          noise_std=0.0, scale=prm["scale"])
        flat = out.cpu().double().reshape(-1).numpy()
        want = np.asarray(rec["image"]["samples"])
        assert np.abs(flat[_idx(flat.size)] - want).max() <= 6.0 * prm["noise_std"] * prm["scale"] + 1e-5   # 6 sigma of the reference's noise
        assert [int(v) for v in oseg.cpu().reshape(-1)[_idx(lab16.size)]] == rec["label_samples"]
    # the device noise field: zero mean, the requested standard deviation, different per channel / voxel / seed, reproducible
    base = torch.zeros(4, 32, 32, 32, device=DEV)
    a = U.apply_augmentations(base, None, noise_std=0.05, scale=1.0, seed=123)
    b = U.apply_augmentations(base, None, noise_std=0.05, scale=1.0, seed=123)
    c = U.apply_augmentations(base, None, noise_std=0.05, scale=1.0, seed=124)
    assert torch.equal(a, b) and not torch.equal(a, c)
    assert abs(float(a.mean())) < 1e-3 and abs(float(a.std()) - 0.05) < 1e-3
    assert abs(float(torch.corrcoef(torch.stack([a[0].reshape(-1), a[1].reshape(-1)]))[0, 1])) < 0.02
    kurt = float(((a / a.std()) ** 4).mean())
    assert abs(kurt - 3.0) < 0.1      # Gaussian


def test_preprocess_has_no_cpu_fallback_and_times_a_brats_sized_case():
    with pytest.raises(Exception):
        U.apply_augmentations(torch.zeros(4, 8, 8, 8), None)
    vols = [np.round(np.random.RandomState(c).gamma(2.0, 300.0, size=(240, 240, 155))).astype(np.float32) for c in range(4)]
    seg = np.random.RandomState(9).choice([0, 1, 2, 4], size=(240, 240, 155)).astype(np.float32)
    dv = [torch.from_numpy(v).to(DEV) for v in vols]
    ds = torch.from_numpy(seg).to(DEV)
    for _ in range(2):
        U.preprocess_case(dv, ds, device=DEV)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        image, mask = U.preprocess_case(dv, ds, device=DEV)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print("PREPROCESS 4 x 240x240x155 -> 4 x 128^3 + mask: %.3f ms per case on the device" % ms)
    assert ms < 20.0
