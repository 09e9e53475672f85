"""Golden fixtures of the input pipeline produced by EXECUTING THE REFERENCE's own BraTSDataset methods
(/root/reference/training.py:117-172, sliced by oracle/ref_slice.py) on seeded synthetic volumes — container-only.

    python tests/golden/make_golden_preprocess.py      ->  tests/golden/preprocess.json (+ preprocess.npz, small)

Inputs come from `make_volume(seed)` below (reference-independent); only OUTPUT summaries / samples are stored.
Augmentation: the reference flips the 3-D label map with the IMAGE's axis numbers (training.py:157-160: axis 3 of a 3-D array
raises, axes 1/2 hit H/W instead of D/H), so full-function goldens use (a) a draw with a rotation and no flip, and (b) a draw
with all three flips and the label map passed as [1,D,H,W] — the aligned behaviour the oracle and the GPU path implement."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import preprocess_oracle as P  # noqa: E402
from oracle import ref_slice  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def make_volume(seed, shape):
    """MRI-like non-negative integer intensities with a zero background, and a BraTS label map with values {0,1,2,4}."""
    rng = np.random.RandomState(seed)
    vol = np.round(rng.gamma(2.0, 300.0, size=shape)).astype(np.float64)
    vol[rng.rand(*shape) < 0.3] = 0
    seg = rng.choice([0, 1, 2, 4], size=shape, p=[0.7, 0.1, 0.1, 0.1]).astype(np.float64)
    return vol, seg


def sample_idx(numel, k=512):
    g = np.random.RandomState(12345)
    return np.sort(g.choice(numel, size=min(k, numel), replace=False))


def summarize(a):
    a = np.asarray(a, dtype=np.float64).reshape(-1)
    return {"sum": float(a.sum()), "abs_sum": float(np.abs(a).sum()), "l2": float(np.sqrt((a * a).sum())),
            "samples": [float(v) for v in a[sample_idx(a.size)]]}


def main():
    ns = ref_slice.load()
    golden = {"cases": []}
    arrays = {}
    for seed, shape in ((3, (30, 36, 28)), (4, (155, 48, 40)), (5, (128, 128, 128))):
        vol, seg = make_volume(seed, shape)
        img = ns["_preprocess_image"](None, vol.copy())
        lab = ns["_preprocess_segmentation"](None, seg.copy())
        oimg, st = P.preprocess_image(vol)
        olab = P.preprocess_segmentation(seg)
        assert np.abs(img - oimg).max() <= 1e-6 and np.array_equal(lab, olab), "oracle disagrees with the reference"
        golden["cases"].append({"seed": seed, "shape": list(shape), "stats": {k: float(v) for k, v in st.items()},
                                "image": summarize(img), "label_counts": [int(v) for v in np.bincount(lab.reshape(-1), minlength=4)],
                                "label_samples": [int(v) for v in lab.reshape(-1)[sample_idx(lab.size)]]})
    # augmentation on a 4 x 16^3 crop of case 3 (the numpy global RNG drives the reference)
    vol, seg = make_volume(3, (30, 36, 28))
    img, _ = P.preprocess_image(vol)
    lab = P.preprocess_segmentation(seg)
    image4 = np.stack([img[:16, :16, :16] * (1 + 0.1 * c) for c in range(4)], 0).astype(np.float64)
    lab16 = lab[:16, :16, :16].copy()
    aug = []
    for name, seed, seg_in in (("rot_noflip", 25, lab16), ("flips_norot", 26, lab16[None])):
        np.random.seed(seed)
        a_ref, s_ref = ns["_apply_augmentations"](None, image4.copy(), seg_in.copy())
        np.random.seed(seed)
        a_or, s_or, prm = P.apply_augmentations(image4.copy(), lab16.copy())
        assert np.abs(a_ref - a_or).max() <= 1e-12 and np.array_equal(np.asarray(s_ref).reshape(lab16.shape), s_or), name
        aug.append({"name": name, "seed": seed, "params": {"k": prm["k"], "flips": prm["flips"], "noise_std": prm["noise_std"],
                                                            "scale": prm["scale"]},
                    "image": summarize(a_ref), "label_samples": [int(v) for v in np.asarray(s_ref).reshape(-1)[sample_idx(lab16.size)]]})
    golden["augment"] = aug
    with open(os.path.join(OUT, "preprocess.json"), "w") as fh:
        json.dump(golden, fh, indent=1)
    print("wrote preprocess.json:", [c["stats"] for c in golden["cases"]])


if __name__ == "__main__":
    main()
