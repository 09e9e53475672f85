"""oracle/preprocess_oracle.py — TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

CPU restatement (numpy, float64 like the reference: nibabel's get_fdata() returns float64) of the input pipeline in front of the
hot path, /root/reference/training.py:117-172:

    preprocess_image         training.py:117-132   np.percentile(1, 99) clip -> z-score -> ndimage.zoom(order=1) to 128^3 -> float32
    preprocess_segmentation  training.py:134-146   label 4 -> 3 -> ndimage.zoom(order=0) -> uint8
    apply_augmentations      training.py:148-172   rot90 in (D,H), flips, Gaussian noise, intensity scale (numpy global RNG)

The two zooms are restated WITHOUT scipy (explicit index arithmetic, scipy's NI_ZoomShift conventions: output index o samples
input coordinate o * (in - 1) / (out - 1); order 1 = linear interpolation, order 0 = floor(c + 0.5)), so the CUDA kernels have
a line-by-line reference.  Parity status: PINNED — tests/golden/make_golden_preprocess.py executes the reference's own three
methods (which do call scipy) on seeded volumes and tests/test_oracle_golden.py checks this file against those outputs.
"""
import numpy as np

TARGET = (128, 128, 128)


def percentiles(image, qs=(1.0, 99.0)):
    """np.percentile(..., method='linear') from explicit order statistics: rank r = q/100 * (n-1), lerp(a[floor r], a[ceil r])."""
    a = np.sort(np.asarray(image, dtype=np.float64).reshape(-1))
    n = a.size
    out = []
    for q in qs:
        r = q / 100.0 * (n - 1)
        lo = int(np.floor(r))
        hi = min(lo + 1, n - 1)
        out.append(a[lo] + (a[hi] - a[lo]) * (r - lo))
    return out


def zoom_linear(v, shape):
    """ndimage.zoom(v, [t/s], order=1): separable linear interpolation at coordinates o * (in-1)/(out-1) (training.py:128-130)."""
    out = np.asarray(v, dtype=np.float64)
    for ax, o_n in enumerate(shape):
        i_n = out.shape[ax]
        if o_n == i_n:
            continue
        zoom = (i_n - 1) / (o_n - 1) if o_n > 1 else 0.0
        c = np.arange(o_n) * zoom
        i0 = np.minimum(np.floor(c).astype(np.int64), i_n - 1)
        i1 = np.minimum(i0 + 1, i_n - 1)
        f = c - i0
        sh = [1] * out.ndim
        sh[ax] = o_n
        out = np.take(out, i0, axis=ax) * (1.0 - f.reshape(sh)) + np.take(out, i1, axis=ax) * f.reshape(sh)
    return out


def zoom_nearest(v, shape):
    """ndimage.zoom(v, [t/s], order=0): index floor(o * (in-1)/(out-1) + 0.5) per axis (training.py:142-144)."""
    out = np.asarray(v)
    for ax, o_n in enumerate(shape):
        i_n = out.shape[ax]
        if o_n == i_n:
            continue
        zoom = (i_n - 1) / (o_n - 1) if o_n > 1 else 0.0
        idx = np.minimum(np.floor(np.arange(o_n) * zoom + 0.5).astype(np.int64), i_n - 1)
        out = np.take(out, idx, axis=ax)
    return out


def preprocess_image(image, target=TARGET):
    """training.py:117-132.  Returns (float32 [target], dict(p1, p99, mean, std)) — the statistics are what the GPU path must match."""
    image = np.asarray(image, dtype=np.float64)
    p1, p99 = percentiles(image)
    image = np.clip(image, p1, p99)
    mean, std = float(np.mean(image)), float(np.std(image))
    image = (image - mean) / (std + 1e-8)
    if image.shape != tuple(target):
        image = zoom_linear(image, target)
    return image.astype(np.float32), {"p1": p1, "p99": p99, "mean": mean, "std": std}


def preprocess_segmentation(seg, target=TARGET):
    """training.py:134-146."""
    seg = np.array(seg, dtype=np.float64, copy=True)
    seg[seg == 4] = 3
    if seg.shape != tuple(target):
        seg = zoom_nearest(seg, target)
    return seg.astype(np.uint8)


def draw_augmentation(rng=np.random):
    """The random decisions of training.py:148-172 in the reference's exact draw order (so the same numpy RNG state gives the same
    decisions); the Gaussian noise field itself is drawn by apply_augmentations."""
    k = 0
    if rng.rand() > 0.5:
        k = int(rng.randint(1, 4))
    flips = [bool(rng.rand() > 0.5) for _ in range(3)]
    noise_std = float(rng.uniform(0, 0.1))
    return {"k": k, "flips": flips, "noise_std": noise_std}


def spatial_augment(image, seg, k, flips):
    """rot90 by k in the (D,H) plane of every channel, then flips along D, H, W (training.py:150-160)."""
    image, seg = np.asarray(image), np.asarray(seg)
    if k:
        image = np.stack([np.rot90(image[i], k, axes=(0, 1)) for i in range(image.shape[0])], axis=0)
        seg = np.rot90(seg, k, axes=(0, 1))
    for ax, f in enumerate(flips):
        if f:
            image = np.flip(image, axis=ax + 1)
            seg = np.flip(seg, axis=ax)
    return image, seg


def apply_augmentations(image, seg, rng=np.random):
    """training.py:148-172 with the reference's draw order: [rot?][k] flip x3, noise_std, noise field, scale."""
    p = draw_augmentation(rng)
    image, seg = spatial_augment(image, seg, p["k"], p["flips"])
    noise = rng.normal(0, p["noise_std"], image.shape)
    image = image + noise
    scale = float(rng.uniform(0.9, 1.1))
    image = image * scale
    p["scale"] = scale
    return image.copy(), seg.copy(), p
