"""GPU input pipeline in front of the hot path — the work of `BraTSDataset.__getitem__`
(/root/reference/training.py:76-115) after the NIfTI files are read:

    preprocess_image          training.py:117-132   percentile clip (1, 99) -> z-score -> trilinear resize to 128^3 -> float32
    preprocess_segmentation   training.py:134-146   label 4 -> 3 -> nearest resize -> integer mask
    preprocess_case           training.py:82-104    4 modalities stacked + mask, as the tensors the trainer feeds the model
    draw_augmentation / apply_augmentations   training.py:148-172

On the host this costs ~1 s per case (a full sort for np.percentile and two scipy zooms per modality) while one B200 trains on
a case every ~9 ms; here a case is ~10 kernel launches per modality on data that stays L2 resident (csrc/preprocess.cu).
Inputs may be numpy arrays (any float dtype; uploaded as fp32 — MRI intensities are integers < 2^24, exactly representable) or
torch tensors.  No CPU fallback: without the CUDA library every function raises.
"""
import ctypes

import numpy as np
import torch

from . import _lib
from ._lib import c_int, c_ll, c_sz, check, ptr, stream_ptr

c_double, c_float, c_ull = ctypes.c_double, ctypes.c_float, ctypes.c_ulonglong
TARGET = (128, 128, 128)
_WORK_BYTES = 4 * 2048 * 4 + 64


def _dev(device):
    device = torch.device("cuda" if device is None else device)
    if device.type != "cuda":
        raise _lib.B3DError("preprocess: a CUDA (sm_100) device is required — the b200 path has no CPU fallback")
    if device.index is None:
        device = torch.device("cuda", torch.cuda.current_device())
    _lib.require_device(device)
    return device


def _f32(vol, device):
    if isinstance(vol, np.ndarray):
        vol = torch.from_numpy(np.ascontiguousarray(vol, dtype=np.float32))
    if not torch.is_tensor(vol):
        raise TypeError("expected a numpy array or a torch tensor")
    return vol.to(device=device, dtype=torch.float32, non_blocking=True).contiguous()


def clip_stats(x, q_lo=1.0, q_hi=99.0):
    """x: fp32 CUDA tensor (any shape).  Returns float64[4] on the device: p_lo, p_hi (np.percentile, linear), mean and
    population std of clip(x, p_lo, p_hi) — training.py:119-124."""
    work = torch.empty(_WORK_BYTES, dtype=torch.uint8, device=x.device)
    stats = torch.empty(4, dtype=torch.float64, device=x.device)
    check(_lib.lib().b3d_clip_stats(ptr(x), c_ll(x.numel()), c_double(q_lo), c_double(q_hi), ptr(work), c_sz(_WORK_BYTES), ptr(stats),
                                    stream_ptr()))
    return stats


def preprocess_image(volume, target=TARGET, out=None, device=None, return_stats=False):
    """One modality [D,H,W] -> float32 [target] on the device (training.py:117-132)."""
    device = _dev(device if device is not None else (volume.device if torch.is_tensor(volume) and volume.is_cuda else None))
    x = _f32(volume, device)
    if x.dim() != 3:
        raise ValueError("preprocess_image expects a 3-D volume, got %s" % (tuple(x.shape),))
    with torch.cuda.device(device):
        stats = clip_stats(x)
        if out is None:
            out = torch.empty(tuple(target), dtype=torch.float32, device=device)
        assert out.is_contiguous() and out.dtype == torch.float32 and tuple(out.shape) == tuple(target)
        d, h, w = x.shape
        check(_lib.lib().b3d_zoom_normalize(ptr(x), c_int(d), c_int(h), c_int(w), ptr(stats), ptr(out), c_int(target[0]),
                                            c_int(target[1]), c_int(target[2]), stream_ptr()))
    return (out, stats) if return_stats else out


def preprocess_segmentation(seg, target=TARGET, dtype=torch.int64, out=None, device=None):
    """BraTS label map [D,H,W] with values {0,1,2,4} -> {0,1,2,3} at [target] (training.py:134-146).  dtype: torch.int64 (what the
    trainer feeds the loss, training.py:103) or torch.uint8 (what the reference's numpy array holds)."""
    device = _dev(device if device is not None else (seg.device if torch.is_tensor(seg) and seg.is_cuda else None))
    x = _f32(seg, device)
    if x.dim() != 3:
        raise ValueError("preprocess_segmentation expects a 3-D label map, got %s" % (tuple(x.shape),))
    if dtype not in (torch.int64, torch.uint8):
        raise ValueError("dtype must be torch.int64 or torch.uint8")
    with torch.cuda.device(device):
        if out is None:
            out = torch.empty(tuple(target), dtype=dtype, device=device)
        d, h, w = x.shape
        check(_lib.lib().b3d_zoom_labels(ptr(x), c_int(d), c_int(h), c_int(w), ptr(out), c_int(0 if dtype == torch.uint8 else 1),
                                         c_int(target[0]), c_int(target[1]), c_int(target[2]), stream_ptr()))
    return out


def preprocess_case(modalities, seg, target=TARGET, device=None):
    """4 modalities (list of [D,H,W] volumes or one [4,D,H,W] array, order t1/t1ce/t2/flair as training.py:35,82-91) + label map
    -> (image float32 [4,*target], mask int64 [*target]) on the device: the tensors of training.py:102-103."""
    device = _dev(device)
    n = len(modalities)
    image = torch.empty((n,) + tuple(target), dtype=torch.float32, device=device)
    for c in range(n):
        preprocess_image(modalities[c], target, out=image[c], device=device)
    mask = preprocess_segmentation(seg, target, torch.int64, device=device)
    return image, mask


def draw_augmentation(rng=np.random):
    """The random decisions of training.py:148-172 drawn on the host in the reference's order (rotation?, k, three flips,
    noise_std, scale).  NOTE: the reference draws the whole Gaussian noise field from the same numpy stream BETWEEN noise_std and
    scale; the device draws the field from a counter RNG instead, so `scale` is not the value the reference would draw next."""
    k = 0
    if rng.rand() > 0.5:
        k = int(rng.randint(1, 4))
    flips = [bool(rng.rand() > 0.5) for _ in range(3)]
    noise_std = float(rng.uniform(0, 0.1))
    scale = float(rng.uniform(0.9, 1.1))
    seed = int(rng.randint(0, 2 ** 31 - 1))
    return {"k": k, "flips": flips, "noise_std": noise_std, "scale": scale, "seed": seed}


def apply_augmentations(image, seg=None, k=0, flips=(False, False, False), noise_std=0.0, scale=1.0, seed=0):
    """image float32 [C,D,H,W] (+ optional mask [D,H,W] uint8 / int64) on the device -> augmented copies (training.py:148-172):
    rot90 by k in the (D,H) plane, flips along D/H/W (the mask follows the image — the reference's own label flips use the image's
    axis numbers on the 3-D mask, training.py:157-160, which raises for W and misaligns D/H; not reproduced), additive
    N(0, noise_std) noise, intensity scale."""
    if not image.is_cuda:
        raise _lib.B3DError("apply_augmentations: CUDA tensors required — the b200 path has no CPU fallback")
    _lib.require_device(image.device)
    image = image.contiguous()
    if image.dtype != torch.float32 or image.dim() != 4:
        raise ValueError("image must be float32 [C,D,H,W]")
    c, d, h, w = image.shape
    out = torch.empty_like(image)
    lab_dtype, out_seg = 0, None
    if seg is not None:
        if seg.dtype not in (torch.uint8, torch.int64) or tuple(seg.shape) != (d, h, w) or seg.device != image.device:
            raise ValueError("seg must be uint8 / int64 [D,H,W] on the image's device")
        seg = seg.contiguous()
        lab_dtype = 0 if seg.dtype == torch.uint8 else 1
        out_seg = torch.empty_like(seg)
    with torch.cuda.device(image.device):
        check(_lib.lib().b3d_augment(ptr(image), ptr(seg), c_int(lab_dtype), ptr(out), ptr(out_seg), c_int(c), c_int(d), c_int(h),
                                     c_int(w), c_int(int(k)), c_int(int(bool(flips[0]))), c_int(int(bool(flips[1]))),
                                     c_int(int(bool(flips[2]))), c_float(float(noise_std)), c_float(float(scale)), c_ull(int(seed)),
                                     stream_ptr()))
    return (out, out_seg) if seg is not None else out
