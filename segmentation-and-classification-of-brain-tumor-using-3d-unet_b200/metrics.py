"""Dice / voxel-count metrics on the GPU confusion-histogram kernel (integer-exact counts).

    calculate_dice_score       /root/reference/training.py:351-364
    voxel counts / volumes     /root/reference/main.py:470-474,588-591 ; utils/visualization.py:217-221,253
"""
import numpy as np
import torch

from . import _lib, ops


def confusion_matrix(outputs, targets):
    """int64 [K,K] tensor H[pred, true] over the whole batch (device tensor, no sync)."""
    if not outputs.is_cuda:
        raise _lib.B3DError("metrics: CUDA (sm_100) tensors required — the b200 path has no CPU fallback")
    hist, _ = ops.confusion(outputs.detach(), targets)
    return hist


def calculate_dice_score(outputs, targets):
    """Mean Dice of classes 1..3 exactly as the reference computes it (counts cast to fp32, +1e-8) -> python float.
    One device->host copy (16 integers) instead of three .item() calls."""
    h = confusion_matrix(outputs, targets).cpu().numpy()
    scores = []
    for c in range(1, 4):
        inter = np.float32(h[c, c])
        ps, ts = np.float32(h[c, :].sum()), np.float32(h[:, c].sum())
        scores.append(float((np.float32(2.0) * inter) / (ps + ts + np.float32(1e-8))))
    return float(np.mean(scores))


dice_score = calculate_dice_score


def segment(model, volume, return_logits=False):
    """Full-volume inference call site (main.py:382-398): eval forward + argmax -> uint8 mask [N,D,H,W] on device."""
    was_training = model.training
    model.eval()
    with torch.no_grad():
        logits = model(volume)
    if was_training:
        model.train()
    _, mask = ops.confusion(logits, None, want_mask=True)
    return (mask, logits) if return_logits else mask


def tumor_volumes(mask):
    """mask uint8 [D,H,W] -> dict(total tumour voxels, per-class counts, per-slice counts along the last axis)."""
    cls, sl = ops.voxel_counts(mask)
    cls, sl = cls.cpu().tolist(), sl.cpu().tolist()
    return {"tumor_voxels": int(sum(cls[1:])), "class_voxels": cls, "slice_voxels": sl}


CLASS_NAMES = ("Background", "Necrotic Core", "Peritumoral Edema", "Enhancing Tumor")   # main.py:417


def classify(model, volume):
    """classify_tumor call site (main.py:408-418): eval forward of BrainTumorClassifier, softmax, arg-max and its probability.
    Returns (predicted class [N] int64, confidence [N] fp32, probabilities [N,K]) on the device — one D2H copy for the caller."""
    was_training = model.training
    model.eval()
    with torch.no_grad():
        prob = torch.softmax(model(volume), dim=1)
    if was_training:
        model.train()
    conf, pred = prob.max(dim=1)
    return pred, conf, prob
