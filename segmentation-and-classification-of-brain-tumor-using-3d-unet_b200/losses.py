"""Loss classes with the reference's names, constructors and return types, computed by the fused sm_100a loss kernels.

    CombinedLoss3D / TverskyLoss3D / DeepSupervisionLoss3D      /root/reference/losses.py:7-126
    CombinedLoss / DiceLoss / FocalLoss (trainer-local)          /root/reference/training.py:517-566
"""
import os

import torch
import torch.nn as nn

from . import _lib, ops
from .lazy import LazyDeepOutput


class _LossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, target, cfg):
        if not logits.is_cuda:
            raise _lib.B3DError("loss: CUDA (sm_100) tensors required — the b200 path has no CPU fallback")
        _lib.require_device(logits.device)
        lg = logits.detach()
        if lg.dtype != torch.float32:
            lg = lg.float()
        values, saved = ops.loss_fwd(lg, target, cfg)
        ctx.saved_state, ctx.cfg, ctx.shape, ctx.in_dtype = saved, cfg, tuple(logits.shape), logits.dtype
        ctx.mark_non_differentiable(values)
        total = values[0].clone()
        return total, values

    @staticmethod
    def backward(ctx, gtotal, _gvalues):
        g = gtotal.contiguous().float().reshape(1)
        dl = ops.loss_bwd(ctx.saved_state, ctx.cfg, g, 1.0, ctx.shape)
        ctx.saved_state = None
        return dl.to(ctx.in_dtype), None, None


def _run(pred, target, cfg):
    return _LossFn.apply(pred, target, cfg)


class _DsLossFn(torch.autograd.Function):
    """Fused F.interpolate(trilinear) + loss of one deep-supervision output, straight from its low-res logits (dsloss.cu)."""

    @staticmethod
    def forward(ctx, lo, tgt_u8, cfg, size):
        _lib.require_device(lo.device)
        lod = lo.detach().contiguous()
        values, acc = ops.dsloss_fwd(lod, tgt_u8, cfg, size)
        ctx.saved_state, ctx.cfg, ctx.size = (lod, tgt_u8, acc), cfg, size
        ctx.mark_non_differentiable(values)
        return values[0].clone(), values

    @staticmethod
    def backward(ctx, gtotal, _gvalues):
        lod, tgt_u8, acc = ctx.saved_state
        g = gtotal.contiguous().float().reshape(1)
        dlo = ops.dsloss_bwd(lod, tgt_u8, acc, ctx.cfg, g, 1.0, ctx.size)
        ctx.saved_state = None
        return dlo, None, None, None


class CombinedLoss3D(nn.Module):
    """alpha*dice + beta*focal(0.25, 2) + gamma*boundary — losses.py:7-75.  forward -> (tensor, dict of 4 floats)."""

    def __init__(self, alpha=0.5, beta=0.3, gamma=0.2, smooth=1e-5):
        super().__init__()
        self.alpha, self.beta, self.gamma, self.smooth = alpha, beta, gamma, smooth

    def _cfg(self):
        return ops.loss_cfg(w_dice=self.alpha, smooth=self.smooth, w_focal=self.beta, f_alpha=0.25, f_gamma=2.0,
                            w_boundary=self.gamma)

    def loss_tensor(self, pred, target):
        """Total loss only — no host synchronisation (used by DeepSupervisionLoss3D)."""
        return _run(pred, target, self._cfg())[0]

    def dice_loss(self, pred, target):
        return _run(pred, target, ops.loss_cfg(w_dice=1.0, smooth=self.smooth))[0]

    def focal_loss(self, pred, target, alpha=0.25, gamma=2.0):
        return _run(pred, target, ops.loss_cfg(w_focal=1.0, f_alpha=alpha, f_gamma=gamma))[0]

    def boundary_loss(self, pred, target):
        return _run(pred, target, ops.loss_cfg(w_boundary=1.0))[0]

    def forward(self, pred, target):
        total, values = _run(pred, target, self._cfg())
        v = values.tolist()  # ONE device->host copy instead of the reference's four .item() calls
        return total, {"dice_loss": v[1], "focal_loss": v[2], "boundary_loss": v[3], "total_loss": v[0]}


class TverskyLoss3D(nn.Module):
    """losses.py:77-97."""

    def __init__(self, alpha=0.7, beta=0.3, smooth=1e-5):
        super().__init__()
        self.alpha, self.beta, self.smooth = alpha, beta, smooth

    def forward(self, pred, target):
        return _run(pred, target, ops.loss_cfg(w_tv=1.0, tv_alpha=self.alpha, tv_beta=self.beta, tv_smooth=self.smooth))[0]


FUSED_DS_LOSS = os.environ.get("B3D_FUSED_DS_LOSS", "1") != "0"   # 0: materialise the up-sampled maps (round-1 path)


class DeepSupervisionLoss3D(nn.Module):
    """w0*L(main) + sum_{i<len(w)-1} w_{i+1}*L(deep_i) — losses.py:99-126 (the 4th deep output is unused, as there)."""

    def __init__(self, weights=[1.0, 0.8, 0.6, 0.4], loss_fn=None):
        super().__init__()
        self.weights = weights
        self.loss_fn = loss_fn or CombinedLoss3D()

    def _one(self, pred, target, tgt_u8=None):
        if isinstance(pred, LazyDeepOutput) and tgt_u8 is not None:   # fused upsample + loss from the low-res head logits
            return _DsLossFn.apply(pred.lo, tgt_u8, self.loss_fn._cfg(), pred.full_size)[0]
        if hasattr(self.loss_fn, "loss_tensor"):
            return self.loss_fn.loss_tensor(pred, target)
        out = self.loss_fn(pred, target)
        return out[0] if isinstance(out, tuple) else out

    def forward(self, predictions, target):
        if isinstance(predictions, tuple):
            main_pred, deep_preds = predictions
            total = self._one(main_pred, target) * self.weights[0]
            used = [pred for i, pred in enumerate(deep_preds) if i < len(self.weights) - 1]
            fused = isinstance(self.loss_fn, CombinedLoss3D) and FUSED_DS_LOSS and any(isinstance(q, LazyDeepOutput) for q in used)
            tgt_u8 = ops.target_u8(target) if fused else None
            for i, pred in enumerate(used):
                if tuple(pred.shape[2:]) != tuple(target.shape[1:]):
                    raise ValueError("b200 DeepSupervisionLoss3D: deep outputs must be full resolution (UNet3D "
                                     "up-samples them, main.py:165-170)")
                total = total + self._one(pred, target, tgt_u8) * self.weights[i + 1]
            return total
        return self._one(predictions, target)


class DiceLoss(nn.Module):
    """training.py:536-553."""

    def __init__(self, smooth=1e-6):
        super().__init__()
        self.smooth = smooth

    def forward(self, outputs, targets):
        return _run(outputs, targets, ops.loss_cfg(w_dice=1.0, smooth=self.smooth))[0]


class FocalLoss(nn.Module):
    """training.py:555-566."""

    def __init__(self, alpha=1, gamma=2):
        super().__init__()
        self.alpha, self.gamma = alpha, gamma

    def forward(self, outputs, targets):
        return _run(outputs, targets, ops.loss_cfg(w_focal=1.0, f_alpha=float(self.alpha), f_gamma=float(self.gamma)))[0]


class CombinedLoss(nn.Module):
    """0.5*Dice(1e-6) + 0.3*CrossEntropy + 0.2*Focal(1,2) — training.py:517-534 (tensor in, tensor out)."""

    def __init__(self, weights=[0.5, 0.3, 0.2]):
        super().__init__()
        self.weights = weights
        self.dice_loss = DiceLoss()
        self.ce_loss = nn.CrossEntropyLoss()  # attribute kept for API parity; the fused kernel computes the CE term
        self.focal_loss = FocalLoss()

    def forward(self, outputs, targets):
        cfg = ops.loss_cfg(w_dice=self.weights[0], smooth=self.dice_loss.smooth, w_ce=self.weights[1],
                           w_focal=self.weights[2], f_alpha=float(self.focal_loss.alpha),
                           f_gamma=float(self.focal_loss.gamma))
        return _run(outputs, targets, cfg)[0]
