"""Optimizer glue of the train step: AdamW exactly as the reference trainer builds it
(/root/reference/training.py:186-191: lr, weight_decay=1e-4, betas=(0.9, 0.999)) with `CosineAnnealingWarmRestarts`
(training.py:194-196, stepped per epoch at :252).

`make_adamw` returns an optimizer whose learning rate is a DEVICE scalar, so the schedule keeps working when the step is
replayed from a CUDA graph (graph.py).
"""
import torch


def make_adamw(model, lr=1e-4, weight_decay=1e-4, betas=(0.9, 0.999), eps=1e-8, capturable=True):
    params = [p for p in model.parameters() if p.requires_grad]
    dev = params[0].device
    lr_t = torch.tensor(float(lr), dtype=torch.float32, device=dev) if (capturable and dev.type == "cuda") else lr
    return torch.optim.AdamW(params, lr=lr_t, weight_decay=weight_decay, betas=betas, eps=eps, fused=dev.type == "cuda",
                             capturable=capturable and dev.type == "cuda")


def make_scheduler(optimizer, T_0=10, T_mult=2, eta_min=1e-6):
    """training.py:194-196."""
    return torch.optim.lr_scheduler.CosineAnnealingWarmRestarts(optimizer, T_0=T_0, T_mult=T_mult, eta_min=eta_min)
