"""Optimizer glue of the train step: AdamW exactly as the reference trainer builds it
(/root/reference/training.py:186-191: lr, weight_decay=1e-4, betas=(0.9, 0.999)) with `CosineAnnealingWarmRestarts`
(training.py:194-196, stepped per epoch at :252).

`FusedAdamW` is a `torch.optim.Optimizer` (same `param_groups` / `state_dict` layout as `torch.optim.AdamW`, so reference
checkpoints' optimizer states load) whose `step()` is TWO kernel launches (csrc/pack.cu):
  * one over every conv weight the tcgen05 kernels read: p, g, m, v are read once, AdamW is applied in fp32, p / m / v are
    written back and BOTH bf16 packed operand layouts (fprop + dgrad) are emitted from shared memory in the same pass — the
    separate re-pack pass (and its ~50 launches) after the optimizer disappears;
  * one multi-tensor launch over every other parameter.
The learning rate and the step count live in device memory, so LR schedulers keep working when the step is replayed from a
CUDA graph (graph.py).  torch's own `AdamW(fused=True)` remains fully supported (the packed copies are then refreshed by
`functional.packed` after each step).
"""
import ctypes

import torch
import torch.nn as nn

from . import _lib, ops
from ._lib import c_int, c_ll, check, ptr, stream_ptr

c_double = ctypes.c_double
FLAT_CHUNK = 4096   # elements per CTA of the flat kernel (ADAM_FLAT_CHUNK in pack.cu)


def packed_conv_params(model):
    """{id(param): (param, is_conv_transpose)} for every weight the conv kernels read through `functional.packed`."""
    from . import modules as M
    out = {}

    def add(w, convt=False):
        out[id(w)] = (w, convt)
    for m in model.modules():
        if isinstance(m, M.DoubleConv3D):
            add(m.double_conv[0].weight)
            add(m.double_conv[3].weight)
            if isinstance(m.residual, nn.Sequential):
                add(m.residual[0].weight)
        elif isinstance(m, M.AttentionGate3D):
            add(m.W_g[0].weight)
            add(m.W_x[0].weight)
        elif isinstance(m, nn.ConvTranspose3d):
            add(m.weight, True)
        elif isinstance(m, M.UNet3D):
            add(m.final_conv[0].weight)
        elif isinstance(m, M.BrainTumorClassifier):
            for idx in (0, 3, 6):
                add(m.features[idx].weight)
    return out


class FusedAdamW(torch.optim.Optimizer):
    """torch.optim.AdamW semantics (decoupled weight decay, bias correction) in two launches, re-pack included."""

    def __init__(self, model, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2):
        if not isinstance(model, nn.Module):
            raise TypeError("FusedAdamW takes the nn.Module (it needs to know which weights the conv kernels read)")
        params = [p for p in model.parameters() if p.requires_grad]
        if not params or not params[0].is_cuda:
            raise _lib.B3DError("FusedAdamW: parameters must live on a CUDA (sm_100) device — no CPU fallback")
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, amsgrad=False, maximize=False, foreach=None,
                        capturable=True, differentiable=False, fused=True)
        super().__init__(params, defaults)
        self._packed = packed_conv_params(model)
        dev = params[0].device
        self._dev = dev
        self._step_t = torch.zeros((), dtype=torch.float32, device=dev)
        self._lr_t = {}
        self._tables = {}      # group index -> (key, (host tables, counts))

    # -- state -----------------------------------------------------------------------------------------------------------
    def _init_state(self, p):
        st = self.state[p]
        if "exp_avg" not in st:
            st["step"] = self._step_t
            st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
        if id(p) in self._packed and "_b3d_bufs" not in st:
            w, convt = self._packed[id(p)]
            if convt:
                cin, cout = w.shape[0], w.shape[1]
                ri = ops.roundup(cin, 16)
                of = torch.empty(8 * cout * ri, dtype=torch.bfloat16, device=w.device)
                od = torch.empty(ri * 8 * cout, dtype=torch.bfloat16, device=w.device)
                st["_b3d_bufs"] = {ops.PACK_CONVT_FPROP: (of, ri, 8 * cout), ops.PACK_CONVT_DGRAD: (od, 8 * cout, ri)}
                st["_b3d_geom"] = (cin, cout, 8, 1, ri, 8 * cout, 8 * cout, ri)          # A, B, T, convT, Kp_f, rows_f, Kp_d, rows_d
            else:
                cout, cin = w.shape[0], w.shape[1]
                t = w.shape[2] * w.shape[3] * w.shape[4]
                ri, ro = ops.roundup(cin, 16), ops.roundup(cout, 16)
                # padded rows / K columns of the packed buffers are never written by the tile kernel: they must be zero
                of = torch.zeros(t * ro * ri, dtype=torch.bfloat16, device=w.device)
                od = torch.zeros(t * ri * ro, dtype=torch.bfloat16, device=w.device)
                st["_b3d_bufs"] = {ops.PACK_FPROP: (of, ri, ro), ops.PACK_DGRAD: (od, ro, ri)}
                st["_b3d_geom"] = (cout, cin, t, 0, ri, ro, ro, ri)
        return st

    def state_dict(self):
        sd = super().state_dict()
        for st in sd["state"].values():   # private buffers are derived data: keep the checkpoint torch.optim.AdamW-compatible
            st.pop("_b3d_bufs", None)
            st.pop("_b3d_geom", None)
            if torch.is_tensor(st.get("step")):
                st["step"] = st["step"].detach().clone()
        return sd

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        steps = [float(st["step"]) for st in self.state.values() if "step" in st]
        self._step_t.fill_(max(steps) if steps else 0.0)
        for p, st in self.state.items():
            st["step"] = self._step_t
            for k in ("exp_avg", "exp_avg_sq"):
                if k in st and (st[k].device != p.device or st[k].dtype != torch.float32):
                    st[k] = st[k].to(device=p.device, dtype=torch.float32)
        self._tables.clear()

    # -- step ------------------------------------------------------------------------------------------------------------
    def _lr_tensor(self, gi, group):
        lr = group["lr"]
        if torch.is_tensor(lr):
            if not lr.is_cuda:
                raise _lib.B3DError("FusedAdamW: a tensor lr must live on the device")
            return lr if lr.dtype == torch.float32 else lr.float()
        t = self._lr_t.get(gi)
        if t is None:
            t = self._lr_t[gi] = torch.empty((), dtype=torch.float32, device=self._dev)
        t.fill_(float(lr))   # NOTE: baked into a CUDA-graph capture — GraphedTrainStep converts the lr to a tensor first
        return t

    def _build_tables(self, gi, plist):
        pack_rows, flat_rows = [], []
        tiles = blocks = 0
        for p in plist:
            st = self._init_state(p)
            g = p.grad
            if g.dtype != torch.float32 or not g.is_contiguous() or not p.is_contiguous():
                raise _lib.B3DError("FusedAdamW: parameters and gradients must be contiguous fp32")
            if "_b3d_bufs" in st:
                a, b, t, convt, kpf, rf, kpd, rd = st["_b3d_geom"]
                bufs = list(st["_b3d_bufs"].values())
                ta, tb = (a + 15) // 16, (b + 15) // 16
                pack_rows.append([p.data_ptr(), g.data_ptr(), st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr(),
                                  bufs[0][0].data_ptr(), bufs[1][0].data_ptr(), a, b, t, convt, kpf, rf, kpd, rd, tiles, tb])
                tiles += ta * tb
            else:
                n = p.numel()
                flat_rows.append([p.data_ptr(), g.data_ptr(), st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr(), n, blocks, 0, 0])
                blocks += (n + FLAT_CHUNK - 1) // FLAT_CHUNK
        # HOST tables: the C side copies them into kernel parameters (no device copy, graph-capture safe)
        host_p = torch.tensor(pack_rows if pack_rows else [[0] * 16], dtype=torch.int64)
        host_f = torch.tensor(flat_rows if flat_rows else [[0] * 8], dtype=torch.int64)
        return (host_p, host_f, len(pack_rows), tiles, len(flat_rows), blocks)

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        _lib.require_device(self._dev)
        self._step_t.add_(1.0)
        for gi, group in enumerate(self.param_groups):
            plist = [p for p in group["params"] if p.grad is not None]
            if not plist:
                continue
            key = tuple((p.data_ptr(), p.grad.data_ptr()) for p in plist)
            cached = self._tables.get(gi)
            if cached is None or cached[0] != key:
                cached = (key, self._build_tables(gi, plist))
                self._tables[gi] = cached
            host_p, host_f, n_pack, tiles, n_flat, blocks = cached[1]
            beta1, beta2 = group["betas"]
            check(_lib.lib().b3d_adamw_step(ptr(host_p), c_int(n_pack), c_ll(tiles), ptr(host_f), c_int(n_flat),
                                            c_ll(blocks), ptr(self._lr_tensor(gi, group)), ptr(self._step_t), c_double(beta1),
                                            c_double(beta2), c_double(group["eps"]), c_double(group["weight_decay"]), stream_ptr()))
            for p in plist:   # the packed copies written by the kernel ARE the current ones: functional.packed returns them
                st = self.state[p]
                if "_b3d_bufs" in st:
                    p.__dict__["_b3d_pack_pinned"] = (p._version, p.data_ptr(), st["_b3d_bufs"])
        return loss


def make_adamw(model, lr=1e-4, weight_decay=1e-4, betas=(0.9, 0.999), eps=1e-8, capturable=True, fused_pack=True):
    """The trainer's optimizer (training.py:186-191).  On the device: `FusedAdamW` with a device-scalar learning rate;
    `fused_pack=False` gives torch's own fused AdamW (the round-1 path)."""
    params = [p for p in model.parameters() if p.requires_grad]
    dev = params[0].device
    if dev.type != "cuda":
        return torch.optim.AdamW(params, lr=lr, weight_decay=weight_decay, betas=betas, eps=eps)
    lr_t = torch.tensor(float(lr), dtype=torch.float32, device=dev) if capturable else lr
    if fused_pack:
        return FusedAdamW(model, lr=lr_t, betas=betas, eps=eps, weight_decay=weight_decay)
    return torch.optim.AdamW(params, lr=lr_t, weight_decay=weight_decay, betas=betas, eps=eps, fused=True, capturable=capturable)


def make_scheduler(optimizer, T_0=10, T_mult=2, eta_min=1e-6):
    """training.py:194-196."""
    return torch.optim.lr_scheduler.CosineAnnealingWarmRestarts(optimizer, T_0=T_0, T_mult=T_mult, eta_min=eta_min)
