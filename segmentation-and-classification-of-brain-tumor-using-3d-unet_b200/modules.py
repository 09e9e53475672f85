"""nn.Module surface of the reference (drop-in): UNet3D, DoubleConv3D, AttentionGate3D.

Same constructors, attribute names, parameter registration order (hence identical RNG consumption and initial values
for a given seed) and state_dict keys as /root/reference/main.py:102-299, so reference checkpoints load both ways and
the reference call sites (train_model.py:174, training.py:294-304, main.py:390-393) work unchanged.  The sub-modules are
PARAMETER CONTAINERS only: `forward` never calls them, it runs the hand-written sm_100a kernels of libb3d.so through
functional.py.  There is no fallback — on a non-B200 device (or CPU tensors) forward raises.
"""
import torch
import torch.nn as nn

from . import _lib, functional as Fn, ops
from .lazy import LazyDeepOutput


def _conv(cin, cout, k, bias=True):
    return nn.Conv3d(cin, cout, kernel_size=k, stride=1, padding=k // 2, bias=bias)


def _require_cuda(x, who):
    if not x.is_cuda:
        raise _lib.B3DError("%s: CUDA (sm_100) tensors required — the b200 path has no CPU fallback" % who)
    _lib.require_device(x.device)


def _param_dict(module, prefix=""):
    return {prefix + k: v for k, v in module.named_parameters()}


class _BlockFn(torch.autograd.Function):
    """Generic autograd bridge: fwd(ctx_args) -> outputs, bwd(saved, grads) -> (input grads..., param grads dict)."""

    @staticmethod
    def forward(ctx, impl, n_inputs, names, *tensors):
        inputs, params = tensors[:n_inputs], tensors[n_inputs:]
        p = dict(zip(names, params))
        need_bwd = any(ctx.needs_input_grad)  # grad mode is off inside Function.forward; this reflects the caller's
        ctx.set_materialize_grads(False)  # unused outputs (e.g. the 4th deep-supervision map) arrive as None
        outs, saved = impl.fwd(inputs, p, need_bwd)
        ctx.impl, ctx.saved_state, ctx.names, ctx.params, ctx.n_inputs = impl, saved, names, p, n_inputs
        ctx.mark_non_differentiable(*[o for o in outs if not o.is_floating_point()])
        return tuple(outs)

    @staticmethod
    def backward(ctx, *gouts):
        if ctx.saved_state is None:
            raise RuntimeError("backward through a forward that ran without grad enabled")
        din, grads = ctx.impl.bwd(ctx.saved_state, gouts, ctx.params)
        ctx.saved_state = None
        pg = []
        for name in ctx.names:
            g = grads.get(name)
            if g is not None:
                g = g.reshape(ctx.params[name].shape)
            pg.append(g)
        return (None, None, None) + tuple(din) + tuple(pg)


def _as_ndhwc(x, cpad=None):
    """fp32/bf16 NCDHW module input -> NDHWC bf16 activation (channels zero-padded to a multiple of 16)."""
    c = x.shape[1]
    return ops.to_ndhwc_bf16(x.detach(), ops.roundup(c, 16) if cpad is None else cpad)


class _DoubleConvImpl:
    def __init__(self, cin, cout):
        self.cin, self.cout = cin, cout

    def fwd(self, inputs, p, need_bwd):
        x = _as_ndhwc(inputs[0])
        out, saved = Fn.double_conv_fwd(x, p, "", self.cin, need_bwd)
        return [ops.to_ncdhw_f32(out)], saved

    def bwd(self, saved, gouts, p):
        dout = _as_ndhwc(gouts[0].contiguous())
        dx, grads = Fn.double_conv_bwd(saved, dout, p, "", need_dx=True)
        ops.wgrad_join()
        return [ops.to_ncdhw_f32(dx)[:, :self.cin]], grads


class DoubleConv3D(nn.Module):
    """(Conv3x3x3 -> GroupNorm(8) -> ReLU) x 2 plus a GroupNorm'd 1x1x1 residual projection — main.py:205-242."""

    def __init__(self, in_channels, out_channels, mid_channels=None, use_residual=True):
        super().__init__()
        if not mid_channels:
            mid_channels = out_channels
        if mid_channels != out_channels:
            raise NotImplementedError("b200 DoubleConv3D: mid_channels != out_channels is not used by UNet3D")
        self.in_channels, self.out_channels = in_channels, out_channels
        self.use_residual = use_residual and (in_channels == out_channels)
        self.double_conv = nn.Sequential(
            _conv(in_channels, mid_channels, 3, bias=False), nn.GroupNorm(8, mid_channels), nn.ReLU(inplace=True),
            _conv(mid_channels, out_channels, 3, bias=False), nn.GroupNorm(8, out_channels), nn.ReLU(inplace=True))
        if self.use_residual:
            self.residual = nn.Identity()
        elif in_channels != out_channels:
            self.residual = nn.Sequential(_conv(in_channels, out_channels, 1, bias=False), nn.GroupNorm(8, out_channels))
        else:
            self.residual = None
            raise NotImplementedError("b200 DoubleConv3D: use_residual=False with in == out is not used by UNet3D")

    def forward(self, x):
        _require_cuda(x, "DoubleConv3D")
        names, params = zip(*self.named_parameters())
        return _BlockFn.apply(_DoubleConvImpl(self.in_channels, self.out_channels), 1, names, x, *params)[0]


class _GateImpl:
    def __init__(self, c):
        self.c = c

    def fwd(self, inputs, p, need_bwd):
        g, x = _as_ndhwc(inputs[0]), _as_ndhwc(inputs[1])
        out = torch.empty_like(x)
        _, saved = Fn.gate_fwd(g, x, p, "", out, need_bwd)
        return [ops.to_ncdhw_f32(out)], saved

    def bwd(self, saved, gouts, p):
        dout = _as_ndhwc(gouts[0].contiguous())
        dg, dx, grads = Fn.gate_bwd(saved, dout, p, "")
        ops.wgrad_join()
        return [ops.to_ncdhw_f32(dg), ops.to_ncdhw_f32(dx)], grads


class AttentionGate3D(nn.Module):
    """Spatial attention gate with a squeeze-excite channel-attention branch — main.py:244-299."""

    def __init__(self, F_g, F_l, F_int):
        super().__init__()
        if F_g != F_l:
            raise NotImplementedError("b200 AttentionGate3D: F_g must equal F_l (as in UNet3D)")
        self.W_g = nn.Sequential(_conv(F_g, F_int, 1), nn.GroupNorm(4, F_int))
        self.W_x = nn.Sequential(_conv(F_l, F_int, 1), nn.GroupNorm(4, F_int))
        self.psi = nn.Sequential(_conv(F_int, 1, 1), nn.GroupNorm(1, 1), nn.Sigmoid())
        self.channel_attention = nn.Sequential(nn.AdaptiveAvgPool3d(1), _conv(F_l, F_l // 8, 1), nn.ReLU(inplace=True),
                                               _conv(F_l // 8, F_l, 1), nn.Sigmoid())
        self.relu = nn.ReLU(inplace=True)
        self.F_l = F_l

    def forward(self, g, x):
        _require_cuda(x, "AttentionGate3D")
        if g.shape != x.shape:
            raise ValueError("b200 AttentionGate3D: g and x must have identical shapes (no interpolate fix-up)")
        names, params = zip(*self.named_parameters())
        return _BlockFn.apply(_GateImpl(self.F_l), 2, names, g, x, *params)[0]


class _UNetImpl:
    def __init__(self, model):
        self.model = model
        self.on_grads = None

    def fwd(self, inputs, p, need_bwd):
        m = self.model
        x = inputs[0]
        training = m.training
        masks = None
        if training and m.dropout.p > 0:
            # identical Bernoulli stream to nn.Dropout3d: one [N,C,1,1,1] draw per encoder level, in order
            # (F.dropout3d noise = x.new_empty(N,C,1,1,1).bernoulli_(1-p).div_(1-p); SURVEY hard part 6)
            pr = m.dropout.p
            masks = [x.new_empty((x.shape[0], f, 1, 1, 1), dtype=torch.float32).bernoulli_(1 - pr).div_(1 - pr)
                     .reshape(x.shape[0], f).contiguous() for f in m.features]
            m._last_dropout_masks = masks  # kept for inspection / parity tests (tiny)
        bufs = dict(m.named_buffers())
        logits, deep, saved = Fn.unet_fwd(x.detach(), p, bufs, list(m.features), training, masks, need_bwd)
        return [logits] + deep, saved

    def bwd(self, saved, gouts, p):
        dmain = gouts[0]
        if dmain is None:  # only deep outputs fed the loss
            x, hbuf = saved["final"][0], saved["final"][1]
            k = p["final_conv.3.weight"].shape[0]
            dmain = torch.zeros((x.shape[0], k) + tuple(x.shape[1:4]), dtype=torch.float32, device=x.device)
        Fn.GRAD_SINK = self.model._grad_sink   # data parallel: weight-gradient kernels write into the all-reduce buckets
        try:
            grads = Fn.unet_bwd(saved, dmain, list(gouts[1:]), p, list(self.model.features), on_grads=self.model._on_grads)
        finally:
            Fn.GRAD_SINK = None
        ops.wgrad_join()
        if self.model._on_backward_end is not None:
            self.model._on_backward_end()  # data parallel: join the gradient all-reduces before autograd sees the grads
        if self.model._grad_views is not None:
            grads = self.model._grad_views(grads)
        return [None], grads


class UNet3D(nn.Module):
    """Enhanced 3D U-Net (residual double convs, attention gates, deep supervision) — main.py:102-203.

    forward(x: float[N,Cin,D,H,W]) -> Tensor[N,Cout,D,H,W] in eval, (Tensor, [4 Tensors]) in train (main.py:200-203).
    D, H, W must be multiples of 32 (the reference's interpolate fix-ups for other sizes are not reproduced).
    """

    def __init__(self, in_channels=1, out_channels=4, features=[32, 64, 128, 256, 512], dropout_rate=0.2):
        super().__init__()
        self.ups = nn.ModuleList()
        self.downs = nn.ModuleList()
        self.pool = nn.MaxPool3d(kernel_size=2, stride=2)
        self.dropout = nn.Dropout3d(dropout_rate)
        self.features = features
        cin = in_channels
        for f in features:
            self.downs.append(DoubleConv3D(cin, f))
            cin = f
        for f in reversed(features):
            self.ups.append(nn.ConvTranspose3d(f * 2, f, kernel_size=2, stride=2))
            self.ups.append(AttentionGate3D(f, f, f // 2))
            self.ups.append(DoubleConv3D(f * 2, f))
        self.bottleneck = DoubleConv3D(features[-1], features[-1] * 2)
        self.final_conv = nn.Sequential(_conv(features[0], features[0] // 2, 3), nn.BatchNorm3d(features[0] // 2),
                                        nn.ReLU(inplace=True), _conv(features[0] // 2, out_channels, 1))
        self.deep_supervision = nn.ModuleList([_conv(f, out_channels, 1) for f in features[:-1]])
        self.apply(self._init_weights)
        self._on_grads = None  # hooks for data-parallel gradient bucketing (parallel.py)
        self._on_backward_end = None
        self._grad_sink = None
        self._grad_views = None
        for a, b in zip(features[:-1], features[1:]):
            if b != 2 * a:
                raise ValueError("UNet3D: features must double at every level (the reference decoder requires it)")

    def _init_weights(self, m):
        if isinstance(m, nn.Conv3d):
            nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)
        elif isinstance(m, nn.BatchNorm3d):
            nn.init.constant_(m.weight, 1)
            nn.init.constant_(m.bias, 0)

    def forward(self, x):
        _require_cuda(x, "UNet3D")
        names, params = zip(*self.named_parameters())
        outs = _BlockFn.apply(_UNetImpl(self), 1, names, x, *params)
        if self.training and len(outs) > 1:
            # main.py:200-201 returns the up-sampled head maps; here they are lazy views of the low-res head logits (lazy.py)
            size = tuple(x.shape[2:])
            return outs[0], [LazyDeepOutput(lo, size) for lo in outs[1:]]
        return outs[0]


class BrainTumorClassifier(nn.Module):
    """CNN tumour-type classifier — main.py:301-328, used by `classify_tumor` (main.py:398-425) in eval mode under no_grad.

    Same layers / state_dict keys as the reference (`features.{0,3,6}`, `classifier.{0,3}`).  The three 3x3x3 convolutions run
    on the tcgen05 implicit-GEMM kernels, ReLU+MaxPool and ReLU+AdaptiveAvgPool on the bandwidth kernels of pool_layout.cu;
    the two Linear layers (8192x512, 512xK: 8 MFLOP per sample) are plain library GEMMs.  Inference only: the reference never
    trains this model, and train mode (Dropout(0.5) active) or a forward that needs gradients raises.  D, H, W must be
    multiples of 4 (the two MaxPool3d(2) stages; nn.MaxPool3d would floor other sizes)."""

    def __init__(self, num_classes=4):
        super().__init__()
        self.features = nn.Sequential(
            nn.Conv3d(4, 32, 3, 1, 1), nn.ReLU(), nn.MaxPool3d(2),
            nn.Conv3d(32, 64, 3, 1, 1), nn.ReLU(), nn.MaxPool3d(2),
            nn.Conv3d(64, 128, 3, 1, 1), nn.ReLU(), nn.AdaptiveAvgPool3d((4, 4, 4)))
        self.classifier = nn.Sequential(nn.Linear(128 * 4 * 4 * 4, 512), nn.ReLU(), nn.Dropout(0.5), nn.Linear(512, num_classes))

    def forward(self, x):
        _require_cuda(x, "BrainTumorClassifier")
        if self.training:
            raise NotImplementedError("b200 BrainTumorClassifier: inference only (call .eval(); the reference never trains it)")
        n, cin, d, h, w = x.shape
        if cin != 4:
            raise ValueError("BrainTumorClassifier expects 4 input channels, got %d" % cin)
        if d % 4 or h % 4 or w % 4 or min(d, h, w) < 16:
            raise ValueError("BrainTumorClassifier (b200 path): D,H,W must be multiples of 4 and >= 16, got %s" % ((d, h, w),))
        with torch.no_grad():
            a = ops.to_ndhwc_bf16(x.detach(), 16)
            for idx, last in ((0, False), (3, False), (6, True)):
                conv = self.features[idx]
                wp, _, rows = Fn.packed(conv.weight, ops.PACK_FPROP)
                a, _ = ops.conv_fprop(a, wp, rows, conv.out_channels, 3, bias=conv.bias)
                a = ops.relu_adaptive_avgpool(a, (4, 4, 4)) if last else ops.relu_pool_fwd(a)
            fc1, fc2 = self.classifier[0], self.classifier[3]
            hdn = torch.relu(torch.nn.functional.linear(a, fc1.weight, fc1.bias))
            return torch.nn.functional.linear(hdn, fc2.weight, fc2.bias)
