"""Functional CPU restatement of the reference hot path (TEST INFRASTRUCTURE — see oracle/__init__.py).

Each function cites the reference lines it follows.  Parameters are taken from a plain state_dict with the
reference's key names, so the same weights drive the reference, this oracle and the CUDA path.
"""
import math

import torch
import torch.nn.functional as F


# ----------------------------------------------------------------------------------------------
# Parameter inventory (reference: main.py:105-152 module tree; state_dict order = registration order)
# ----------------------------------------------------------------------------------------------
def _double_conv_shapes(prefix, cin, cout):
    """main.py:208-233 — DoubleConv3D: conv3(no bias) GN8 ReLU conv3(no bias) GN8 ReLU + residual conv1(no bias) GN8."""
    s = [
        (prefix + "double_conv.0.weight", (cout, cin, 3, 3, 3)),
        (prefix + "double_conv.1.weight", (cout,)), (prefix + "double_conv.1.bias", (cout,)),
        (prefix + "double_conv.3.weight", (cout, cout, 3, 3, 3)),
        (prefix + "double_conv.4.weight", (cout,)), (prefix + "double_conv.4.bias", (cout,)),
    ]
    if cin != cout:
        s += [(prefix + "residual.0.weight", (cout, cin, 1, 1, 1)),
              (prefix + "residual.1.weight", (cout,)), (prefix + "residual.1.bias", (cout,))]
    return s


def _gate_shapes(prefix, fg, fl, fint):
    """main.py:247-278 — AttentionGate3D."""
    return [
        (prefix + "W_g.0.weight", (fint, fg, 1, 1, 1)), (prefix + "W_g.0.bias", (fint,)),
        (prefix + "W_g.1.weight", (fint,)), (prefix + "W_g.1.bias", (fint,)),
        (prefix + "W_x.0.weight", (fint, fl, 1, 1, 1)), (prefix + "W_x.0.bias", (fint,)),
        (prefix + "W_x.1.weight", (fint,)), (prefix + "W_x.1.bias", (fint,)),
        (prefix + "psi.0.weight", (1, fint, 1, 1, 1)), (prefix + "psi.0.bias", (1,)),
        (prefix + "psi.1.weight", (1,)), (prefix + "psi.1.bias", (1,)),
        (prefix + "channel_attention.1.weight", (fl // 8, fl, 1, 1, 1)), (prefix + "channel_attention.1.bias", (fl // 8,)),
        (prefix + "channel_attention.3.weight", (fl, fl // 8, 1, 1, 1)), (prefix + "channel_attention.3.bias", (fl,)),
    ]


def param_shapes(in_channels=1, out_channels=4, features=(32, 64, 128, 256, 512)):
    """Ordered (key, shape) list of UNet3D.state_dict() (main.py:105-140).  Buffers of BatchNorm3d included."""
    features = list(features)
    out = []
    # registration order in __init__: ups, downs, (pool, dropout), bottleneck, final_conv, deep_supervision
    ups = []
    for i, f in enumerate(reversed(features)):
        ups += [("ups.%d.weight" % (3 * i), (2 * f, f, 2, 2, 2)), ("ups.%d.bias" % (3 * i), (f,))]
        ups += _gate_shapes("ups.%d." % (3 * i + 1), f, f, f // 2)
        ups += _double_conv_shapes("ups.%d." % (3 * i + 2), 2 * f, f)
    out += ups
    cin = in_channels
    for i, f in enumerate(features):
        out += _double_conv_shapes("downs.%d." % i, cin, f)
        cin = f
    out += _double_conv_shapes("bottleneck.", features[-1], 2 * features[-1])
    f0 = features[0]
    out += [
        ("final_conv.0.weight", (f0 // 2, f0, 3, 3, 3)), ("final_conv.0.bias", (f0 // 2,)),
        ("final_conv.1.weight", (f0 // 2,)), ("final_conv.1.bias", (f0 // 2,)),
        ("final_conv.1.running_mean", (f0 // 2,)), ("final_conv.1.running_var", (f0 // 2,)),
        ("final_conv.1.num_batches_tracked", ()),
        ("final_conv.3.weight", (out_channels, f0 // 2, 1, 1, 1)), ("final_conv.3.bias", (out_channels,)),
    ]
    for i, f in enumerate(features[:-1]):
        out += [("deep_supervision.%d.weight" % i, (out_channels, f, 1, 1, 1)),
                ("deep_supervision.%d.bias" % i, (out_channels,))]
    return out


def make_state_dict(in_channels=1, out_channels=4, features=(32, 64, 128, 256, 512), seed=0, dtype=torch.float32):
    """Deterministic, reference-independent weights for parity tests: every tensor is drawn from its own generator
    seeded by (seed, key index), so the recipe is reproducible on any box.  Conv weights ~ Kaiming fan_out scale
    (as main.py:143-152), affine norm params near (1, 0) but NOT exactly so, biases small non-zero so that bias
    paths are exercised."""
    sd = {}
    for idx, (key, shape) in enumerate(param_shapes(in_channels, out_channels, features)):
        g = torch.Generator().manual_seed(seed * 100003 + idx)
        if key.endswith("num_batches_tracked"):
            sd[key] = torch.tensor(0, dtype=torch.long)
        elif key.endswith("running_mean"):
            sd[key] = torch.zeros(shape, dtype=dtype)
        elif key.endswith("running_var"):
            sd[key] = torch.ones(shape, dtype=dtype)
        elif len(shape) == 5:
            if key.startswith("ups.") and key.count(".") == 2:  # ConvTranspose3d [Cin, Cout, 2,2,2]
                fan = shape[0] * 1.0
                std = math.sqrt(1.0 / fan)
            else:
                fan_out = shape[0] * shape[2] * shape[3] * shape[4]
                std = math.sqrt(2.0 / fan_out)
            sd[key] = (torch.randn(shape, generator=g, dtype=torch.float64) * std).to(dtype)
        elif key.endswith(".weight"):  # norm scale
            sd[key] = (1.0 + 0.1 * torch.randn(shape, generator=g, dtype=torch.float64)).to(dtype)
        else:  # biases (conv and norm)
            sd[key] = (0.1 * torch.randn(shape, generator=g, dtype=torch.float64)).to(dtype)
    return sd


def make_inputs(n, d, h, w, in_channels=4, num_classes=4, seed=0, dtype=torch.float32):
    """SURVEY §8c/8d recipe: z-scored-MRI-like N(0,1) volumes and uniform labels, own generators."""
    g = torch.Generator().manual_seed(1000 + seed)
    x = torch.randn(n, in_channels, d, h, w, generator=g, dtype=torch.float64).to(dtype)
    y = torch.randint(0, num_classes, (n, d, h, w), generator=g)
    return x, y


def make_dropout_masks(n, features, p, seed=0, dtype=torch.float32):
    """Per-(sample, channel) Dropout3d keep masks scaled by 1/(1-p) (main.py:110,174), one per encoder level."""
    masks = []
    for i, f in enumerate(features):
        g = torch.Generator().manual_seed(77 + 13 * seed + i)
        keep = (torch.rand(n, f, generator=g) >= p).to(dtype)
        masks.append(keep / (1.0 - p) if p < 1.0 else keep * 0)
    return masks


# ----------------------------------------------------------------------------------------------
# Model forward
# ----------------------------------------------------------------------------------------------
def double_conv(x, sd, prefix):
    """main.py:235-242: relu(GN8(conv3(relu(GN8(conv3(x)))))) + GN8(conv1(x)); no ReLU after the add."""
    y = F.conv3d(x, sd[prefix + "double_conv.0.weight"], None, padding=1)
    y = F.relu(F.group_norm(y, 8, sd[prefix + "double_conv.1.weight"], sd[prefix + "double_conv.1.bias"], 1e-5))
    y = F.conv3d(y, sd[prefix + "double_conv.3.weight"], None, padding=1)
    y = F.relu(F.group_norm(y, 8, sd[prefix + "double_conv.4.weight"], sd[prefix + "double_conv.4.bias"], 1e-5))
    if (prefix + "residual.0.weight") in sd:
        r = F.conv3d(x, sd[prefix + "residual.0.weight"], None)
        r = F.group_norm(r, 8, sd[prefix + "residual.1.weight"], sd[prefix + "residual.1.bias"], 1e-5)
        y = y + r
    else:  # in == out: identity residual (main.py:225-226)
        y = y + x
    return y


def attention_gate(g, x, sd, prefix):
    """main.py:280-299: x * sigmoid(GN1(conv1(relu(GN4(conv1 g) + GN4(conv1 x))))) * SE(x)."""
    g1 = F.group_norm(F.conv3d(g, sd[prefix + "W_g.0.weight"], sd[prefix + "W_g.0.bias"]), 4,
                      sd[prefix + "W_g.1.weight"], sd[prefix + "W_g.1.bias"], 1e-5)
    x1 = F.group_norm(F.conv3d(x, sd[prefix + "W_x.0.weight"], sd[prefix + "W_x.0.bias"]), 4,
                      sd[prefix + "W_x.1.weight"], sd[prefix + "W_x.1.bias"], 1e-5)
    q = F.relu(g1 + x1)
    p = F.conv3d(q, sd[prefix + "psi.0.weight"], sd[prefix + "psi.0.bias"])
    p = torch.sigmoid(F.group_norm(p, 1, sd[prefix + "psi.1.weight"], sd[prefix + "psi.1.bias"], 1e-5))
    m = x.mean(dim=(2, 3, 4), keepdim=True)
    ca = F.relu(F.conv3d(m, sd[prefix + "channel_attention.1.weight"], sd[prefix + "channel_attention.1.bias"]))
    ca = torch.sigmoid(F.conv3d(ca, sd[prefix + "channel_attention.3.weight"], sd[prefix + "channel_attention.3.bias"]))
    return x * p * ca


def unet_forward(x, sd, features=(32, 64, 128, 256, 512), training=False, dropout_masks=None, bn_momentum=0.1):
    """main.py:154-203.  Returns (main, deep_list, bn_update) — bn_update is the (running_mean, running_var) the
    reference's BatchNorm3d would hold after this call in training mode (None in eval)."""
    features = list(features)
    skips, deep = [], []
    full = x.shape[2:]
    for i, f in enumerate(features):
        x = double_conv(x, sd, "downs.%d." % i)
        skips.append(x)
        if i < len(features) - 1:  # main.py:164-171
            d = F.conv3d(x, sd["deep_supervision.%d.weight" % i], sd["deep_supervision.%d.bias" % i])
            deep.append(F.interpolate(d, size=full, mode="trilinear", align_corners=False))
        x = F.max_pool3d(x, 2, 2)
        if training and dropout_masks is not None:
            x = x * dropout_masks[i][:, :, None, None, None].to(x.dtype)
    x = double_conv(x, sd, "bottleneck.")
    for i, f in enumerate(reversed(features)):
        skip = skips[len(features) - 1 - i]
        x = F.conv_transpose3d(x, sd["ups.%d.weight" % (3 * i)], sd["ups.%d.bias" % (3 * i)], stride=2)
        xa = attention_gate(x, skip, sd, "ups.%d." % (3 * i + 1))
        x = double_conv(torch.cat((xa, x), dim=1), sd, "ups.%d." % (3 * i + 2))
    # final_conv: main.py:129-134
    h = F.conv3d(x, sd["final_conv.0.weight"], sd["final_conv.0.bias"], padding=1)
    bn_update = None
    if training:
        mean = h.mean(dim=(0, 2, 3, 4))
        var_b = h.var(dim=(0, 2, 3, 4), unbiased=False)
        m = h.numel() // h.shape[1]
        bn_update = ((1 - bn_momentum) * sd["final_conv.1.running_mean"] + bn_momentum * mean.detach(),
                     (1 - bn_momentum) * sd["final_conv.1.running_var"] + bn_momentum * var_b.detach() * m / (m - 1))
        hn = (h - mean[None, :, None, None, None]) * torch.rsqrt(var_b + 1e-5)[None, :, None, None, None]
    else:
        hn = (h - sd["final_conv.1.running_mean"][None, :, None, None, None]) * torch.rsqrt(
            sd["final_conv.1.running_var"] + 1e-5)[None, :, None, None, None]
    hn = hn * sd["final_conv.1.weight"][None, :, None, None, None] + sd["final_conv.1.bias"][None, :, None, None, None]
    main = F.conv3d(F.relu(hn), sd["final_conv.3.weight"], sd["final_conv.3.bias"])
    return main, deep, bn_update


# ----------------------------------------------------------------------------------------------
# Losses (losses.py:7-126, training.py:517-566) and metrics (training.py:351-364, main.py:470-474)
# ----------------------------------------------------------------------------------------------
def _one_hot(target, c, dtype):
    return F.one_hot(target, num_classes=c).permute(0, 4, 1, 2, 3).to(dtype)


def dice_loss(pred, target, smooth):
    """losses.py:17-28 / training.py:543-553."""
    p = F.softmax(pred, dim=1)
    o = _one_hot(target, pred.shape[1], pred.dtype)
    inter = (p * o).sum(dim=(2, 3, 4))
    union = p.sum(dim=(2, 3, 4)) + o.sum(dim=(2, 3, 4))
    return 1 - ((2.0 * inter + smooth) / (union + smooth)).mean()


def focal_loss(pred, target, alpha, gamma):
    """losses.py:30-35 / training.py:562-566."""
    ce = F.cross_entropy(pred, target, reduction="none")
    pt = torch.exp(-ce)
    return (alpha * (1 - pt) ** gamma * ce).mean()


def _fwd_grad_mag(t):
    """losses.py:42-52: |forward difference| along D, H, W, zero at the far face, summed."""
    gx = F.pad((t[:, :, 1:] - t[:, :, :-1]).abs(), (0, 0, 0, 0, 0, 1))
    gy = F.pad((t[:, :, :, 1:] - t[:, :, :, :-1]).abs(), (0, 0, 0, 1, 0, 0))
    gz = F.pad((t[:, :, :, :, 1:] - t[:, :, :, :, :-1]).abs(), (0, 1, 0, 0, 0, 0))
    return gx + gy + gz


def boundary_loss(pred, target):
    """losses.py:37-61."""
    p = F.softmax(pred, dim=1)
    o = _one_hot(target, pred.shape[1], pred.dtype)
    return F.mse_loss(_fwd_grad_mag(p), _fwd_grad_mag(o))


def combined_loss3d(pred, target, alpha=0.5, beta=0.3, gamma=0.2, smooth=1e-5):
    """losses.py:63-75 — returns (total, parts dict of tensors)."""
    d = dice_loss(pred, target, smooth)
    f = focal_loss(pred, target, 0.25, 2.0)
    b = boundary_loss(pred, target)
    total = alpha * d + beta * f + gamma * b
    return total, {"dice_loss": d, "focal_loss": f, "boundary_loss": b, "total_loss": total}


def deep_supervision_loss(main, deep, target, weights=(1.0, 0.8, 0.6, 0.4)):
    """losses.py:107-126: main + first len(weights)-1 deep outputs (targets already full-res ⇒ nearest resize = id)."""
    total = combined_loss3d(main, target)[0] * weights[0]
    for i, p in enumerate(deep):
        if i < len(weights) - 1:
            total = total + combined_loss3d(p, target)[0] * weights[i + 1]
    return total


def tversky_loss(pred, target, alpha=0.7, beta=0.3, smooth=1e-5):
    """losses.py:86-97."""
    p = F.softmax(pred, dim=1)
    o = _one_hot(target, pred.shape[1], pred.dtype)
    tp = (p * o).sum(dim=(2, 3, 4))
    fp = (p * (1 - o)).sum(dim=(2, 3, 4))
    fn = ((1 - p) * o).sum(dim=(2, 3, 4))
    return 1 - ((tp + smooth) / (tp + alpha * fp + beta * fn + smooth)).mean()


def trainer_combined_loss(pred, target, weights=(0.5, 0.3, 0.2)):
    """training.py:517-534: 0.5*Dice(1e-6) + 0.3*CE + 0.2*Focal(alpha=1,gamma=2)."""
    return (weights[0] * dice_loss(pred, target, 1e-6) + weights[1] * F.cross_entropy(pred, target)
            + weights[2] * focal_loss(pred, target, 1.0, 2.0))


def confusion_counts(logits, target, num_classes=4):
    """Integer confusion histogram H[pred, true] of the argmax mask — exact integers."""
    pred = torch.argmax(logits, dim=1)
    idx = (pred * num_classes + target).reshape(-1)
    return torch.bincount(idx, minlength=num_classes * num_classes).reshape(num_classes, num_classes)


def dice_score(logits, target):
    """training.py:351-364: mean over classes 1..3 of 2|P∧T| / (|P|+|T|+1e-8), counts cast to fp32 first."""
    pred = torch.argmax(logits, dim=1)
    scores = []
    for c in range(1, 4):
        pm = (pred == c).float()
        tm = (target == c).float()
        inter = (pm * tm).sum()
        scores.append(((2.0 * inter) / (pm.sum() + tm.sum() + 1e-8)).item())
    return float(sum(scores) / len(scores))


def voxel_counts(mask, num_classes=4):
    """main.py:470-474,588-591; utils/visualization.py:217-221,253: total tumour voxels, per-class counts and
    per-axial-slice tumour counts of a [D,H,W] label mask (reference slices the LAST axis: seg[:, :, z])."""
    per_class = [int((mask == c).sum()) for c in range(num_classes)]
    tumour = int((mask > 0).sum())
    per_slice = [(int((mask[:, :, z] > 0).sum())) for z in range(mask.shape[2])]
    return tumour, per_class, per_slice


# ---------------------------------------------------------------------------------------------------------------------
# BrainTumorClassifier (main.py:301-328; call site classify_tumor main.py:398-425): eval-mode forward only — the reference
# never trains it.  Conv3d(4,32,3,1,1) ReLU MaxPool3d(2) Conv3d(32,64) ReLU MaxPool3d(2) Conv3d(64,128) ReLU
# AdaptiveAvgPool3d(4) flatten Linear(8192,512) ReLU Dropout(0.5) Linear(512,num_classes)
# ---------------------------------------------------------------------------------------------------------------------
def classifier_param_shapes(num_classes=4):
    return [("features.0.weight", (32, 4, 3, 3, 3)), ("features.0.bias", (32,)),
            ("features.3.weight", (64, 32, 3, 3, 3)), ("features.3.bias", (64,)),
            ("features.6.weight", (128, 64, 3, 3, 3)), ("features.6.bias", (128,)),
            ("classifier.0.weight", (512, 128 * 4 * 4 * 4)), ("classifier.0.bias", (512,)),
            ("classifier.3.weight", (num_classes, 512)), ("classifier.3.bias", (num_classes,))]


def make_classifier_state_dict(num_classes=4, seed=0, dtype=torch.float32):
    """Deterministic, reference-independent weights (own generator per tensor; Kaiming-like scales, small biases)."""
    sd = {}
    for idx, (key, shape) in enumerate(classifier_param_shapes(num_classes)):
        g = torch.Generator().manual_seed(seed * 100003 + 5000 + idx)
        if len(shape) == 5:
            std = math.sqrt(2.0 / (shape[1] * 27))
        elif len(shape) == 2:
            std = math.sqrt(2.0 / shape[1])
        else:
            std = 0.1
        sd[key] = (torch.randn(shape, generator=g, dtype=torch.float64) * std).to(dtype)
    return sd


def classifier_forward(x, sd):
    """main.py:324-328 in eval mode (Dropout is the identity).  x: [N,4,D,H,W] -> logits [N,num_classes]."""
    x = F.max_pool3d(F.relu(F.conv3d(x, sd["features.0.weight"], sd["features.0.bias"], padding=1)), 2)
    x = F.max_pool3d(F.relu(F.conv3d(x, sd["features.3.weight"], sd["features.3.bias"], padding=1)), 2)
    x = F.relu(F.conv3d(x, sd["features.6.weight"], sd["features.6.bias"], padding=1))
    x = F.adaptive_avg_pool3d(x, (4, 4, 4))
    x = x.reshape(x.shape[0], -1)
    x = F.relu(F.linear(x, sd["classifier.0.weight"], sd["classifier.0.bias"]))
    return F.linear(x, sd["classifier.3.weight"], sd["classifier.3.bias"])
