"""A few launches of each kernel whose `ncu --set full` capture is committed under profiles/ (one target per invocation):
    python scripts/ncu_targets.py zs | pwadd | dsloss | adamw | gnbwd | deep | convT | convTd | wgdeep"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import b3d  # noqa
import unet3d_b200 as U
from unet3d_b200 import ops
dev = torch.device("cuda:0")
torch.manual_seed(0)
what = sys.argv[1]


def b3d_plan():
    import ctypes
    from unet3d_b200 import _lib
    f = _lib.lib().b3d_last_plan
    f.restype = ctypes.c_char_p
    return f().decode()


bf = torch.bfloat16
if what == "zs":        # 32 -> 32 3x3x3 conv @2x128^3 with GroupNorm statistics: zs_kernel<32,32>
    x = torch.randn(2, 128, 128, 128, 32, device=dev).to(bf)
    w = torch.randn(32, 32, 3, 3, 3, device=dev) * 0.05
    wp, kp, rows = ops.pack_weight(w, ops.PACK_FPROP)
    for _ in range(4):
        ops.conv_fprop(x, wp, rows, 32, 3, groups=8)
elif what == "pwadd":   # 1x1 dgrad 16 -> 32 with the addend on the tensor core (attention-gate W_x dgrad): igemm_kernel
    x = torch.randn(2, 128, 128, 128, 16, device=dev).to(bf)
    w = torch.randn(16, 32, 1, 1, 1, device=dev) * 0.1           # layer weight [Cout=16][Cin=32]; dgrad maps 16 -> 32
    wp, kp, rows = ops.pack_weight(w, ops.PACK_DGRAD)
    dx = torch.randn(2, 128, 128, 128, 32, device=dev).to(bf)
    for _ in range(4):
        ops.conv_fprop(x, wp, rows, 32, 1, out=dx, add=dx)
elif what == "dsloss":  # fused deep-supervision loss, scale 2, forward + backward
    lo = torch.randn(2, 64, 64, 64, 4, device=dev)
    y = torch.randint(0, 4, (2, 128, 128, 128), device=dev)
    t8 = ops.target_u8(y)
    cfg = ops.loss_cfg(w_dice=0.5, smooth=1e-5, w_focal=0.3, f_alpha=0.25, f_gamma=2.0, w_boundary=0.2)
    g = torch.ones(1, device=dev)
    for _ in range(4):
        v, acc = ops.dsloss_fwd(lo, t8, cfg, (128, 128, 128))
        ops.dsloss_bwd(lo, t8, acc, cfg, g, 0.8, (128, 128, 128))
elif what == "adamw":   # fused AdamW + re-pack over the whole default model
    model = U.UNet3D(4, 4).to(dev)
    opt = U.make_adamw(model)
    for p in model.parameters():
        p.grad = torch.randn_like(p) * 0.01
    for _ in range(4):
        opt.step()
elif what == "gnbwd":   # dual GroupNorm backward @2x128^3x32
    y2 = torch.randn(2, 128, 128, 128, 32, device=dev).to(bf)
    r = torch.randn(2, 128, 128, 128, 32, device=dev).to(bf)
    dy = torch.randn(2, 128, 128, 128, 32, device=dev).to(bf)
    gam, bet = torch.ones(32, device=dev), torch.zeros(32, device=dev)
    def stats(t):
        f = t.float().reshape(2, -1, 8, 4)
        return torch.stack([f.sum(dim=(1, 3)), (f * f).sum(dim=(1, 3))], dim=-1).double()
    sa, sb = stats(y2), stats(r)
    for _ in range(4):
        ops.gn_bwd_dual(dy, y2, sa, gam, bet, r, sb, gam, 8)
elif what == "deep":    # weight-streaming deep-level conv: 3x3x3 1024 -> 512 @2x8^3 (ups.2 conv1), split-K igemm_kernel
    x = torch.randn(2, 8, 8, 8, 1024, device=dev).to(bf)
    w = torch.randn(512, 1024, 3, 3, 3, device=dev) * 0.01
    wp, kp, rows = ops.pack_weight(w, ops.PACK_FPROP)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for _ in range(4):
        flush.zero_()                      # weights come from HBM, as inside a real step
        ops.conv_fprop(x, wp, rows, 512, 3, groups=8)
elif what == "convT":   # ConvTranspose3d(k2,s2) fprop 64 -> 32 @2x64^3 -> 2x128^3 (ups.12), pixel-shuffle epilogue into a concat slice
    x = torch.randn(2, 64, 64, 64, 64, device=dev).to(bf)
    w = torch.randn(64, 32, 2, 2, 2, device=dev) * 0.05
    packs = ops.pack_weight_pair(w, True)
    wp = packs[ops.PACK_CONVT_FPROP][0]
    cat = torch.empty(2, 128, 128, 128, 64, device=dev, dtype=bf)
    bias = torch.zeros(32, device=dev)
    for _ in range(4):
        ops.convT2_fprop(x, wp, bias, 32, out=cat[..., 32:])
    print(b3d_plan())
elif what == "convTd":  # its data gradient: 2x128^3x32 -> 2x64^3x64 (8 strided K-maps)
    dy = torch.randn(2, 128, 128, 128, 32, device=dev).to(bf)
    w = torch.randn(64, 32, 2, 2, 2, device=dev) * 0.05
    packs = ops.pack_weight_pair(w, True)
    wd, _, rows = packs[ops.PACK_CONVT_DGRAD]
    for _ in range(4):
        ops.convT2_dgrad(dy, wd, rows, 64)
    print(b3d_plan())
elif what == "wgdeep":  # bottleneck weight gradient: 3x3x3 1024 -> 1024 @2x4^3 (113 MB of fp32 dW): wgrad_kernel + finalize
    x = torch.randn(2, 4, 4, 4, 1024, device=dev).to(bf)
    dy = torch.randn(2, 4, 4, 4, 1024, device=dev).to(bf)
    ops.WGRAD_SIDE = False
    for _ in range(4):
        ops.conv_wgrad(x, dy, 1024, 1024, 3)
torch.cuda.synchronize()
print("ok", what)
