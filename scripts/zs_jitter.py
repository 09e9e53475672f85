"""Quantify the run-to-run jitter of the z-marching 3x3x3 conv kernel (two ping-pong MMA issuers): repeat one conv on the same
input, count differing output elements, and bound every run's error against an fp32 cuDNN reference (TF32 off)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
import b3d  # noqa
from unet3d_b200 import ops, _lib
torch.backends.cudnn.allow_tf32 = False
dev = torch.device("cuda:0")
for (n, s, cin, cout) in ((1, 128, 16, 32), (2, 128, 32, 32), (2, 64, 64, 64)):
    g = torch.Generator().manual_seed(1)
    x = torch.randn(n, s, s, s, cin, generator=g).to(dev).bfloat16()
    w = (torch.randn(cout, cin, 3, 3, 3, generator=g) / (cin * 27) ** 0.5).to(dev).bfloat16().float()
    wp, kp, rows = ops.pack_weight(w, ops.PACK_FPROP)
    ref = F.conv3d(x.float().permute(0, 4, 1, 2, 3), w, padding=1).permute(0, 2, 3, 4, 1)
    scale = float(ref.abs().max())
    for mode in (0, 1, 2):
        _lib.set_ordered_issue(mode)
        outs = []
        for _ in range(6):
            y, st = ops.conv_fprop(x, wp, rows, cout, 3, groups=8)
            outs.append((y.clone(), st.clone()))
        torch.cuda.synchronize()
        ndiff = [int((o[0] != outs[0][0]).sum()) for o in outs[1:]]
        maxd = [float((o[0].float() - outs[0][0].float()).abs().max()) for o in outs[1:]]
        err = [float((o[0].float() - ref).abs().max()) / scale for o in outs]
        sd = [float(((o[1] - outs[0][1]).abs() / (outs[0][1].abs() + 1)).max()) for o in outs[1:]]
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            ops.conv_fprop(x, wp, rows, cout, 3, groups=8)
        e1.record(); torch.cuda.synchronize()
        print("   %.3f ms per launch" % (e0.elapsed_time(e1) / 10))
        print("conv %dx%d^3 %d->%d ordered=%d: differing elements vs run 0: %s of %d, max |d| %s, max err vs fp32 / scale %s, "
              "stats rel diff %s" % (n, s, cin, cout, mode, ndiff, outs[0][0].numel(), ["%.2e" % v for v in maxd],
                                     ["%.2e" % v for v in err], ["%.1e" % v for v in sd]), flush=True)
_lib.set_ordered_issue(False)
