"""cfg 2 (1 x 4 x 128^3 eval forward) a few times, eager, for the ncu launch-list pass; prints the CUDA-event latency."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import b3d  # noqa
import unet3d_b200 as U
dev = torch.device("cuda:0")
torch.manual_seed(0)
model = U.UNet3D(4, 4).to(dev).eval()
x = torch.randn(1, 4, 128, 128, 128, device=dev)
with torch.no_grad():
    for _ in range(3):
        model(x)
    torch.cuda.synchronize()
    g = U.GraphedInference(model, x)
    for _ in range(3):
        g(x)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(10):
        g(x)
    e1.record()
    torch.cuda.synchronize()
    print("graphed inference: %.3f ms per volume" % (e0.elapsed_time(e1) / 10))
    torch.cuda.synchronize()
    model(x)   # one eager forward at the end: the launch list's last `to_ndhwc` segment
    torch.cuda.synchronize()
