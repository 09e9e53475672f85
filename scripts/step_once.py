"""One training step (+1 warm-up) of cfg 3 for the ncu launch-list pass."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import b3d  # noqa
import unet3d_b200 as U
dev = torch.device("cuda:0")
torch.manual_seed(0)
model = U.UNet3D(4, 4).to(dev).train()
crit = U.DeepSupervisionLoss3D()
x = torch.randn(2, 4, 128, 128, 128, device=dev)
y = torch.randint(0, 4, (2, 128, 128, 128), device=dev)
for it in range(int(sys.argv[1]) if len(sys.argv) > 1 else 2):
    model.zero_grad(set_to_none=True)
    loss = crit(model(x), y)
    loss.backward()
torch.cuda.synchronize()
print("loss", float(loss))
