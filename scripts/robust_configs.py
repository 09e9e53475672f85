"""BASELINE configs 4 and 5 smoke + timing: big per-GPU batch, and the wide model at 160x192x160."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import b3d  # noqa
import unet3d_b200 as U
dev = torch.device("cuda:0")
def run(name, feats, n, d, h, w, iters=3):
    torch.manual_seed(0)
    model = U.UNet3D(4, 4, features=feats).to(dev).train()
    crit = U.DeepSupervisionLoss3D()
    opt = torch.optim.AdamW(model.parameters(), lr=1e-4, fused=True)
    x = torch.randn(n, 4, d, h, w, device=dev)
    y = torch.randint(0, 4, (n, d, h, w), device=dev)
    def step():
        opt.zero_grad(set_to_none=True)
        loss = crit(model(x), y); loss.backward(); opt.step(); return loss
    l0 = float(step()); step(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): l = step()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    l = float(l)
    ok = l == l and abs(l) < 1e3 and l < l0 + 0.5
    print("%s: N%d %dx%dx%d feats %s: %.2f ms/step, %.3g voxels/s, loss %.4f -> %.4f, mem %.1f GB %s" % (
        name, n, d, h, w, feats[0], ms, n * d * h * w / ms * 1e3, l0, l, torch.cuda.max_memory_allocated() / 2**30, "OK" if ok else "BAD"), flush=True)
    del model, opt, x, y
    torch.cuda.empty_cache()
run("cfg4 batch 8", [32, 64, 128, 256, 512], 8, 128, 128, 128)
run("cfg5 wide", [64, 128, 256, 512, 1024], 1, 160, 192, 160)
run("odd dims", [32, 64, 128, 256, 512], 1, 96, 160, 64)
