"""Where does the end-to-end step (pinned host batch -> H2D -> graph replay -> loss .item()) spend its time?  Replays bench.py's
e2e loop with CUDA events on the copy stream / compute stream and host timestamps around every call."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import b3d  # noqa
import unet3d_b200 as U
dev = torch.device("cuda:0")
torch.manual_seed(0)
model = U.UNet3D(4, 4, dropout_rate=0.2).to(dev).train()
crit = U.DeepSupervisionLoss3D()
opt = U.make_adamw(model, lr=1e-4, weight_decay=1e-4)
g = torch.Generator().manual_seed(1000)
x_host = torch.randn(2, 4, 128, 128, 128, generator=g).pin_memory()
y_host = torch.randint(0, 4, (2, 128, 128, 128), generator=g).pin_memory()
print("pinned:", x_host.is_pinned(), y_host.is_pinned(), flush=True)
xd, yd = x_host.to(dev), y_host.to(dev)
step = U.GraphedTrainStep(model, crit, opt, xd, yd, warmup=3)
for _ in range(3):
    step(xd, yd)
torch.cuda.synchronize()
ev = lambda: torch.cuda.Event(enable_timing=True)
for mode in ("resident", "e2e", "e2e_nosync_item_every_step", "e2e"):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    rows = []
    if mode.startswith("e2e"):
        step.prefetch(x_host, y_host)
    for it in range(8):
        c0 = time.perf_counter()
        a, b = ev(), ev()
        a.record()
        if mode == "resident":
            lt = step(xd, yd)
        else:
            lt = step.step_prefetched()
        b.record()
        c1 = time.perf_counter()
        if mode.startswith("e2e") and it < 7:
            h0, h1 = ev(), ev()
            with torch.cuda.stream(step._copy_stream):
                pass
            step._copy_stream.wait_event(step._consumed)
            h0.record(step._copy_stream)
            step.prefetch(x_host, y_host)
            h1.record(step._copy_stream)
        else:
            h0 = h1 = None
        c2 = time.perf_counter()
        if mode != "e2e_nosync_item_every_step":
            lt.item()
        c3 = time.perf_counter()
        rows.append((a, b, h0, h1, c1 - c0, c2 - c1, c3 - c2))
    torch.cuda.synchronize()
    tot = (time.perf_counter() - t0) * 1e3 / 8
    print("== %s: %.2f ms/step (host clock)" % (mode, tot))
    for i, (a, b, h0, h1, d1, d2, d3) in enumerate(rows):
        print("  it %d: step call->end on stream %.2f ms | H2D on copy stream %s ms | host: step() %.2f prefetch() %.2f item() %.2f ms" % (
            i, a.elapsed_time(b), ("%.2f" % h0.elapsed_time(h1)) if h0 is not None else "-", d1 * 1e3, d2 * 1e3, d3 * 1e3))
