"""CUDA-graph capture of the whole training step (at ~20 ms per step the ~750 kernel launches and the TMA
descriptor encodes of one step cost more host time than the GPU needs for the small layers).

`GraphedTrainStep` captures  zero_grad -> forward -> criterion -> backward (-> gradient all-reduce) -> optimizer.step  once
on static input buffers and replays it; the trainer call sites (training.py:290-304) keep their shape:

    step = GraphedTrainStep(model, criterion, optimizer, images, masks)
    for images, masks in loader:
        loss = step(images, masks)          # device tensor; read it with .item() only when you need the number

Everything the step launches (libb3d kernels, memsets, torch's fused AdamW, NCCL all-reduces) is stream-ordered and free of
host synchronisation, which is what makes it capturable.  The optimizer must be constructed with capturable=True, and its
learning rate must live in a DEVICE tensor: a python-float lr is baked into the captured kernels, so a scheduler
(`CosineAnnealingWarmRestarts`, training.py:195,252) would silently become a no-op behind the replayed graph.  A float lr
is therefore replaced by a CUDA scalar tensor before the capture (torch's schedulers `fill_` a tensor lr in place).
Loss scaling: the path computes in bf16 (fp32 exponent range), so `GradScaler` is unnecessary; the reference's literal
`autocast` + `scaler.scale(loss).backward()` + `scaler.step()` sequence (training.py:292-299) works on the EAGER path
(tests/test_gpu_trainer_loop.py) but is not captured here (`scaler.step` reads the inf flag on the host).
"""
import torch

from . import functional, ops


def _capture_stream():
    """The capture stream of the main chain.  B3D_MAIN_PRIORITY=-1 makes its kernel nodes outrank the weight-gradient side
    stream (ops.WGRAD_STREAM, priority 0) when both have blocks pending (tuning knob; default 0 = equal priority)."""
    import os
    return torch.cuda.Stream(priority=int(os.environ.get("B3D_MAIN_PRIORITY", "0")))


class GraphedTrainStep:
    def __init__(self, model, criterion, optimizer, example_x, example_y, warmup=3):
        self.model, self.criterion, self.optimizer = model, criterion, optimizer
        for group in optimizer.param_groups:   # see the module doc: the lr must be a device scalar or schedulers stop working
            if not torch.is_tensor(group["lr"]):
                group["lr"] = torch.tensor(float(group["lr"]), dtype=torch.float32, device=example_x.device)
            elif not group["lr"].is_cuda:
                group["lr"] = group["lr"].to(example_x.device)
        self.x = example_x.detach().clone()
        self.y = example_y.detach().clone()
        # input pipelining (prefetch / step_prefetched): staging buffers, copy stream and events exist before the first timed
        # step — allocating them lazily cost a one-off 100+ ms (allocator growth next to the graph's private pool)
        self._copy_stream = torch.cuda.Stream()
        self._sx, self._sy = torch.empty_like(self.x), torch.empty_like(self.y)
        self._staged = torch.cuda.Event()
        self._consumed = torch.cuda.Event()
        self._consumed.record()
        self.graph = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):  # eager warm-up on a side stream (torch's capture protocol); also fills every cache
            for _ in range(warmup):
                self._eager()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.optimizer.zero_grad(set_to_none=True)
        # thread_local: CUDA calls of other host threads (e.g. the NCCL watchdog polling events) must not invalidate the capture
        ops.reset_scratch()   # the capture must zero-fill its own scratch arenas (and nothing eager may slice them later)
        with torch.cuda.graph(self.graph, stream=_capture_stream(), capture_error_mode="thread_local"):
            self.loss = self._eager()
        ops.reset_scratch()
        torch.cuda.synchronize()

    def _eager(self):
        self.optimizer.zero_grad(set_to_none=True)
        loss = self.criterion(self.model(self.x), self.y)
        if isinstance(loss, tuple):
            loss = loss[0]
        loss.backward()
        self.optimizer.step()
        return loss.detach()

    def __call__(self, x, y):
        """Copies the batch into the static buffers (device->device or pinned host->device, stream ordered) and replays."""
        if x is not self.x:
            self.x.copy_(x, non_blocking=True)
        if y is not self.y:
            self.y.copy_(y, non_blocking=True)
        self.graph.replay()
        functional.clear_pack_cache()   # parameters changed behind the version counters: eager calls must re-pack
        return self.loss

    def close(self):
        """Release the captured graph (and the NCCL work it holds) — call before `destroy_process_group()`."""
        torch.cuda.synchronize()
        self.graph.reset()
        self.graph = None
        functional.clear_pack_cache()

    # -- input pipelining: the H2D copy of batch i+1 runs on a copy stream while the graph of batch i executes ----------
    def prefetch(self, x_host, y_host):
        """Start copying the NEXT batch (pinned host tensors) into staging buffers on a side stream."""
        self._copy_stream.wait_event(self._consumed)       # the previous staged batch has been moved into the static buffers
        with torch.cuda.stream(self._copy_stream):
            self._sx.copy_(x_host, non_blocking=True)
            self._sy.copy_(y_host, non_blocking=True)
            self._staged.record()

    def step_prefetched(self):
        """Run one step on the batch handed to the last prefetch() (device-to-device move, then replay)."""
        cur = torch.cuda.current_stream()
        cur.wait_event(self._staged)
        self.x.copy_(self._sx, non_blocking=True)
        self.y.copy_(self._sy, non_blocking=True)
        self._consumed.record()
        self.graph.replay()
        functional.clear_pack_cache()
        return self.loss


class GraphedInference:
    """Eval-mode forward (main.py:390-393 / train_model.py:216-217 call sites) captured once and replayed:
        infer = GraphedInference(model, volume);  logits = infer(volume)   # static output buffer, fp32 NCDHW
    The bf16 packed weights are frozen at capture time (serving); call `refresh()` after the parameters changed.  The graph
    OWNS those packed copies: the captured kernels hold raw pointers into them, so the instance keeps a reference to every
    packed tensor it captured — a later cache miss of `functional.packed` (optimizer step, `clear_pack_cache()`, a second
    GraphedInference on the same model) replaces the per-parameter cache entries but can no longer free these buffers."""

    def __init__(self, model, example_x, warmup=3):
        assert not model.training, "GraphedInference captures the eval-mode forward"
        self.model = model
        self.x = example_x.detach().clone()
        self._warmup = warmup
        self.refresh()

    def refresh(self):
        warmup = self._warmup
        functional.clear_pack_cache()
        self.graph = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(warmup):
                self.model(self.x)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        ops.reset_scratch()
        with torch.no_grad(), torch.cuda.graph(self.graph, capture_error_mode="thread_local"):
            self.out = self.model(self.x)
        ops.reset_scratch()
        torch.cuda.synchronize()
        self._packs = [entry for prm in self.model.parameters() for entry in prm.__dict__.get("_b3d_pack", {}).values()]

    def close(self):
        torch.cuda.synchronize()
        self.graph.reset()
        self.graph, self._packs = None, None

    def __call__(self, x):
        if x is not self.x:
            self.x.copy_(x, non_blocking=True)
        self.graph.replay()
        return self.out
