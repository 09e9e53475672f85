"""Data-parallel training: one process per GPU, bucketed gradient all-reduce overlapped with backward.

The reference has no distributed code; the hot path shards naturally by batch: every rank owns a
full replica and `global_batch / world` volumes, and the only exchange is the all-reduce (average) of the parameter
gradients.  UNet3D's hand-scheduled backward hands every block's gradients to `GradientBuckets.add` the moment they
exist (reverse-topological order: final_conv, ups.14 ... ups.0, bottleneck, downs.4 ... downs.0), so each ~bucket_mb
bucket is all-reduced by NCCL over NVLink on a side stream while the remaining dgrad/wgrad kernels run; the
optimizer-facing gradients are only touched after `finish()` joined the streams.  BatchNorm statistics of final_conv stay
per replica (the reference has no SyncBN).
"""
import torch
import torch.distributed as dist


class GradientBuckets:
    def __init__(self, process_group=None, bucket_mb=32.0, average=True):
        self.group = process_group
        self.bucket_bytes = int(bucket_mb * (1 << 20))
        self.average = average
        self.pending, self.pending_bytes = [], 0
        self.inflight = []
        self.comm_stream = None
        self.buckets_launched = 0

    def world(self):
        return dist.get_world_size(self.group) if dist.is_initialized() else 1

    def add(self, grads):
        """grads: dict name -> tensor (a block's freshly computed parameter gradients)."""
        if self.world() == 1:
            return
        for g in grads.values():
            if g is None:
                continue
            self.pending.append(g)
            self.pending_bytes += g.numel() * g.element_size()
        if self.pending_bytes >= self.bucket_bytes:
            self._launch()

    def _launch(self):
        if not self.pending:
            return
        tensors, self.pending, self.pending_bytes = self.pending, [], 0
        cuda = tensors[0].is_cuda
        if cuda:
            if self.comm_stream is None:
                self.comm_stream = torch.cuda.Stream()
            ev = torch.cuda.Event()
            ev.record()  # the gradients were produced on the current (compute) stream
            self.comm_stream.wait_event(ev)
            from . import ops as _ops
            if _ops.WGRAD_STREAM is not None:   # ... or on the weight-gradient side stream
                self.comm_stream.wait_stream(_ops.WGRAD_STREAM)
            ctx = torch.cuda.stream(self.comm_stream)
        else:
            ctx = _Null()
        with ctx:
            flat = torch.cat([t.reshape(-1).float() for t in tensors])
            if self.average and cuda:   # NCCL averages inside the collective: no separate pass over the bucket
                work = dist.all_reduce(flat, op=dist.ReduceOp.AVG, group=self.group, async_op=True)
            else:
                if self.average:
                    flat.div_(self.world())
                work = dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
            self.inflight.append((work, flat, tensors))
            if cuda:
                for t in tensors:
                    t.record_stream(self.comm_stream)
        self.buckets_launched += 1

    def finish(self):
        """Flush the last partial bucket, wait for every all-reduce and scatter the averaged values back in place."""
        if self.world() == 1:
            return
        self._launch()
        for work, flat, tensors in self.inflight:
            cuda = flat.is_cuda
            ctx = torch.cuda.stream(self.comm_stream) if cuda else _Null()
            with ctx:
                work.wait()
                off = 0
                for t in tensors:
                    n = t.numel()
                    t.copy_(flat[off:off + n].reshape(t.shape))
                    off += n
        if self.comm_stream is not None:
            torch.cuda.current_stream().wait_stream(self.comm_stream)
        self.inflight = []


class _Null:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


class DataParallel(torch.nn.Module):
    """Wrap a UNet3D replica: forward is unchanged; backward all-reduces gradients bucket by bucket (see module doc)."""

    def __init__(self, module, process_group=None, bucket_mb=32.0, broadcast_parameters=True):
        super().__init__()
        self.module = module
        self.buckets = GradientBuckets(process_group, bucket_mb)
        module._on_grads = self.buckets.add
        module._on_backward_end = self.buckets.finish
        if broadcast_parameters and dist.is_initialized() and dist.get_world_size(process_group) > 1:
            for t in list(module.parameters()) + list(module.buffers()):
                dist.broadcast(t.data, src=0, group=process_group)
            from . import functional   # `.data` writes bump no version counter: drop bf16 packs made before the broadcast
            functional.clear_pack_cache()

    def forward(self, *args, **kwargs):
        return self.module(*args, **kwargs)
