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
    """Bucketed all-reduce of parameter gradients in emission order.

    First backward pass: the (name, shape) sequence in which the model emits its gradients is recorded and that pass is reduced
    the simple way (gather -> all-reduce -> scatter back).  From then on the layout is PERSISTENT: flat fp32 buckets are
    allocated once, every gradient has a fixed slice, the weight-gradient kernels write straight into their slices
    (`sink`), the remaining small gradients are moved with one multi-tensor copy per block, each bucket is all-reduced IN
    PLACE (ncclAvg) the moment its last gradient exists, and autograd receives views of the buckets — no `torch.cat`, no
    per-tensor copy back, and `.grad` addresses that never change (CUDA graphs, FusedAdamW tables)."""

    def __init__(self, process_group=None, bucket_mb=32.0, average=True, persistent=True):
        self.group = process_group
        self.bucket_bytes = int(bucket_mb * (1 << 20))
        self.average = average
        self.persistent = persistent
        self.pending, self.pending_bytes = [], 0
        self.inflight = []
        self.comm_stream = None
        self.buckets_launched = 0
        self._order = []          # first pass: [(name, shape, dtype, device)]
        self._layout = None       # name -> (bucket index, offset, numel, shape)
        self._flat = []           # bucket tensors
        self._remaining = []      # per bucket: gradients still missing in this pass
        self._count = []          # per bucket: gradients per pass
        self._views_valid = False

    def world(self):
        return dist.get_world_size(self.group) if dist.is_initialized() else 1

    # -- persistent layout ---------------------------------------------------------------------------------------------
    def _build_layout(self):
        layout, flat, count = {}, [], []
        cur, cur_bytes, dev = [], 0, None
        def close():
            nonlocal cur, cur_bytes
            if not cur:
                return
            total = sum(n for _, n, _ in cur)
            buf = torch.zeros(total, dtype=torch.float32, device=dev)
            off = 0
            for name, n, shape in cur:
                layout[name] = (len(flat), off, n, shape)
                off += n
            flat.append(buf)
            count.append(len(cur))
            cur, cur_bytes = [], 0
        for name, shape, dtype, device in self._order:
            if name in layout or any(name == c[0] for c in cur):
                continue
            n = 1
            for d_ in shape:
                n *= int(d_)
            dev = device
            cur.append((name, n, tuple(shape)))
            cur_bytes += 4 * n
            if cur_bytes >= self.bucket_bytes:
                close()
        close()
        self._layout, self._flat, self._count = layout, flat, count
        self._remaining = list(count)

    def sink(self, name, shape):
        """The slice of the persistent bucket a kernel may write gradient `name` into (None until the layout exists)."""
        if self._layout is None or self.world() == 1:
            return None
        ent = self._layout.get(name)
        if ent is None or tuple(ent[3]) != tuple(shape):
            return None
        b, off, n, shp = ent
        return self._flat[b][off:off + n].view(shp)

    def view(self, name):
        b, off, n, shp = self._layout[name]
        return self._flat[b][off:off + n].view(shp)

    # -- per pass ------------------------------------------------------------------------------------------------------
    def add(self, grads):
        """grads: dict name -> tensor (a block's freshly computed parameter gradients)."""
        if self.world() == 1:
            return
        if self._layout is not None:
            return self._add_persistent(grads)
        for name, g in grads.items():
            if g is None:
                continue
            if self.persistent:
                self._order.append((name, tuple(g.shape), g.dtype, g.device))
            self.pending.append(g)
            self.pending_bytes += g.numel() * g.element_size()
        if self.pending_bytes >= self.bucket_bytes:
            self._launch()

    def _comm_ctx(self, cuda):
        if not cuda:
            return _Null()
        if self.comm_stream is None:
            self.comm_stream = torch.cuda.Stream()
        ev = torch.cuda.Event()
        ev.record()  # the gradients were produced on the current (compute) stream
        self.comm_stream.wait_event(ev)
        from . import ops as _ops
        if _ops.WGRAD_STREAM is not None:   # ... or on the weight-gradient side stream
            self.comm_stream.wait_stream(_ops.WGRAD_STREAM)
        return torch.cuda.stream(self.comm_stream)

    def _allreduce(self, flat):
        if self.average and flat.is_cuda:   # NCCL averages inside the collective: no separate pass over the bucket
            return dist.all_reduce(flat, op=dist.ReduceOp.AVG, group=self.group, async_op=True)
        if self.average:
            flat.div_(self.world())
        return dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True)

    def _add_persistent(self, grads):
        dst, src, touched = [], [], set()
        for name, g in list(grads.items()):
            if g is None:
                continue
            ent = self._layout.get(name)
            if ent is None:
                raise RuntimeError("GradientBuckets: gradient %r was not part of the first backward pass" % name)
            v = self.view(name)
            if g.data_ptr() != v.data_ptr():
                dst.append(v)
                src.append(g.reshape(v.shape) if g.dtype == torch.float32 else g.reshape(v.shape).float())
            grads[name] = v
            touched.add(ent[0])
            self._remaining[ent[0]] -= 1
        if dst:
            torch._foreach_copy_(dst, src)
        for b in sorted(touched):
            if self._remaining[b] == 0:
                flat = self._flat[b]
                with self._comm_ctx(flat.is_cuda):
                    work = self._allreduce(flat)
                    self.inflight.append((work, flat, None))
                self.buckets_launched += 1
                self._remaining[b] = -1   # launched

    def _launch(self):
        if not self.pending:
            return
        tensors, self.pending, self.pending_bytes = self.pending, [], 0
        cuda = tensors[0].is_cuda
        with self._comm_ctx(cuda):
            flat = torch.cat([t.reshape(-1).float() for t in tensors])
            work = self._allreduce(flat)
            self.inflight.append((work, flat, tensors))
            if cuda:
                for t in tensors:
                    t.record_stream(self.comm_stream)
        self.buckets_launched += 1

    def finish(self):
        """Flush what is left, wait for every all-reduce; first pass: scatter the averaged values back and fix the layout."""
        if self.world() == 1:
            return
        self._views_valid = self._layout is not None   # this pass's gradients live in the buckets (not so in the first pass)
        if self._layout is not None:
            for b, rem in enumerate(self._remaining):   # a bucket whose gradients did not all arrive this pass
                if rem >= 0 and rem < self._count[b]:
                    flat = self._flat[b]
                    with self._comm_ctx(flat.is_cuda):
                        self.inflight.append((self._allreduce(flat), flat, None))
                    self.buckets_launched += 1
            self._remaining = list(self._count)
        else:
            self._launch()
        for work, flat, tensors in self.inflight:
            cuda = flat.is_cuda
            ctx = torch.cuda.stream(self.comm_stream) if cuda else _Null()
            with ctx:
                work.wait()
                if tensors is not None:
                    off = 0
                    for t in tensors:
                        n = t.numel()
                        t.copy_(flat[off:off + n].reshape(t.shape))
                        off += n
        if self.comm_stream is not None:
            torch.cuda.current_stream().wait_stream(self.comm_stream)
        self.inflight = []
        if self._layout is None and self.persistent and self._order:
            self._build_layout()

    def fresh_views(self, grads):
        """Replace bucket-resident gradients by NEW view objects (autograd adopts a gradient it holds the only reference to
        without copying; `.grad` then aliases the bucket)."""
        if self._layout is None or self.world() == 1 or not self._views_valid:
            return grads
        return {k: (self.view(k) if (v is not None and k in self._layout) else v) for k, v in grads.items()}


class _Null:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


class DataParallel(torch.nn.Module):
    """Wrap a UNet3D replica: forward is unchanged; backward all-reduces gradients bucket by bucket (see module doc)."""

    def __init__(self, module, process_group=None, bucket_mb=32.0, broadcast_parameters=True, reserved_sms=0):
        """reserved_sms > 0: size every grid of the library for (SM count - reserved_sms) SMs (`b3d_set_reserved_sms`, process-wide)
        so NCCL's channel CTAs get SMs of their own; pair it with NCCL_MAX_CTAS=reserved_sms set before the communicator is
        created.  Measured on 2 x B200 (profiles/dp_reserved_sms_r2.txt): no gain, so the default leaves it off."""
        super().__init__()
        self.module = module
        if reserved_sms > 0:
            from . import _lib
            _lib.set_reserved_sms(reserved_sms)
        self.buckets = GradientBuckets(process_group, bucket_mb)
        module._on_grads = self.buckets.add
        module._on_backward_end = self.buckets.finish
        module._grad_sink = self.buckets.sink
        module._grad_views = self.buckets.fresh_views
        if broadcast_parameters and dist.is_initialized() and dist.get_world_size(process_group) > 1:
            for t in list(module.parameters()) + list(module.buffers()):
                dist.broadcast(t.data, src=0, group=process_group)
            from . import functional   # `.data` writes bump no version counter: drop bf16 packs made before the broadcast
            functional.clear_pack_cache()

    def forward(self, *args, **kwargs):
        return self.module(*args, **kwargs)
