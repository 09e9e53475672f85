"""Host-side logic of the data-parallel path on CPU: world_size-2 gloo, bucketed all-reduce of gradient dicts."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    import b3d  # noqa: F401
    from unet3d_b200.parallel import GradientBuckets
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(rank)
    b = GradientBuckets(bucket_mb=0.001)  # ~1 KB buckets -> several launches
    blocks = []
    for i in range(5):
        g = {"w%d" % i: torch.randn(16, 9) + rank, "b%d" % i: torch.randn(300), "none%d" % i: None}
        blocks.append(g)
        b.add(g)
    b.finish()
    assert b.buckets_launched >= 2
    flat = torch.cat([t.reshape(-1) for g in blocks for t in g.values() if t is not None])
    gathered = [torch.zeros_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    same = all(torch.equal(gathered[0], t) for t in gathered)
    out[rank] = (same, float(flat.mean()))
    dist.destroy_process_group()


def test_gradient_buckets_average_across_two_ranks():
    world = 2
    port = 29500 + (os.getpid() % 2000)
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    assert out[0][0] and out[1][0]           # every rank ends with identical (averaged) gradients
    assert abs(out[0][1] - out[1][1]) < 1e-7
    # the rank-dependent +rank offset averages to 0.5 on the weight blocks
    assert 0.1 < out[0][1] < 0.4


def _worker_persistent(rank, world, port, out):
    sys.path.insert(0, ROOT)
    import b3d  # noqa: F401
    from unet3d_b200.parallel import GradientBuckets
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    b = GradientBuckets(bucket_mb=0.002)
    names = [("w%d" % i, (16, 9)) for i in range(6)] + [("b%d" % i, (40,)) for i in range(6)]
    ok = True
    for step in range(3):
        torch.manual_seed(100 * step + rank)
        emitted = {}
        for i in range(0, len(names), 3):                       # four "blocks" of three gradients each
            block = {}
            for name, shape in names[i:i + 3]:
                sink = b.sink(name, shape)                      # None in the first pass, a bucket slice afterwards
                g = torch.randn(shape) + rank
                if sink is not None and name.startswith("w"):   # "weight-gradient kernels" write into the bucket directly
                    sink.copy_(g)
                    g_out = sink
                else:
                    g_out = g.clone()
                block[name] = g_out
                emitted[name] = g
            b.add(block)
            if step > 0:
                ok = ok and all(v.data_ptr() == b.view(k).data_ptr() for k, v in block.items())   # add() swapped in bucket views
        b.finish()
        final = {k: b.view(k).clone() for k, _ in names} if step > 0 else None
        if step == 0:
            ok = ok and b._layout is not None and len(b._flat) >= 2
            continue
        # expected: mean over ranks of what each rank emitted
        for k, _ in names:
            mine = emitted[k]
            gathered = [torch.zeros_like(mine) for _ in range(world)]
            dist.all_gather(gathered, mine)
            ok = ok and torch.allclose(final[k], sum(gathered) / world, atol=1e-6)
        fv = b.fresh_views({k: b.view(k) for k, _ in names})
        ok = ok and all(fv[k].data_ptr() == b.view(k).data_ptr() for k, _ in names)
    out[rank] = bool(ok)
    dist.destroy_process_group()


def test_persistent_buckets_allreduce_in_place_across_two_ranks():
    """Second and later passes: fixed bucket slices (sink), one multi-tensor copy per block for the rest, in-place all-reduce,
    autograd-facing views alias the buckets."""
    world = 2
    port = 31500 + (os.getpid() % 2000)
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker_persistent, args=(world, port, out), nprocs=world, join=True)
    assert out[0] and out[1]


def test_single_process_is_a_noop():
    sys.path.insert(0, ROOT)
    import b3d  # noqa: F401
    from unet3d_b200.parallel import GradientBuckets
    b = GradientBuckets()
    g = {"w": torch.ones(4)}
    b.add(g)
    b.finish()
    assert torch.equal(g["w"], torch.ones(4)) and b.buckets_launched == 0
