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


def test_single_process_is_a_noop():
    sys.path.insert(0, ROOT)
    import b3d  # noqa: F401
    from unet3d_b200.parallel import GradientBuckets
    b = GradientBuckets()
    g = {"w": torch.ones(4)}
    b.add(g)
    b.finish()
    assert torch.equal(g["w"], torch.ones(4)) and b.buckets_launched == 0
