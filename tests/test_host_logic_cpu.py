"""Host-side logic of the package that needs no GPU: the weight-gradient zero-padding rule, the per-parameter packed-weight
cache and its invalidation, and the C-side planners' invariants mirrored in Python."""
import torch

import b3d  # noqa: F401  (registers the package as `unet3d_b200`)
from unet3d_b200 import functional, ops


def test_pad_w16_is_exact_for_the_weight_gradient():
    """ops._pad_w16: zero-padding W to a multiple of 16 does not change  dW = sum_v X[v + tap] * dY[v]  (the padded dY
    columns are zero, the padded X columns equal the convolution's own zero padding)."""
    g = torch.Generator().manual_seed(0)
    x = torch.randn(1, 3, 5, 24, 8, generator=g)        # NDHWC, W = 24
    dy = torch.randn(1, 3, 5, 24, 4, generator=g)
    xp, dyp = ops._pad_w16(x, dy, 1)
    assert xp.shape[3] == 32 and dyp.shape[3] == 32 and xp.shape[:3] == x.shape[:3]
    assert torch.equal(xp[:, :, :, :24], x) and float(xp[:, :, :, 24:].abs().max()) == 0.0

    def wgrad(a, b):
        w = torch.zeros(b.shape[-1], a.shape[-1], 3, 3, 3, requires_grad=True)
        torch.nn.functional.conv3d(a.permute(0, 4, 1, 2, 3), w, padding=1).backward(b.permute(0, 4, 1, 2, 3))
        return w.grad
    assert torch.allclose(wgrad(x, dy), wgrad(xp, dyp), atol=1e-5)
    # transposed conv: dy is twice as fine as x
    dyt = torch.randn(1, 6, 10, 48, 4, generator=g)
    xq, dyq = ops._pad_w16(x, dyt, 2)
    assert xq.shape[3] == 32 and dyq.shape[3] == 64
    # multiples of 16 and tiny planes are left alone
    a = torch.randn(1, 2, 2, 16, 8); b = torch.randn(1, 2, 2, 16, 4)
    assert ops._pad_w16(a, b, 1)[0] is a
    a = torch.randn(1, 2, 2, 6, 8); b = torch.randn(1, 2, 2, 6, 4)
    assert ops._pad_w16(a, b, 1)[0] is a


def test_packed_weight_cache_lives_on_the_parameter(monkeypatch):
    calls = []

    def fake_pack(w, mode):
        calls.append((id(w), mode))
        return torch.zeros(1), 16, 16
    monkeypatch.setattr(ops, "pack_weight", fake_pack)
    p = torch.nn.Parameter(torch.randn(4, 4, 3, 3, 3))
    functional.packed(p, 0); functional.packed(p, 0)
    assert len(calls) == 1                       # cached
    functional.packed(p, 1)
    assert len(calls) == 2                       # per mode
    with torch.no_grad():
        p.add_(1.0)                              # in-place update bumps the version counter
    functional.packed(p, 0)
    assert len(calls) == 3
    functional.clear_pack_cache()                # a replayed CUDA graph changes parameters behind the version counters
    functional.packed(p, 0)
    assert len(calls) == 4
    q = torch.nn.Parameter(p.detach().clone())   # another parameter never sees p's cache
    functional.packed(q, 0)
    assert len(calls) == 5 and "_b3d_pack" in q.__dict__ and q.__dict__["_b3d_pack"] is not p.__dict__["_b3d_pack"]


def test_optimizer_step_invalidates_the_pack_cache(monkeypatch):
    """torch's fused AdamW updates parameters WITHOUT bumping their version counters (checked here), so the packed-weight
    cache is also invalidated by a global optimizer post-step hook — for every torch optimizer, fused or not."""
    calls = []

    def fake_pack(w, mode):
        calls.append(mode)
        return torch.zeros(1), 16, 16
    monkeypatch.setattr(ops, "pack_weight", fake_pack)
    for kwargs in ({"fused": True}, {"foreach": True}, {}):
        p = torch.nn.Parameter(torch.randn(4, 4, 3, 3, 3))
        opt = torch.optim.AdamW([p], lr=1e-3, **kwargs)
        calls.clear()
        functional.packed(p, 0); functional.packed(p, 0)
        assert len(calls) == 1
        p.grad = torch.randn_like(p)
        before = p.detach().clone()
        opt.step()
        assert not torch.equal(before, p.detach())
        functional.packed(p, 0)
        assert len(calls) == 2, "packed weights survived an optimizer step (%s)" % (kwargs,)
    sgd_p = torch.nn.Parameter(torch.randn(4, 4, 1, 1, 1))
    sgd = torch.optim.SGD([sgd_p], lr=0.1)
    calls.clear()
    functional.packed(sgd_p, 0)
    sgd_p.grad = torch.ones_like(sgd_p)
    sgd.step()
    functional.packed(sgd_p, 0)
    assert len(calls) == 2


def test_zeros_scratch_and_side_branch_are_inert_on_cpu():
    """ops.zeros_scratch falls back to torch.zeros off-GPU; ops.side_branch(False, ...) is a no-op context."""
    z = ops.zeros_scratch((2, 3, 2), torch.float64, torch.device("cpu"))
    assert z.shape == (2, 3, 2) and z.dtype == torch.float64 and float(z.abs().sum()) == 0.0
    with ops.side_branch(False, z) as br:
        z += 1
    br.join()
    assert float(z.sum()) == 12.0


def test_seed0_init_is_bit_identical_to_the_reference(golden):
    """Same constructor tree / registration order / init calls as main.py:102-153 => the same torch seed yields bit-identical
    initial parameters and buffers (hashes of the reference's own state_dict, tests/golden/make_golden.py:init_case)."""
    import hashlib
    import unet3d_b200 as U
    torch.manual_seed(0)
    m = U.UNet3D(4, 4, features=[16, 32, 64, 128, 256])
    want = golden["init_small"]
    sd = m.state_dict()
    assert list(sd.keys()) == list(want.keys())
    for k, v in sd.items():
        assert hashlib.sha1(v.numpy().tobytes()).hexdigest()[:16] == want[k], k
