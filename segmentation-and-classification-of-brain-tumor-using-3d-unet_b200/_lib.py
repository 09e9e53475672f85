"""ctypes binding of libb3d.so — the C-ABI boundary of the hot path (see include/b3d.h).

There is deliberately NO fallback: if the shared object is missing or the device is not sm_100, every op raises.
"""
import ctypes
import os

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

c_int, c_ll, c_vp, c_sz = ctypes.c_int, ctypes.c_longlong, ctypes.c_void_p, ctypes.c_size_t


class B3DError(RuntimeError):
    pass


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(HERE, "libb3d.so")
        if not os.path.exists(path):
            raise B3DError(
                "libb3d.so not built (%s). Run `python __graft_entry__.py build`; there is no CPU/cuDNN fallback." % path)
        _LIB = ctypes.CDLL(path)
        _LIB.b3d_last_error_string.restype = ctypes.c_char_p
        for fn in ("b3d_conv_fprop_workspace_bytes", "b3d_convT2_dgrad_workspace_bytes", "b3d_conv_wgrad_workspace_bytes",
                   "b3d_convT2_wgrad_workspace_bytes"):
            getattr(_LIB, fn).restype = ctypes.c_size_t
    return _LIB


def check(rc):
    if rc != 0:
        raise B3DError("libb3d error %d: %s" % (rc, lib().b3d_last_error_string().decode()))


def stream_ptr():
    return c_vp(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    if t is None:
        return c_vp(0)
    return c_vp(t.data_ptr())


_ERR_FLAG = {}


def err_flag(device):
    """Device int the kernels' watchdogs write a code into before trapping."""
    key = str(device)
    if key not in _ERR_FLAG:
        _ERR_FLAG[key] = torch.zeros(1, dtype=torch.int32, device=device)
    return _ERR_FLAG[key]


_DEVICE_OK = {}


def require_device(device):
    key = str(device)
    if key not in _DEVICE_OK:
        if not torch.cuda.is_available():
            raise B3DError("CUDA device required: the b200 path has no CPU fallback")
        with torch.cuda.device(device):
            check(lib().b3d_check_device())
        _DEVICE_OK[key] = True


def set_ordered_issue(on):
    """Bit-reproducible forward / input gradients: ONE MMA-issuing thread in the z-marching conv kernel instead of two ping-pong
    issuers, whose accumulation order jitters by an fp32 ulp from run to run.  0 / False: two issuers; 1 / True: one issuer;
    2: one issuer fed by a scout warp that does the barrier waits and descriptor arithmetic.  Returns the previous mode."""
    return int(lib().b3d_set_ordered_issue(c_int(int(on))))


def set_reserved_sms(n):
    """Size every later grid of the library for (SM count - n) SMs, leaving n SMs to NCCL's channel CTAs (data parallel:
    set it to NCCL_MAX_CTAS before the step is captured).  Returns the previous reservation."""
    return int(lib().b3d_set_reserved_sms(c_int(int(n))))
