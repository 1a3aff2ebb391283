"""Time-domain and constellation views on the GPU (kernel K4, csrc/spx_timedomain.cu).

``iq_hist2d`` replaces the random 2000-point constellation scatter of the reference
(/root/reference/app/dashboard/callbacks.py:199-214) with an exact 2-D density
(np.histogram2d semantics); ``frame_stats`` gives the per-frame mean / peak power of
SURVEY.md A9.  No CPU fallback.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _native as nat
from ._native import FMT_CF32, FMT_CI16, MEM_HOST, DeviceArray


def _input(x, in_fmt):
    if isinstance(x, (DeviceArray, nat.DeviceView)) or (hasattr(x, "data_ptr") and hasattr(x, "is_cuda")):
        ptr, mem = nat.as_ptr(x)
        nbytes = x.nbytes if isinstance(x, (DeviceArray, nat.DeviceView)) else x.numel() * x.element_size()
        return x, ptr, mem, nbytes // (4 if in_fmt == FMT_CI16 else 8)
    a = np.asarray(x)
    if in_fmt == FMT_CI16:
        if a.dtype != np.int16:
            raise TypeError("ci16 input must be int16 interleaved I,Q")
        a = np.ascontiguousarray(a)
        return a, a.ctypes.data, MEM_HOST, a.size // 2
    a = np.ascontiguousarray(a, dtype=np.complex64)
    return a, a.ctypes.data, MEM_HOST, a.size


def iq_hist2d(x, r: float, bins: int = 256, in_fmt: int = FMT_CF32, in_scale: float = 1.0, out=None,
              accumulate: bool = False, device: int = 0, stream: int = 0):
    """uint32 [bins, bins] counts, H[i, j] with i <-> I and j <-> Q."""
    nat.require_device()
    keep, ptr, mem, n = _input(x, in_fmt)
    if out is None:
        out = np.zeros((bins, bins), np.uint32) if mem == MEM_HOST else DeviceArray((bins, bins), np.uint32, device, zero=True)
    optr, omem = nat.as_ptr(out)
    if omem != mem:
        raise ValueError("output must live where the input lives")
    from .spectral import _check_buffer
    _check_buffer(out, (bins, bins), np.uint32, "out")
    nat.check(nat.lib().spx_iq_hist2d(device, mem, ptr, in_fmt, float(in_scale), n, float(r), int(bins), optr,
                                      1 if accumulate else 0, stream or None))
    return out


def frame_stats(x, frame_len: int, hop: int = 0, in_fmt: int = FMT_CF32, in_scale: float = 1.0, device: int = 0,
                stream: int = 0, out=None):
    """(mean_pow, peak_pow): float32 [F] each, F = (n - frame_len)//hop + 1.  ``out=(mean, peak)`` reuses caller
    buffers living where the input lives."""
    nat.require_device()
    hop = hop or frame_len
    keep, ptr, mem, n = _input(x, in_fmt)
    F = nat.frame_count(n, frame_len, hop)
    if out is not None:
        mean, peak = out
        from .spectral import _check_buffer
        for name, buf in (("mean", mean), ("peak", peak)):
            if nat.as_ptr(buf)[1] != mem:
                raise ValueError(f"out {name} must live where the input lives")
            _check_buffer(buf, (F,), np.float32, "out " + name)
    elif mem == MEM_HOST:
        mean, peak = np.zeros(F, np.float32), np.zeros(F, np.float32)
    else:
        mean, peak = DeviceArray((F,), np.float32, device), DeviceArray((F,), np.float32, device)
    nf = C.c_int64(0)
    nat.check(nat.lib().spx_frame_stats(device, mem, ptr, in_fmt, float(in_scale), n, int(frame_len), int(hop),
                                        nat.as_ptr(mean)[0], nat.as_ptr(peak)[0], C.byref(nf), stream or None))
    return mean, peak
