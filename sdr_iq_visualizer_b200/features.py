"""Classifier measurements on the GPU (kernel K3, csrc/spx_features.cu) -- host binding.

One call measures a batch of spectra: noise floor (exact 20th percentile), peak, SNR, occupied
bandwidth edges, flatness, kurtosis and the greedy peak pick of
/root/reference/app/processing/classifier.py:45-58,163-219.  Indices come back as integers; turning
them into Hz with the caller's ``freqs`` is scalar host work done in ``classifier.py``.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional

import numpy as np

from . import _native as nat

_FIELDS = [f[0] for f in nat.spx_features._fields_ if f[0] != "reserved"]


def _to_dict(f: nat.spx_features, peaks: Optional[np.ndarray]) -> dict:
    d = {k: getattr(f, k) for k in _FIELDS}
    d["peaks"] = [] if peaks is None else [int(v) for v in peaks[: f.peaks_stored]]
    return d


def measure_batch(power_db, n: Optional[int] = None, batch: Optional[int] = None, stride: Optional[int] = None,
                  device: int = 0, want_peaks: bool = True, stream: int = 0, drops=(3.0, 10.0, 20.0),
                  peak_threshold_db: Optional[float] = None, min_distance_bins: int = 0) -> List[dict]:
    """Measure ``batch`` spectra of ``n`` bins.  ``power_db`` is a numpy array ([n] or [batch, n],
    float32/float64), a DeviceArray or a CUDA tensor (then n/batch/stride describe it)."""
    nat.require_device()
    if isinstance(power_db, np.ndarray) or not (isinstance(power_db, nat.DeviceArray) or hasattr(power_db, "data_ptr")):
        a = np.asarray(power_db)
        if a.dtype not in (np.float32, np.float64):
            a = a.astype(np.float64)
        a = np.ascontiguousarray(a)
        if a.ndim == 1:
            a = a[None, :]
        batch, n = a.shape
        stride = n
        dtype = 0 if a.dtype == np.float32 else 1
        ptr, mem = a.ctypes.data, nat.MEM_HOST
    else:
        ptr, mem = nat.as_ptr(power_db)
        if isinstance(power_db, nat.DeviceArray):
            dt, shape = power_db.dtype, power_db.shape
        else:
            import torch  # plumbing only
            dt = {torch.float32: np.dtype(np.float32), torch.float64: np.dtype(np.float64)}[power_db.dtype]
            shape = tuple(power_db.shape)
        dtype = 0 if dt == np.float32 else 1
        if n is None:
            n = shape[-1]
        if batch is None:
            batch = int(np.prod(shape[:-1])) if len(shape) > 1 else 1
        if stride is None:
            stride = n
    if batch == 0:
        return []
    out = (nat.spx_features * batch)()
    # strict local maxima are >= 2 bins apart: at most (n+1)//2 of them survive a min distance of 1 or 2, at most
    # n//min_dist + 1 otherwise (the default distance is max(3, n//300))
    md = int(min_distance_bins)
    cap = (max(1, (n + 1) // 2) if 0 < md < 3 else max(1, n // 3 + 2)) if want_peaks else 0
    peaks = np.zeros((batch, cap), np.int32) if want_peaks else None
    opts = nat.spx_feature_opts()
    opts.drop_db[0], opts.drop_db[1], opts.drop_db[2] = (float(d) for d in drops)
    opts.use_peak_threshold = 0 if peak_threshold_db is None else 1
    opts.peak_threshold_db = 0.0 if peak_threshold_db is None else float(peak_threshold_db)
    opts.min_distance_bins = int(min_distance_bins)
    nat.check(nat.lib().spx_classify_features(device, mem, ptr, dtype, int(n), int(batch), int(stride), out,
                                              peaks.ctypes.data if want_peaks else None, cap, C.byref(opts),
                                              stream or None))
    res = [_to_dict(out[b], peaks[b] if want_peaks else None) for b in range(batch)]
    if want_peaks:
        for d in res:
            if d["peaks_stored"] != d["peak_count"]:   # cannot happen with the caps above; never hand back a cut list
                raise nat.SpectralError(nat.E_INVALID,
                                        f"peak list truncated: {d['peaks_stored']} of {d['peak_count']} stored")
    return res


def measure(power_db, device: int = 0, **opts) -> dict:
    return measure_batch(power_db, device=device, **opts)[0]


class FeatureQueue:
    """Enqueue-only feature measurement for pipelines that must not stall the host (one Welch block after the
    other on one stream): ``enqueue`` launches the kernel on device data and an asynchronous copy of the result
    structs into pinned memory; ``results`` is read after the stream has been synchronised."""

    def __init__(self, batch: int = 1, device: int = 0, stream: int = 0, slots: int = 1):
        nat.require_device()
        self.batch, self.device, self.stream, self.slots = int(batch), int(device), stream or None, int(slots)
        self._sz = C.sizeof(nat.spx_features) * self.batch
        self._dev = nat.DeviceArray((self.slots * self._sz,), np.uint8, device)
        self._host = nat.pinned_empty(self.slots * self._sz, np.uint8)
        self._opts = nat.spx_feature_opts()
        self._opts.drop_db[0], self._opts.drop_db[1], self._opts.drop_db[2] = 3.0, 10.0, 20.0
        self._next = 0

    def enqueue(self, power_db, n: int, stride: Optional[int] = None, dtype=np.float64) -> int:
        """Measure ``batch`` device-resident spectra of ``n`` bins; returns the slot the results will land in."""
        slot = self._next % self.slots
        self._next += 1
        ptr, mem = nat.as_ptr(power_db)
        if mem != nat.MEM_DEVICE:
            raise ValueError("FeatureQueue needs device-resident spectra")
        d_out = self._dev.ptr + slot * self._sz
        nat.check(nat.lib().spx_classify_features_dev(self.device, ptr, 0 if np.dtype(dtype) == np.float32 else 1, int(n),
                                                      self.batch, int(stride or n), d_out, None, 0, C.byref(self._opts),
                                                      self.stream))
        nat.check(nat.lib().spx_memcpy_d2h_async(self.device, self._host.ctypes.data + slot * self._sz, d_out, self._sz,
                                                 self.stream))
        return slot

    def results(self, slot: int = 0) -> List[dict]:
        """Feature dicts of a slot (call after synchronising the stream the work was enqueued on)."""
        arr = (nat.spx_features * self.batch).from_buffer_copy(bytes(self._host[slot * self._sz:(slot + 1) * self._sz]))
        return [_to_dict(arr[b], None) for b in range(self.batch)]
