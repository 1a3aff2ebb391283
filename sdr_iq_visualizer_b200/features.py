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
    cap = max(1, n // 3 + 2) if want_peaks else 0
    peaks = np.zeros((batch, cap), np.int32) if want_peaks else None
    opts = nat.spx_feature_opts()
    opts.drop_db[0], opts.drop_db[1], opts.drop_db[2] = (float(d) for d in drops)
    opts.use_peak_threshold = 0 if peak_threshold_db is None else 1
    opts.peak_threshold_db = 0.0 if peak_threshold_db is None else float(peak_threshold_db)
    opts.min_distance_bins = int(min_distance_bins)
    nat.check(nat.lib().spx_classify_features(device, mem, ptr, dtype, int(n), int(batch), int(stride), out,
                                              peaks.ctypes.data if want_peaks else None, cap, C.byref(opts),
                                              stream or None))
    return [_to_dict(out[b], peaks[b] if want_peaks else None) for b in range(batch)]


def measure(power_db, device: int = 0, **opts) -> dict:
    return measure_batch(power_db, device=device, **opts)[0]
