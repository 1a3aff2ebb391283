#!/usr/bin/env python3
"""Short device-resident runs of the non-STFT kernels for ncu: python tools/profile_misc.py features|hist|headline"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sdr_iq_visualizer_b200 import _native as nat, features, spectral as sp, timedomain as td  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "features"
rng = np.random.default_rng(1)
if what == "features":
    p = rng.normal(-80, 3, (1, 4096)); p[0, 1000] += 40
    dp = nat.DeviceArray.from_host(p)
    for _ in range(3):
        features.measure_batch(dp, n=4096, batch=1, want_peaks=False)
elif what == "hist":
    x = (0.7 * (rng.standard_normal(1 << 24) + 1j * rng.standard_normal(1 << 24))).astype(np.complex64)
    dx = nat.DeviceArray.from_host(x)
    dh = nat.DeviceArray((256, 256), np.uint32, zero=True)
    for _ in range(3):
        td.iq_hist2d(dx, 4.0, 256, out=dh)
        td.frame_stats(dx, 4096, 4096)
    nat.device_sync(0)
else:
    L = 61_440_000
    x = nat.DeviceArray.from_host(rng.standard_normal(2 * L).astype(np.float32).view(np.complex64))
    pl = sp.SpectralPlan(4096, 4096, "hann", sp.FMT_CF32)
    db = nat.DeviceArray((L // 4096, 4096), np.float32)
    res, ms = pl.time_stft(x, warmup=2, iters=3, db_rows=db)
    print("headline ms", ms)
print("ok", what)
