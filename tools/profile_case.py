#!/usr/bin/env python3
"""One short device-resident run of a named case, for ncu.  python tools/profile_case.py <case> [variant] [samples]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sdr_iq_visualizer_b200 import _native as nat, spectral as sp  # noqa: E402

case = sys.argv[1] if len(sys.argv) > 1 else "headline"
variant = int(sys.argv[2]) if len(sys.argv) > 2 else 0
L = int(sys.argv[3]) if len(sys.argv) > 3 else 1 << 24
rng = np.random.default_rng(1)
if case == "c2":
    x = nat.DeviceArray.from_host(rng.integers(-2047, 2048, size=2 * L, dtype=np.int16))
    pl = sp.SpectralPlan(4096, 1024, "hann", sp.FMT_CI16, variant=variant)
    F = pl.frame_count(L)
    kw = dict(wf_rows=nat.DeviceArray((F, 4096), np.uint8), welch=nat.DeviceArray((1, 4096), np.float64),
              maxhold=nat.DeviceArray((1, 4096), np.float32), vmin=20.0, vmax=130.0)
elif case == "c5":
    x = nat.DeviceArray.from_host(rng.standard_normal(2 * L).astype(np.float32).view(np.complex64))
    pl = sp.SpectralPlan(65536, 32768, "hann", sp.FMT_CF32, variant=variant)
    F = pl.frame_count(L)
    kw = dict(wf_rows=nat.DeviceArray((F, 65536), np.uint8), welch=nat.DeviceArray((1, 65536), np.float64),
              maxhold=nat.DeviceArray((1, 65536), np.float32), vmin=-20.0, vmax=110.0)
else:
    x = nat.DeviceArray.from_host(rng.standard_normal(2 * L).astype(np.float32).view(np.complex64))
    pl = sp.SpectralPlan(4096, 4096, "hann", sp.FMT_CF32, variant=variant)
    F = pl.frame_count(L)
    kw = dict(db_rows=nat.DeviceArray((F, 4096), np.float32))
res, ms = pl.time_stft(x, warmup=2, iters=3, **kw)
print(case, variant, "frames", F, "ms", ms, "GS/s", L / (np.median(ms) * 1e-3) / 1e9)
