#!/usr/bin/env python3
"""Randomised exactness check of spx_iq_hist2d against np.histogram2d (the reference semantics): bins, range, scale, length,
format, planted edge values.  python tools/fuzz_hist2d.py [cases] [seed]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sdr_iq_visualizer_b200 import _native as nat, timedomain as td

cases = int(sys.argv[1]) if len(sys.argv) > 1 else 200
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
bad = 0
for c in range(cases):
    bins = int(rng.choice([1, 2, 3, 7, 16, 64, 100, 128, 192, 255, 256, 300, 318, 319, 400]))
    n = int(rng.choice([1, 5, 1000, 32759, 32761, 70001, 262144, 300007]))
    if rng.random() < 0.5:   # ci16
        scale = float(rng.choice([1.0, 2.0**-15, 2.0**-11, 0.5, 1e-3, 3.0, 1.0 / 2048]))
        r = float(rng.choice([2048.0, 1000.0, 32768.0, 100.0, 4096.0, 777.7])) * scale * float(rng.choice([1.0, 1.0, 0.5, 2.0]))
        raw = rng.integers(-32768, 32768, 2 * n).astype(np.int16)
        if rng.random() < 0.5:
            raw = (rng.normal(0, r / scale / 3, 2 * n)).clip(-32768, 32767).astype(np.int16)
        k = int(round(r / scale))
        sp = np.array([kk for kk in (-k, k, -k - 1, k + 1, 0, -32768, 32767) if -32768 <= kk <= 32767], np.int16)
        m = min(sp.size, n)
        raw[0:2 * m:2] = sp[:m]; raw[1:2 * m:2] = sp[:m][::-1]
        got = td.iq_hist2d(raw, r, bins, in_fmt=nat.FMT_CI16, in_scale=scale)
        i = raw[0::2].astype(np.float64) * scale; q = raw[1::2].astype(np.float64) * scale
        desc = f"ci16 scale {scale} r {r}"
    else:
        r = float(rng.choice([4.0, 1.0, 0.7, 3.3, 1e-3, 1e4, 2.5]))
        x = (rng.normal(0, r / 3, n) + 1j * rng.normal(0, r / 3, n)).astype(np.complex64)
        if rng.random() < 0.3:
            x = (rng.uniform(-1.2 * r, 1.2 * r, n) + 1j * rng.uniform(-1.2 * r, 1.2 * r, n)).astype(np.complex64)
        edges = np.linspace(-r, r, bins + 1)
        sp = np.concatenate([edges, [np.nextafter(r, np.inf), np.nextafter(-r, -np.inf), np.nan, np.inf, -np.inf, 1e30, -1e30]]).astype(np.float32)
        m = min(sp.size, n)
        x[:m] = sp[:m] + 1j * sp[:m][::-1]
        got = td.iq_hist2d(x, r, bins)
        i = x.real.astype(np.float64); q = x.imag.astype(np.float64)
        desc = f"cf32 r {r}"
    ok = np.isfinite(i) & np.isfinite(q)
    want = np.histogram2d(i[ok], q[ok], bins=bins, range=[[-r, r], [-r, r]])[0].astype(np.uint32)
    if not np.array_equal(got, want):
        bad += 1
        print("MISMATCH", desc, "bins", bins, "n", n, "diff bins", int((got != want).sum()), "sum", int(got.sum()), int(want.sum()), flush=True)
print(f"fuzz_hist2d: {cases} cases, {bad} mismatches")
sys.exit(1 if bad else 0)
