#!/usr/bin/env python3
"""CPU baselines of SURVEY.md 8(d), timed on the host this runs on (run it on the GPU box so that the numbers sit next
to the GPU ones): (i) the reference's per-buffer lines (streamer.py:119-121) in a loop, (ii) the float64 numpy STFT of
the checker, (iii) the classifier measurements of the checker on one spectrum, (iv) np.histogram2d.  One core each
(numpy's pocketfft is single-threaded), best of 3.  Reported baselines only -- no speed-up target."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import classifier_ref as cref, spectral_ref as sref  # noqa: E402


def best(fn, reps=3):
    t = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        t.append(time.perf_counter() - t0)
    return min(t)


out = {"host_cores": os.cpu_count(), "numpy": np.__version__}
rng = np.random.default_rng(0)
n, nbuf = 4096, 2000
bufs = (rng.integers(-2047, 2048, (nbuf, n)) + 1j * rng.integers(-2047, 2048, (nbuf, n))).astype(np.complex128)


def per_buffer():
    for b in bufs:
        sref.stream_frame(b, 61.44e6, 2.4e9)


t = best(per_buffer)
out["i_per_buffer_stream_lines"] = {"us_per_4096_buffer": round(t / nbuf * 1e6, 1), "Msamples_per_s": round(nbuf * n / t / 1e6, 2)}
x = sref.to_ci16(sref.synth_iq(1 << 23, seed=2))
t = best(lambda: sref.stft_power_rows(sref.as_complex128(x, sref.FMT_CI16), 4096, 1024, "hann"), reps=2)
out["ii_float64_stft_4096_hann_75pct"] = {"Msamples_per_s": round((1 << 23) / t / 1e6, 2), "sample": "2^23 int16 IQ samples"}
f = sref.freq_axis(4096, 61.44e6, 2.4e9)
p = rng.normal(-80, 3, 4096); p[1000] += 40
t = best(lambda: [cref.features(f, p) for _ in range(200)])
out["iii_classifier_measurements_4096_bins"] = {"us_per_spectrum": round(t / 200 * 1e6, 1)}
xc = (0.7 * (rng.standard_normal(1 << 24) + 1j * rng.standard_normal(1 << 24))).astype(np.complex64)
t = best(lambda: sref.iq_hist2d(xc, 4.0, 256), reps=2)
out["iv_histogram2d_2^24_samples"] = {"s": round(t, 3), "Msamples_per_s": round((1 << 24) / t / 1e6, 2)}
print(json.dumps(out))
