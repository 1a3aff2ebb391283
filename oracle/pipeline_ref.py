"""CPU oracle for the whole config-2 style step -- TEST / BASELINE INFRASTRUCTURE, NOT PRODUCT CODE.

One call does, in float64 numpy, what one GPU step does: unpack int16 IQ -> Hann window ->
overlapped FFT frames (np.fft.fft, the call at /root/reference/app/sdr/streamer.py:119) ->
|X|^2 -> Welch sum + max-hold -> 20*log10(|X|+1e-12) (streamer.py:121) -> fftshift -> uint8 rows,
then the classifier measurements (classifier.py:45-58) on the Welch PSD.  Used by bench.py's
``cpu_baseline`` / ``--impl reference`` legs and by tests; never by the product package.
"""
from __future__ import annotations

import numpy as np

from . import classifier_ref as cref
from . import spectral_ref as sref


def stft_block(x, nfft, hop, kind, in_fmt, scale, vmin, vmax, want_rows=True, chunk=256):
    """Returns (welch_sum[N], maxhold[N], u8 rows [F,N] or None, F) for one contiguous block."""
    xs = sref.as_complex128(x, in_fmt, scale)
    fr = sref.frames(xs, nfft, hop)
    w = sref.window(kind, nfft)
    F = fr.shape[0]
    acc = np.zeros(nfft)
    mx = np.zeros(nfft)
    rows = np.empty((F, nfft), np.uint8) if want_rows else None
    for i in range(0, F, chunk):
        X = np.fft.fftshift(np.fft.fft(fr[i:i + chunk] * w, axis=1), axes=1)
        P = X.real**2 + X.imag**2
        acc += P.sum(axis=0)
        np.maximum(mx, P.max(axis=0), out=mx)
        if want_rows:
            rows[i:i + chunk] = sref.waterfall_u8(20 * np.log10(np.sqrt(P) + 1e-12), vmin, vmax)
    return acc, mx, rows, F


def c2_step(x_ci16, nfft=4096, hop=1024, kind="hann", sample_rate=61.44e6, center_freq=2.4e9, vmin=20.0, vmax=130.0,
            want_rows=True):
    acc, mx, rows, F = stft_block(x_ci16, nfft, hop, kind, sref.FMT_CI16, 1.0, vmin, vmax, want_rows)
    w = sref.window(kind, nfft)
    pxx = acc / F / (sample_rate * np.sum(w**2))
    pxx_db = 10 * np.log10(pxx)
    freqs = sref.freq_axis(nfft, sample_rate, center_freq)
    feats = cref.features(freqs, pxx_db)
    return {"n_frames": F, "welch_acc": acc, "maxhold": mx, "wf_rows": rows, "pxx_db": pxx_db, "features": feats}


def _worker(args):
    x, nfft, hop, kind, vmin, vmax = args
    acc, mx, rows, F = stft_block(x, nfft, hop, kind, sref.FMT_CI16, 1.0, vmin, vmax, True)
    return acc, mx, F, int(rows[::64, ::64].sum())  # rows are produced; only a checksum travels back


def c2_step_parallel(x_ci16, pool, n_workers, nfft=4096, hop=1024, kind="hann", vmin=20.0, vmax=130.0):
    """Frame-block sharding with (N - hop)-sample halos across a multiprocessing pool (the CPU
    analogue of the multi-GPU split of SURVEY.md 8(e))."""
    L = x_ci16.size // 2
    F = sref.frame_count(L, nfft, hop)
    per = -(-F // n_workers)
    jobs = []
    for k in range(n_workers):
        f0, f1 = k * per, min(F, (k + 1) * per)
        if f0 >= f1:
            break
        jobs.append((x_ci16[2 * f0 * hop: 2 * ((f1 - 1) * hop + nfft)], nfft, hop, kind, vmin, vmax))
    parts = pool.map(_worker, jobs)
    acc = sum(p[0] for p in parts)
    mx = np.maximum.reduce([p[1] for p in parts])
    return acc, mx, sum(p[2] for p in parts)
