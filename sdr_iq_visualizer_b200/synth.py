"""Synthetic IQ workloads of BASELINE.json / SURVEY.md 8(d): QPSK + CW tone + AWGN.

Input generators only (numpy on the host); nothing here is on the compute path.  bench.py and the
examples use them; the test-side checker keeps its own copy.
"""
from __future__ import annotations

import numpy as np


def synth_iq(n: int, seed: int, snr_db: float = 20.0, tone_cycles_per_sample: float = 0.2,
             sps: int = 8, tone_amp: float = 0.5) -> np.ndarray:
    """QPSK (rectangular pulses, ``sps`` samples/symbol, amplitude 1) + CW tone + complex AWGN; complex128."""
    rng = np.random.default_rng(seed)
    nsym = (n + sps - 1) // sps
    bits = rng.integers(0, 2, size=(nsym, 2))
    sym = ((2 * bits[:, 0] - 1) + 1j * (2 * bits[:, 1] - 1)) / np.sqrt(2.0)
    out = np.repeat(sym, sps)[:n]
    t = np.arange(n, dtype=np.float64)
    out = out + tone_amp * np.exp(2j * np.pi * tone_cycles_per_sample * t)
    sigma2 = 10.0 ** (-snr_db / 10.0)
    out += np.sqrt(sigma2 / 2.0) * (rng.standard_normal(n) + 1j * rng.standard_normal(n))
    return out


def to_ci16(x: np.ndarray, gain: float = 1024.0, clip: int = 2047) -> np.ndarray:
    """12-bit-range interleaved int16 I,Q (Pluto style): round(x*gain) clipped to +-clip."""
    iq = np.empty(2 * len(x), dtype=np.int16)
    iq[0::2] = np.clip(np.rint(x.real * gain), -clip, clip).astype(np.int16)
    iq[1::2] = np.clip(np.rint(x.imag * gain), -clip, clip).astype(np.int16)
    return iq


def tiled_ci16(n: int, seed: int, block_log2: int = 22, tone_cycles_per_sample: float = 1500.37 / 4096) -> np.ndarray:
    """``n`` int16 IQ samples: one 2^block_log2-sample synthetic block tiled (cheap to generate at 61.44 MS)."""
    base = to_ci16(synth_iq(1 << block_log2, seed=seed, tone_cycles_per_sample=tone_cycles_per_sample))
    reps = -(-n // (1 << block_log2))
    return np.tile(base.reshape(-1, 2), (reps, 1))[:n].reshape(-1)
