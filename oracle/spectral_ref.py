"""CPU oracle for the spectral hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

This module is a float64 numpy restatement of what the reference computes on the
path  IQ -> (window) -> FFT -> |X|^2 / dB -> fftshift -> Welch / max-hold -> waterfall,
plus the time-domain views (frame power statistics, I/Q density histogram).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu-baseline /
``--impl reference`` legs may import it.  The product package
(``sdr_iq_visualizer_b200``) never does; it fails loudly without its CUDA library.

Parity pinning (see DESIGN.md "Oracle"):
  * stream path  -- pinned by running the reference's own ``SDRDataStreamer._stream_data``
    (``/root/reference/app/sdr/streamer.py:118-121``) with a fake radio and committing its
    outputs as ``tests/golden/stream_frames.npz`` (``tests/golden/make_golden.py``).
  * Welch / Hann path -- the reference calls matplotlib's ``plt.psd``
    (``/root/reference/scripts/process_sigmf_data.py:188``); matplotlib 3.10.6 is a
    third-party dependency that is absent here, so its published ``mlab._spectral_helper``
    algorithm is restated below and anchored on ``scipy.signal.welch`` (installed) in
    ``tests/test_oracle.py``.  Parity for that leg is "restated + scipy-anchored".
  * overlap > 0, Blackman, max-hold, uint8 quantisation, frame stats and the 2-D histogram
    do not exist in the reference; their definitions are the ones in
    SURVEY.md section 8(a) rows A2/A3/A7-A10 and are stated here in numpy terms.

Every FFT is evaluated in complex128 (numpy keeps complex64 input in single precision,
so inputs are upcast first -- SURVEY.md section 8(c)).
"""
from __future__ import annotations

import numpy as np

WINDOW_RECT, WINDOW_HANN, WINDOW_BLACKMAN = 0, 1, 2
FMT_CF32, FMT_CI16 = 0, 1

_WINDOW_NAMES = {"rect": 0, "boxcar": 0, "none": 0, "hann": 1, "hanning": 1, "blackman": 2}


def window_id(kind) -> int:
    if isinstance(kind, str):
        return _WINDOW_NAMES[kind.lower()]
    return int(kind)


# ----------------------------------------------------------------------------- A1
def unpack_ci16(raw, scale: float = 1.0) -> np.ndarray:
    """int16 interleaved I,Q -> complex128.

    Stream path: pyadi-iio ``rx()`` result consumed at reference app/sdr/streamer.py:114
    (raw integer values, ``scale=1``).  SigMF ``ci16_le`` read at
    scripts/process_sigmf_data.py:52 autoscales by 2**-15.
    """
    raw = np.asarray(raw, dtype=np.int16).reshape(-1)
    n = raw.size // 2
    iq = raw[: 2 * n].astype(np.float64).reshape(n, 2)
    return (iq[:, 0] + 1j * iq[:, 1]) * float(scale)


def as_complex128(x, in_fmt: int = FMT_CF32, scale: float = 1.0) -> np.ndarray:
    if in_fmt == FMT_CI16:
        return unpack_ci16(x, scale)
    x = np.asarray(x)
    if x.dtype == np.float32:  # interleaved float pairs
        x = x.reshape(-1, 2)
        return (x[:, 0].astype(np.float64) + 1j * x[:, 1].astype(np.float64)) * float(scale)
    return x.astype(np.complex128) * float(scale)


# ----------------------------------------------------------------------------- A2
def frame_count(n_samples: int, nfft: int, hop: int) -> int:
    """F = (L - N)//hop + 1, tail dropped; 0 when L < N (mlab ``_spectral_helper`` /
    ``scipy.signal.spectrogram(boundary=None)`` framing behind process_sigmf_data.py:188).
    The stream path is the special case hop == N == len(rx buffer) (streamer.py:114-119)."""
    if n_samples < nfft:
        return 0
    return (n_samples - nfft) // hop + 1


def frames(x: np.ndarray, nfft: int, hop: int) -> np.ndarray:
    """Frame f is x[f*hop : f*hop + nfft]."""
    f = frame_count(len(x), nfft, hop)
    if f == 0:
        return np.empty((0, nfft), dtype=x.dtype)
    view = np.lib.stride_tricks.sliding_window_view(x, nfft)[::hop]
    return view[:f]


# ----------------------------------------------------------------------------- A3
def window(kind, nfft: int) -> np.ndarray:
    """Symmetric windows: ``np.hanning`` is mlab's default ``window_hanning``
    (process_sigmf_data.py:188); rect is the stream path (streamer.py:119)."""
    k = window_id(kind)
    if k == WINDOW_RECT:
        return np.ones(nfft, dtype=np.float64)
    if k == WINDOW_HANN:
        return np.hanning(nfft).astype(np.float64)
    if k == WINDOW_BLACKMAN:
        return np.blackman(nfft).astype(np.float64)
    raise ValueError(f"unknown window {kind!r}")


# ----------------------------------------------------------------------------- A4/A5
def stft(x: np.ndarray, nfft: int, hop: int, kind=WINDOW_RECT, chunk: int = 2048) -> np.ndarray:
    """Unnormalised forward DFT of every windowed frame, complex128, NOT shifted
    (``np.fft.fft`` at streamer.py:119)."""
    x = np.asarray(x, dtype=np.complex128)
    fr = frames(x, nfft, hop)
    w = window(kind, nfft)
    out = np.empty((fr.shape[0], nfft), dtype=np.complex128)
    for i in range(0, fr.shape[0], chunk):
        out[i : i + chunk] = np.fft.fft(fr[i : i + chunk] * w, axis=1)
    return out


def shift_bins(a: np.ndarray) -> np.ndarray:
    """fftshift along the last axis (streamer.py:119-120)."""
    return np.fft.fftshift(a, axes=-1)


def freq_axis(nfft: int, sample_rate: float, center_freq: float = 0.0) -> np.ndarray:
    """``fftshift(fftfreq(N, 1/fs)) + fc`` exactly as streamer.py:120 evaluates it."""
    return np.fft.fftshift(np.fft.fftfreq(nfft, 1 / sample_rate)) + center_freq


# ----------------------------------------------------------------------------- A6
def amplitude_db(spec: np.ndarray, eps: float = 1e-12) -> np.ndarray:
    """``20*log10(|X| + eps)`` (streamer.py:121; eps 1e-10 in scripts/sdr_realtime_dash.py:73)."""
    with np.errstate(divide="ignore"):
        return 20 * np.log10(np.abs(spec) + eps)


def stream_frame(samples, sample_rate: float, center_freq: float, eps: float = 1e-12):
    """One rx buffer through the reference's three lines (streamer.py:119-121).
    Returns (freqs, power_db) both float64[N] in fftshift order."""
    s = np.asarray(samples).astype(np.complex128)
    spec = np.fft.fftshift(np.fft.fft(s))
    return freq_axis(len(s), sample_rate, center_freq), amplitude_db(spec, eps)


def stft_db_rows(x, nfft, hop, kind=WINDOW_RECT, eps: float = 1e-12) -> np.ndarray:
    """Waterfall rows in dB, fftshift order, float64 [F][N] (what callbacks.py:176 appends)."""
    return amplitude_db(shift_bins(stft(x, nfft, hop, kind)), eps)


def stft_power_rows(x, nfft, hop, kind=WINDOW_RECT) -> np.ndarray:
    """|X|^2 rows, fftshift order, float64 [F][N]."""
    s = shift_bins(stft(x, nfft, hop, kind))
    return s.real**2 + s.imag**2


# ----------------------------------------------------------------------------- A7
def waterfall_u8(db_rows: np.ndarray, vmin: float, vmax: float) -> np.ndarray:
    """uint8 colormap index: clip(floor((db - vmin) * 256/(vmax - vmin)), 0, 255)
    (matplotlib Normalize+LUT rule; the reference leaves colouring to
    Plotly.js, callbacks.py:182-190)."""
    q = np.floor((np.asarray(db_rows, dtype=np.float64) - vmin) * (256.0 / (vmax - vmin)))
    q = np.nan_to_num(q, nan=0.0, posinf=255.0, neginf=0.0)
    return np.clip(q, 0, 255).astype(np.uint8)


def waterfall_prefloor(db_rows: np.ndarray, vmin: float, vmax: float) -> np.ndarray:
    """Pre-floor value; tests exclude bins whose pre-floor value is within 1e-3 of an integer."""
    return (np.asarray(db_rows, dtype=np.float64) - vmin) * (256.0 / (vmax - vmin))


_VIRIDIS_STOPS = ["#440154", "#482878", "#3e4989", "#31688e", "#26828e",
                  "#1f9e89", "#35b779", "#6ece58", "#b5de2b", "#fde725"]


def viridis_lut() -> np.ndarray:
    """256x3 uint8 LUT: Plotly 'Viridis' 10 stops, linear RGB interpolation
    (colorscale named at callbacks.py:187)."""
    stops = np.array([[int(h[i : i + 2], 16) for i in (1, 3, 5)] for h in _VIRIDIS_STOPS], dtype=np.float64)
    pos = np.linspace(0.0, 1.0, len(stops))
    t = (np.arange(256) + 0.5) / 256.0
    lut = np.stack([np.interp(t, pos, stops[:, c]) for c in range(3)], axis=1)
    return np.floor(lut + 0.5).astype(np.uint8)


# ----------------------------------------------------------------------------- A8
def welch_sum(x, nfft, hop, kind=WINDOW_HANN):
    """(sum_f |X_f[k]|^2 in fftshift order, F)."""
    p = stft_power_rows(x, nfft, hop, kind)
    return p.sum(axis=0), p.shape[0]


def welch_psd(x, nfft, hop, kind=WINDOW_HANN, sample_rate: float = 1.0, center_freq: float = 0.0):
    """mlab.psd semantics (behind plt.psd at process_sigmf_data.py:188): two-sided density,
    ``mean_f |X_f|^2 / (Fs * sum(w^2))``, rolled so DC is at N/2.  Returns (freqs, Pxx).
    If L < N the data is zero-padded to N (mlab ``_spectral_helper`` behaviour)."""
    x = np.asarray(x, dtype=np.complex128)
    if len(x) < nfft:
        x = np.concatenate([x, np.zeros(nfft - len(x), dtype=np.complex128)])
    acc, f = welch_sum(x, nfft, hop, kind)
    w = window(kind, nfft)
    pxx = acc / f / (sample_rate * np.sum(w**2))
    return freq_axis(nfft, sample_rate, center_freq), pxx


def power_db10(p: np.ndarray) -> np.ndarray:
    with np.errstate(divide="ignore"):
        return 10 * np.log10(p)


def maxhold(x, nfft, hop, kind=WINDOW_HANN) -> np.ndarray:
    """Per-bin max over frames of |X|^2, fftshift order (new)."""
    p = stft_power_rows(x, nfft, hop, kind)
    return p.max(axis=0) if p.shape[0] else np.zeros(nfft)


# ----------------------------------------------------------------------------- A9
def frame_stats(x, nfft: int, hop: int):
    """Per frame mean(I^2+Q^2) and max(I^2+Q^2), float64 [F] each (SURVEY.md 8(a) A9;
    the mean-power idea is scripts/pyad-iio-test.py:93)."""
    x = np.asarray(x, dtype=np.complex128)
    fr = frames(x, nfft, hop)
    p = fr.real**2 + fr.imag**2
    if p.shape[0] == 0:
        return np.zeros(0), np.zeros(0)
    return p.mean(axis=1), p.max(axis=1)


# ----------------------------------------------------------------------------- A10
def iq_hist2d(x, r: float, bins: int = 256) -> np.ndarray:
    """``np.histogram2d(I, Q, bins, range=[[-R,R],[-R,R]])[0]`` as uint32 [bins][bins],
    H[i][j] with i<->I, j<->Q; right edge inclusive, outside dropped (SURVEY.md 8(a) A10;
    replaces the 2000-point scatter at callbacks.py:199-214)."""
    x = np.asarray(x, dtype=np.complex128)
    h, _, _ = np.histogram2d(x.real, x.imag, bins=bins, range=[[-r, r], [-r, r]])
    return h.astype(np.uint32)


# ----------------------------------------------------------------------------- synthetic input
def synth_iq(n: int, seed: int, snr_db: float = 20.0, tone_cycles_per_sample: float = 0.2,
             sps: int = 8, tone_amp: float = 0.5) -> np.ndarray:
    """SURVEY.md 8(d) synthetic input: QPSK (rectangular pulses, sps samples/symbol,
    amplitude 1) + CW tone + complex AWGN at the given SNR.  complex128."""
    rng = np.random.default_rng(seed)
    nsym = (n + sps - 1) // sps
    bits = rng.integers(0, 2, size=(nsym, 2))
    sym = ((2 * bits[:, 0] - 1) + 1j * (2 * bits[:, 1] - 1)) / np.sqrt(2.0)
    qpsk = np.repeat(sym, sps)[:n]
    t = np.arange(n, dtype=np.float64)
    tone = tone_amp * np.exp(2j * np.pi * tone_cycles_per_sample * t)
    sigma2 = 10.0 ** (-snr_db / 10.0)
    noise = np.sqrt(sigma2 / 2.0) * (rng.standard_normal(n) + 1j * rng.standard_normal(n))
    return qpsk + tone + noise


def to_ci16(x: np.ndarray, gain: float = 1024.0, clip: int = 2047) -> np.ndarray:
    """12-bit-range interleaved int16 (Pluto style): round(x*gain) clipped to +-clip."""
    iq = np.empty(2 * len(x), dtype=np.int16)
    iq[0::2] = np.clip(np.rint(x.real * gain), -clip, clip).astype(np.int16)
    iq[1::2] = np.clip(np.rint(x.imag * gain), -clip, clip).astype(np.int16)
    return iq
