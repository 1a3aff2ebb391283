"""SigMF recordings in and out of the GPU spectral path.

The reference reads a recording with sigmf-python (``fromfile(base).read_samples()``,
/root/reference/scripts/process_sigmf_data.py:49-52), looks at the first 10 000 samples only
(:148,154) and hands them to ``plt.psd`` (:188); the dashboard writes ``cf32_le`` recordings
(/root/reference/app/dashboard/callbacks.py:285-311).  This module keeps those file conventions --

  * ``<base>.sigmf-meta``: JSON with ``global["core:datatype"]`` in {``cf32_le``, ``ci16_le``},
    ``global["core:sample_rate"]``, ``captures[0]["core:frequency"]``;
  * ``<base>.sigmf-data``: raw little-endian interleaved I,Q;
  * ``read_samples`` semantics of sigmf-python 1.2.x: ``ci16_le`` samples are scaled by 2**-15 and
    returned as complex64, ``cf32_le`` as is --

but memory-maps the data file and streams it through the fused STFT kernel in frame-aligned chunks
with (N - hop)-sample halos, so a capture of any length is processed without loading it whole and
without the 10 000-sample cap.  ``ci16_le`` files go to the GPU as int16 (4 B/sample over PCIe) and
are scaled by 2**-15 inside the kernel's window multiply.  Compute is libspx only (no CPU fallback).
"""
from __future__ import annotations

import io
import json
import os
import zipfile
from dataclasses import dataclass, field
from datetime import datetime, timezone
from typing import Iterator, Optional, Tuple

import numpy as np

from . import spectral as sp
from . import timedomain as td
from ._native import FMT_CF32, FMT_CI16

_DATATYPES = {"cf32_le": (np.dtype("<f4"), FMT_CF32, 1.0, 8), "ci16_le": (np.dtype("<i2"), FMT_CI16, 2.0 ** -15, 4)}
_EXTS = (".sigmf-data", ".sigmf-meta")


def resolve_base(path) -> str:
    """Base name of a recording given its data file, its meta file or the base itself
    (process_sigmf_data.py:35-44)."""
    p = os.fspath(path)
    for ext in _EXTS:
        if p.endswith(ext):
            return p[: -len(ext)]
    return p


@dataclass
class SigMFRecording:
    """A memory-mapped SigMF recording: metadata accessors + zero-copy sample access."""
    base: str
    meta: dict
    datatype: str
    raw: np.ndarray = field(repr=False)   # memmap: float32 [2L] (cf32_le) or int16 [2L] (ci16_le)

    # ---- metadata, named after the sigmf-python accessors the reference calls (:88-110,129-137)
    def get_global_info(self) -> dict:
        return dict(self.meta.get("global", {}))

    def get_global_field(self, key, default=None):
        return self.meta.get("global", {}).get(key, default)

    def get_captures(self) -> list:
        return list(self.meta.get("captures", []))

    def get_annotations(self) -> list:
        return list(self.meta.get("annotations", []))

    @property
    def sample_rate(self) -> float:
        return float(self.get_global_field("core:sample_rate", 1.0))

    @property
    def center_freq(self) -> float:
        caps = self.get_captures()
        return float(caps[0].get("core:frequency", 0.0)) if caps else 0.0

    @property
    def n_samples(self) -> int:
        return int(self.raw.shape[0] // 2)

    def __len__(self) -> int:
        return self.n_samples

    @property
    def in_fmt(self) -> int:
        return _DATATYPES[self.datatype][1]

    @property
    def in_scale(self) -> float:
        return _DATATYPES[self.datatype][2]

    @property
    def bytes_per_sample(self) -> int:
        return _DATATYPES[self.datatype][3]

    def frequency_info(self) -> dict:
        """get_frequency_info of the reference script (:125-145)."""
        info = {}
        fc = self.center_freq if self.get_captures() and "core:frequency" in self.get_captures()[0] else None
        if fc:
            info["center_frequency"] = fc
        fs = self.get_global_field("core:sample_rate")
        if fs:
            info["sample_rate"] = fs
            info["bandwidth"] = fs
            if fc:
                info["start_frequency"] = fc - fs / 2
                info["end_frequency"] = fc + fs / 2
        return info

    # ---- samples
    def raw_slice(self, start: int = 0, count: Optional[int] = None) -> np.ndarray:
        """Raw interleaved view (no copy, no scaling) of samples [start, start+count)."""
        stop = self.n_samples if count is None else min(self.n_samples, start + count)
        return self.raw[2 * start: 2 * stop]

    def read_samples(self, start: int = 0, count: Optional[int] = None) -> np.ndarray:
        """complex64, as sigmf-python's read_samples returns it (ci16_le scaled by 2**-15)."""
        r = self.raw_slice(start, count)
        if self.datatype == "cf32_le":
            return np.ascontiguousarray(r, dtype=np.float32).view(np.complex64)
        return (r.astype(np.float32) * np.float32(2.0 ** -15)).view(np.complex64)

    def frame_chunks(self, nfft: int, hop: int, max_chunk_samples: int = 1 << 26) -> Iterator[Tuple[int, int, np.ndarray]]:
        """Frame-aligned pieces of the capture: yields (first_frame, n_frames, raw view).  Consecutive
        pieces overlap by the (nfft - hop)-sample halo, so framing is identical to one pass over the file."""
        L = self.n_samples
        F = (L - nfft) // hop + 1 if L >= nfft else 0
        per = max(1, (max_chunk_samples - nfft) // hop + 1)
        f0 = 0
        while f0 < F:
            nf = min(per, F - f0)
            yield f0, nf, self.raw_slice(f0 * hop, (nf - 1) * hop + nfft)
            f0 += nf


def fromfile(path) -> SigMFRecording:
    """Open ``<base>.sigmf-meta`` + ``<base>.sigmf-data`` (same call name as sigmf-python's, :49)."""
    base = resolve_base(path)
    with open(base + ".sigmf-meta", "r") as fh:
        meta = json.load(fh)
    datatype = meta.get("global", {}).get("core:datatype")
    if datatype not in _DATATYPES:
        raise ValueError(f"unsupported core:datatype {datatype!r} (supported: {sorted(_DATATYPES)})")
    dt = _DATATYPES[datatype][0]
    data_path = base + ".sigmf-data"
    n = os.path.getsize(data_path) // dt.itemsize
    n -= n % 2
    raw = np.memmap(data_path, dtype=dt, mode="r", shape=(n,)) if n else np.zeros(0, dt)
    return SigMFRecording(base, meta, datatype, raw)


def build_metadata(sample_rate, center_freq, datatype: str = "cf32_le", hw: str = "", description: Optional[str] = None,
                   when: Optional[datetime] = None) -> dict:
    """The metadata dictionary the dashboard's recorder emits (callbacks.py:285-304)."""
    when = when or datetime.now(timezone.utc)
    return {
        "global": {
            "core:datatype": datatype,
            "core:sample_rate": int(sample_rate),
            "core:version": "1.0.0",
            "core:description": description or "SDR live stream sample from IQ Visualizer",
            "core:author": "SDR IQ Visualizer Dashboard",
            "core:recorder": "PlutoSDR via pyadi-iio",
            "core:hw": hw or "PlutoSDR",
            "core:license": "CC0-1.0",
        },
        "captures": [{"core:sample_start": 0, "core:frequency": int(center_freq),
                      "core:datetime": when.replace(tzinfo=None).isoformat() + "Z"}],
        "annotations": [],
    }


def _data_bytes(samples, datatype: str) -> bytes:
    a = np.asarray(samples)
    if datatype == "cf32_le":
        return np.ascontiguousarray(a, dtype=np.complex64).tobytes()      # callbacks.py:307-310
    if a.dtype == np.int16:
        return np.ascontiguousarray(a).astype("<i2", copy=False).tobytes()
    iq = np.empty(2 * a.size, dtype="<i2")
    iq[0::2] = np.clip(np.rint(a.real * 32768.0), -32768, 32767)
    iq[1::2] = np.clip(np.rint(a.imag * 32768.0), -32768, 32767)
    return iq.tobytes()


def write_recording(base, samples, sample_rate, center_freq, datatype: str = "cf32_le", **meta_kw) -> str:
    """Write ``<base>.sigmf-data`` / ``.sigmf-meta``.  ``samples``: complex array (any datatype) or
    interleaved int16 (``ci16_le``)."""
    if datatype not in _DATATYPES:
        raise ValueError(f"unsupported datatype {datatype!r}")
    base = resolve_base(base)
    with open(base + ".sigmf-data", "wb") as fh:
        fh.write(_data_bytes(samples, datatype))
    with open(base + ".sigmf-meta", "w") as fh:
        json.dump(build_metadata(sample_rate, center_freq, datatype, **meta_kw), fh, indent=2)
    return base


def recording_zip(samples, sample_rate, center_freq, base_filename: str, hw: str = "") -> bytes:
    """The zip the dashboard's "save" button downloads (callbacks.py:313-345): data + meta + README."""
    x = np.asarray(samples)
    meta = build_metadata(sample_rate, center_freq, "cf32_le", hw=hw)
    buf = io.BytesIO()
    with zipfile.ZipFile(buf, "w", zipfile.ZIP_DEFLATED) as z:
        z.writestr(f"{base_filename}.sigmf-data", _data_bytes(x, "cf32_le"))
        z.writestr(f"{base_filename}.sigmf-meta", json.dumps(meta, indent=2))
        z.writestr("README.txt",
                   "SigMF Recording from SDR IQ Visualizer\n"
                   f"Sample Rate: {sample_rate / 1e6:.2f} MHz\nCenter Frequency: {center_freq / 1e6:.2f} MHz\n"
                   f"Number of Samples: {x.size}\nDuration: {x.size / sample_rate:.3f} seconds\n"
                   f"Files: {base_filename}.sigmf-data (complex float32), {base_filename}.sigmf-meta (JSON)\n")
    return buf.getvalue()


@dataclass
class RecordingSpectra:
    freqs: np.ndarray                 # float64 [N], fftshift order, + center frequency
    pxx: np.ndarray                   # float64 [N] two-sided density (mlab.psd); plot 10*log10
    n_frames: int
    maxhold: Optional[np.ndarray] = None     # float32 [N] max_f |X|^2
    wf_rows: Optional[np.ndarray] = None     # uint8 [F, N]
    hist: Optional[np.ndarray] = None        # uint32 [bins, bins]
    sample_rate: float = 1.0
    center_freq: float = 0.0
    h2d_bytes: int = 0
    d2h_bytes: int = 0


def process_recording(path_or_rec, nfft: int = 1024, hop: Optional[int] = None, window="hann", waterfall: bool = False,
                      vmin: float = -120.0, vmax: float = 0.0, maxhold: bool = True, hist_r: Optional[float] = None,
                      hist_bins: int = 256, max_samples: Optional[int] = None, max_chunk_samples: int = 1 << 26,
                      device: int = 0) -> RecordingSpectra:
    """The offline path of process_sigmf_data.py on the GPU, for the whole file: Welch PSD with
    ``plt.psd(data, NFFT, Fs, Fc)`` semantics (Hann, ``noverlap = nfft - hop``, default no overlap), plus
    optional max-hold, uint8 waterfall rows and the I/Q density histogram.  ``max_samples`` restores
    the reference's truncation (10 000 there) when wanted."""
    rec = path_or_rec if isinstance(path_or_rec, SigMFRecording) else fromfile(path_or_rec)
    hop = int(hop or nfft)
    if max_samples is not None and max_samples < rec.n_samples:
        rec = SigMFRecording(rec.base, rec.meta, rec.datatype, rec.raw[: 2 * max_samples])
    fs, fc = rec.sample_rate, rec.center_freq
    freqs = sp.freq_axis(nfft, fs, fc)
    L = rec.n_samples
    if L < nfft:   # mlab zero-pads a short input to one frame
        f, pxx = sp.welch_psd(rec.raw_slice() if rec.in_fmt == FMT_CI16 else rec.read_samples(), nfft, hop, window, fs, fc,
                              in_fmt=rec.in_fmt, in_scale=rec.in_scale, device=device)
        return RecordingSpectra(f, pxx, 1, sample_rate=fs, center_freq=fc)
    pl = sp.get_plan(nfft, hop, window, rec.in_fmt, rec.in_scale, sp.DB_EPS_REFERENCE, device)
    F = (L - nfft) // hop + 1
    welch = np.zeros((1, nfft), np.float64)
    mh = np.zeros((1, nfft), np.float32) if maxhold else False
    rows = np.empty((F, nfft), np.uint8) if waterfall else None
    hist = np.zeros((hist_bins, hist_bins), np.uint32) if hist_r is not None else None
    h2d = d2h = 0
    first = True
    for f0, nf, raw in rec.frame_chunks(nfft, hop, max_chunk_samples):
        x = raw if rec.in_fmt == FMT_CI16 else raw.view(np.complex64)
        r = pl.stft(x, welch=welch, maxhold=mh, wf_rows=(rows[f0:f0 + nf] if waterfall else False), vmin=vmin, vmax=vmax,
                    accumulate=not first)
        h2d += r.h2d_bytes
        d2h += r.d2h_bytes
        first = False
    if hist is not None:   # every sample once (no halo): the histogram is over the capture, not over frames
        step = max_chunk_samples
        for s0 in range(0, L, step):
            raw = rec.raw_slice(s0, step)
            x = raw if rec.in_fmt == FMT_CI16 else raw.view(np.complex64)
            td.iq_hist2d(x, hist_r, hist_bins, in_fmt=rec.in_fmt, in_scale=rec.in_scale, out=hist, accumulate=True, device=device)
    pxx, _ = pl.welch_finalize(welch[0], F, fs, want_db=False)
    return RecordingSpectra(freqs, pxx, F, mh[0] if maxhold else None, rows, hist, fs, fc, h2d, d2h)


def psd(path_or_rec, NFFT: int = 1024, noverlap: int = 0, max_samples: Optional[int] = None, device: int = 0):
    """``plt.psd(data, NFFT=1024, Fs=sample_rate, Fc=center_freq)`` of a recording (:188): returns
    ``(Pxx, freqs)`` in matplotlib's order."""
    out = process_recording(path_or_rec, NFFT, NFFT - noverlap, "hann", maxhold=False, max_samples=max_samples, device=device)
    return out.pxx, out.freqs
