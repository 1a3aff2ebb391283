#!/usr/bin/env python3
"""Generate the golden fixtures in this directory by RUNNING THE REFERENCE ITSELF.

Run in the build container only (needs /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py

Outputs (committed):
  stream_frames.npz       inputs + outputs of the reference's SDRDataStreamer._stream_data
                          (/root/reference/app/sdr/streamer.py:95-133) driven by a fake radio.
  classifier_cases.npz    spectra fed to the reference classifier
  classifier_cases.json   what /root/reference/app/processing/classifier.py returned for them
                          (public results + private-helper intermediates), history cleared
                          before every case.
The reference has no golden vectors of its own for this path; these
fixtures are how the oracle and the CUDA path are pinned to it.
"""
import json
import os
import sys
from unittest.mock import MagicMock

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("SPX_REFERENCE", "/root/reference")


def import_reference():
    for k in [k for k in sys.modules if k == "app" or k.startswith("app.")]:
        del sys.modules[k]
    sys.path.insert(0, REF)
    sys.modules.setdefault("adi", MagicMock())
    from app.processing import classifier  # noqa
    from app.sdr import streamer  # noqa
    sys.path.pop(0)
    return classifier, streamer


class OneShotRadio:
    """rx() hands out the prepared buffers, then stops the streamer."""

    def __init__(self, owner, buffers):
        self.owner, self.buffers, self.i = owner, buffers, 0

    def rx(self):
        buf = self.buffers[self.i]
        self.i += 1
        if self.i >= len(self.buffers):
            self.owner.running = False
        return buf


def run_stream(streamer_mod, buffers, fs, fc):
    s = streamer_mod.SDRDataStreamer(sample_rate=fs, center_freq=fc, rx_buffer_size=len(buffers[0]))
    import queue
    s.data_queue = queue.Queue(maxsize=len(buffers) + 1)
    s.sdr = OneShotRadio(s, buffers)
    s.connected = True
    s.running = True
    s._stream_data()
    out = []
    while True:
        d = s.get_latest_data()
        if d is None:
            break
        out.append(d)
    assert len(out) == len(buffers)
    return out


def pluto_like(rng, n, tone_bin, amp=600.0, noise=20.0):
    """complex128 with raw integer values, like pyadi-iio rx() (unscaled int16)."""
    t = np.arange(n)
    x = amp * np.exp(2j * np.pi * tone_bin * t / n) + noise * (rng.standard_normal(n) + 1j * rng.standard_normal(n))
    return (np.clip(np.rint(x.real), -2047, 2047) + 1j * np.clip(np.rint(x.imag), -2047, 2047)).astype(np.complex128)


def make_stream(streamer_mod):
    rng = np.random.default_rng(20261018)
    blob = {}
    meta = []
    cases = [(64, 1e6, 2.4e9, 5.0), (1024, 1e6, 2.4e9, 100.37), (4096, 61.44e6, 2.4e9, 1500.0), (4096, 1e6, 915e6, 77.77)]
    for ci, (n, fs, fc, tb) in enumerate(cases):
        bufs = [pluto_like(rng, n, tb + k) for k in range(2)]
        if ci == 0:
            bufs.append(np.zeros(n, dtype=np.complex128))  # all-zero buffer: exercises the 1e-12 eps
        outs = run_stream(streamer_mod, bufs, int(fs), int(fc))
        for k, (b, o) in enumerate(zip(bufs, outs)):
            key = f"c{ci}_{k}"
            assert o["samples"] is b
            blob[key + "_samples"] = b
            blob[key + "_freqs"] = np.asarray(o["freqs"], dtype=np.float64)
            blob[key + "_power_db"] = np.asarray(o["power_db"], dtype=np.float64)
            meta.append({"key": key, "n": n, "sample_rate": int(fs), "center_freq": int(fc)})
    blob["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    np.savez_compressed(os.path.join(HERE, "stream_frames.npz"), **blob)
    print("stream_frames.npz:", len(meta), "frames")


def classifier_inputs():
    cases = {}
    # KA-1: CW-like spectrum of the reference's own unit test, seeded
    np.random.seed(0)
    f = np.linspace(100e6, 101e6, 1024)
    p = np.random.normal(-80, 1, 1024)
    p[512] = -20
    p[511] = p[513] = -30
    cases["ka1_cw"] = (f, p)
    # KA-2: deterministic, no RNG
    n = 4096
    f = np.fft.fftshift(np.fft.fftfreq(n, 1 / 61.44e6)) + 2.4e9
    k = np.arange(n)
    p = -90 + 0.002 * k + 40 * np.clip(1 - np.abs(k - 1500) / 200, 0, None) + 25 * np.clip(1 - np.abs(k - 3000) / 3, 0, None)
    cases["ka2_wide"] = (f, p)
    rng = np.random.default_rng(7)
    # the reference's simple-classifier unit tests (tests/test_classifier.py:13-35)
    p = np.zeros(100); p[50] = 50
    cases["t_narrow100"] = (np.linspace(0, 10e6, 100), p)
    cases["t_wide200"] = (np.linspace(0, 20e6, 200), np.full(200, 50.0))
    # pure noise, low SNR
    cases["noise300"] = (np.linspace(433e6, 434e6, 300), rng.normal(-95, 0.4, 300))
    # multitone
    n = 2048
    p = rng.normal(-85, 1.0, n)
    for c in (700, 900, 1100):
        p[c - 1 : c + 2] = [-42, -30, -41]
    cases["multitone2048"] = (np.linspace(144e6, 146e6, n), p)
    # OFDM-like wide plateau with ripple
    n = 4096
    k = np.arange(n)
    p = rng.normal(-90, 0.8, n)
    band = (k > 600) & (k < 3500)
    p[band] = -50 + 3 * np.cos(2 * np.pi * k[band] / 16.0) + rng.normal(0, 0.3, band.sum())
    cases["ofdm4096"] = (np.fft.fftshift(np.fft.fftfreq(n, 1 / 20e6)) + 2.44e9, p)
    # FM broadcast candidate
    n = 1000
    k = np.arange(n)
    p = rng.normal(-100, 1.0, n) + 55 * np.exp(-0.5 * ((k - 500) / 40.0) ** 2)
    cases["fm1000"] = (np.linspace(97.0e6, 98.0e6, n), p)
    # flat (sigma < 1e-9 branch, ties everywhere)
    cases["flat64"] = (np.linspace(0, 1e6, 64), np.full(64, -70.0))
    # tiny inputs
    cases["n1"] = (np.array([1.0e6]), np.array([-3.0]))
    cases["n2"] = (np.array([1.0e6, 2.0e6]), np.array([-3.0, -40.0]))
    cases["n3"] = (np.array([1.0e6, 2.0e6, 3.0e6]), np.array([-50.0, -10.0, -52.0]))
    # plateau peaks / duplicates (strict-inequality behaviour) and close peaks (greedy thinning)
    p = rng.normal(-80, 0.5, 900)
    p[100:103] = -20.0
    p[300] = -25; p[302] = -24; p[304] = -23; p[306] = -26
    cases["plateau900"] = (np.linspace(0, 9e6, 900), p)
    # large n, many peaks
    n = 16384
    p = rng.normal(-75, 2.0, n)
    p[::257] += 30
    cases["comb16384"] = (np.linspace(5.0e9, 5.1e9, n), p)
    return cases


def make_classifier(cls):
    cases = classifier_inputs()
    blob, results = {}, {}
    for name, (f, p) in cases.items():
        f = np.asarray(f, dtype=np.float64)
        p = np.asarray(p, dtype=np.float64)
        blob[name + "_freqs"], blob[name + "_power_db"] = f, p
        cls._CLASS_HISTORY.clear(); cls._CONF_HISTORY.clear()
        adv = cls.classify_signal_advanced(f, p)
        cls._CLASS_HISTORY.clear(); cls._CONF_HISTORY.clear()
        simple = cls.classify_signal_simple(f, p)
        nf = cls._estimate_noise_floor(p)
        snr = float(np.max(p) - nf)
        thr = max(nf + 5.0, np.max(p) - 0.9 * snr + 5.0)
        peaks = cls._find_peaks(p, threshold_db=thr, min_distance_bins=max(3, len(p) // 300))
        results[name] = {
            "advanced": adv,
            "simple": simple,
            "noise_floor_db": nf,
            "adaptive_thr": float(thr),
            "peaks": [int(i) for i in peaks],
            "flatness": cls._spectral_flatness(p),
            "kurtosis": cls._spectral_kurtosis(p),
            "bw": [cls._occupied_bandwidth(f, p, d) for d in (3, 10, 20)],
            "peak_spacing_std_hz": cls._peak_spacing_std(f, peaks),
        }
    # temporal smoothing sequence (module-global history, classifier.py:125-139):
    seq = ["ka1_cw"] * 3 + ["noise300"] * 2 + ["ka1_cw"] + ["multitone2048"] * 4 + ["ka2_wide"] * 8
    cls._CLASS_HISTORY.clear(); cls._CONF_HISTORY.clear()
    hist = []
    for name in seq:
        f, p = cases[name]
        r = cls.classify_signal_advanced(np.asarray(f, float), np.asarray(p, float))
        hist.append({"case": name, "label": r["label"], "confidence": r["confidence"], "reasons": r["reasons"],
                     "explanation": r["explanation"]})
    cls._CLASS_HISTORY.clear(); cls._CONF_HISTORY.clear()
    np.savez_compressed(os.path.join(HERE, "classifier_cases.npz"), **blob)
    with open(os.path.join(HERE, "classifier_cases.json"), "w") as fh:
        json.dump({"cases": results, "sequence": hist,
                   "empty_advanced": cls.classify_signal_advanced(np.array([]), np.array([])),
                   "empty_simple": cls.classify_signal_simple(np.array([]), np.array([]))}, fh, indent=1)
    print("classifier_cases:", len(results), "cases,", len(hist), "sequence steps")


if __name__ == "__main__":
    classifier, streamer = import_reference()
    make_stream(streamer)
    make_classifier(classifier)
