#!/usr/bin/env python3
"""Single-GPU measurements of the BASELINE configs that are not the bench line: C1 (SigMF file ->
PSD + waterfall), C3 (16 Mi samples -> frame stats + 256x256 I/Q histogram), plus the classifier-feature kernel on a
batch of spectra.  Device-resident numbers use CUDA events on the launching stream; end-to-end numbers are host
buffers through the public API.  One JSON line per case."""
import json
import os
import sys
import tempfile
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sdr_iq_visualizer_b200 import _native as nat, features, sigmf_io, spectral as sp, synth, timedomain as td  # noqa: E402

HBM = 6539.9
try:
    HBM = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass


def timed(fn, stream=0, warmup=3, iters=10):
    """Median / best time of fn() in ms.  The time-domain and feature entry points launch on library-internal
    (non-blocking) streams, so CUDA events on another stream would not bracket them: wall clock between device
    synchronisations instead (a few microseconds of launch + sync overhead are included)."""
    for _ in range(warmup):
        fn()
    nat.device_sync(0)
    ms = []
    for _ in range(iters):
        t0 = time.perf_counter()
        fn()
        nat.device_sync(0)
        ms.append((time.perf_counter() - t0) * 1e3)
    return float(np.median(ms)), float(min(ms))


def emit(**kw):
    print(json.dumps(kw), flush=True)


def c1():
    L, n, hop = 1 << 20, 1024, 512
    x = synth.synth_iq(L, seed=1).astype(np.complex64)
    with tempfile.TemporaryDirectory() as d:
        base = sigmf_io.write_recording(os.path.join(d, "c1"), x, 1e6, 2.4e9)
        for _ in range(2):
            sigmf_io.process_recording(base, n, hop, "hann", waterfall=True, vmin=-60.0, vmax=70.0)
        ts = []
        for _ in range(5):
            t0 = time.perf_counter()
            out = sigmf_io.process_recording(base, n, hop, "hann", waterfall=True, vmin=-60.0, vmax=70.0)
            ts.append(time.perf_counter() - t0)
    pl = sp.SpectralPlan(n, hop, "hann", sp.FMT_CF32)
    dx = nat.DeviceArray.from_host(x)
    F = pl.frame_count(L)
    kw = dict(wf_rows=nat.DeviceArray((F, n), np.uint8), welch=nat.DeviceArray((1, n), np.float64), maxhold=nat.DeviceArray((1, n), np.float32))
    _, ms = pl.time_stft(dx, warmup=3, iters=10, vmin=-60.0, vmax=70.0, **kw)
    emit(case="C1 SigMF cf32 2^20 samples, 1024-pt Hann 50%: file -> PSD + u8 waterfall", e2e_ms=round(min(ts) * 1e3, 3),
         e2e_Msps=round(L / min(ts) / 1e6, 1), kernel_us=round(float(np.median(ms)) * 1e3, 2),
         kernel_GSps=round(L / (float(np.median(ms)) * 1e-3) / 1e9, 1), frames=out.n_frames,
         note="8 MiB of input: launch/latency bound on a B200; CPU/oracle config of BASELINE")


def c3():
    L = 1 << 24
    rng = np.random.default_rng(3)
    xc = (0.7 * (rng.standard_normal(L) + 1j * rng.standard_normal(L))).astype(np.complex64)
    xi = synth.to_ci16(xc.astype(np.complex128) * (1.0 / 0.7))
    for name, x, fmt, r, bps in (("cf32", xc, sp.FMT_CF32, 4.0, 8), ("ci16", xi, sp.FMT_CI16, 2048.0, 4)):
        dx = nat.DeviceArray.from_host(x)
        dh = nat.DeviceArray((256, 256), np.uint32, zero=True)
        med, best = timed(lambda: td.iq_hist2d(dx, r, 256, in_fmt=fmt, out=dh))
        emit(case=f"C3 I/Q 256x256 histogram, 2^24 {name} samples (device-resident, input L2-warm)", kernel_us=round(med * 1e3, 2),
             GSps=round(L / (med * 1e-3) / 1e9, 1), bytes_per_sample=bps)
        # cold input: rotate over copies that together are several times the 126 MB L2 (no dirty lines left behind)
        ncopy = 16
        copies = [dx] + [nat.DeviceArray.from_host(x) for _ in range(ncopy - 1)]
        state = {"i": 0}

        def cold_hist():
            state["i"] += 1
            td.iq_hist2d(copies[state["i"] % ncopy], r, 256, in_fmt=fmt, out=dh)

        outs = (nat.DeviceArray((L // 4096,), np.float32), nat.DeviceArray((L // 4096,), np.float32))

        def cold_stats():
            state["i"] += 1
            td.frame_stats(copies[state["i"] % ncopy], 4096, 4096, in_fmt=fmt, out=outs)

        # device time of 32 back-to-back calls on one stream (CUDA events on that stream), input cold
        ss = nat.SideStream(0)
        tm = nat.DeviceTimer(0, ss.handle)
        per_call = []
        for rep in range(4):
            tm.start()
            for k in range(32):
                td.iq_hist2d(copies[k % ncopy], r, 256, in_fmt=fmt, out=dh, stream=ss.handle)
            tm.stop()
            per_call.append(tm.elapsed_ms() / 32)
        emit(case=f"C3 I/Q 256x256 histogram, 2^24 {name} samples: device time per call, 32 calls back to back on one stream (zero + count + merge), input cold",
             kernel_us=round(min(per_call[1:]) * 1e3, 2), GSps=round(L / (min(per_call[1:]) * 1e-3) / 1e9, 1),
             hbm_frac=round(L * bps / (min(per_call[1:]) * 1e-3) / 1e9 / HBM, 3), bytes_per_sample=bps)
        ss.sync()
        med, best = timed(cold_hist, warmup=ncopy, iters=3 * ncopy)
        emit(case=f"C3 I/Q 256x256 histogram, 2^24 {name} samples (device-resident, input cold: {ncopy} rotating copies)",
             kernel_us=round(med * 1e3, 2), GSps=round(L / (med * 1e-3) / 1e9, 1),
             hbm_frac=round(L * bps / (med * 1e-3) / 1e9 / HBM, 3), bytes_per_sample=bps)
        med, best = timed(cold_stats, warmup=ncopy, iters=3 * ncopy)
        emit(case=f"C3 per-frame mean/peak power, 4096-sample frames, 2^24 {name} samples (device-resident, input cold)",
             kernel_us=round(med * 1e3, 2), GSps=round(L / (med * 1e-3) / 1e9, 1),
             hbm_frac=round(L * bps / (med * 1e-3) / 1e9 / HBM, 3), bytes_per_sample=bps)
        for c in copies[1:]:
            c.free()
        t0 = time.perf_counter()
        for _ in range(3):
            td.iq_hist2d(x, r, 256, in_fmt=fmt)
        dt = (time.perf_counter() - t0) / 3
        emit(case=f"C3 histogram end to end from pageable host memory ({name})", e2e_ms=round(dt * 1e3, 2), e2e_GSps=round(L / dt / 1e9, 2))
        dx.free()


def k3():
    rng = np.random.default_rng(4)
    for n, batch in ((4096, 1), (4096, 64), (4096, 1024), (65536, 64)):
        p = rng.normal(-80, 3, (batch, n))
        dp = nat.DeviceArray.from_host(p)
        t0 = time.perf_counter()
        for _ in range(5):
            features.measure_batch(dp, n=n, batch=batch, want_peaks=False)
        dt = (time.perf_counter() - t0) / 5
        emit(case=f"K3 classifier features, {batch} spectra x {n} bins (float64, device-resident, host round trip included)",
             ms=round(dt * 1e3, 3), spectra_per_s=round(batch / dt, 1))
        dp.free()


def c2ring():
    """Config 2 as a stream: one second of 61.44 MS/s int16 IQ through the pinned ring (4 slots x 2^22 samples), every
    slot one Welch block with u8 rows, max-hold and on-device classifier measurements."""
    from sdr_iq_visualizer_b200.ring import StreamRing
    L, N, HOP, SLOT = 61_440_000, 4096, 1024, 1 << 22
    x = synth.tiled_ci16(L, 2)
    pl = sp.SpectralPlan(N, HOP, "hann", sp.FMT_CI16)
    ring = StreamRing(pl, n_slots=4, slot_samples=SLOT, wf_rows=True, welch=True, maxhold=True, vmin=20.0, vmax=130.0,
                      features=True, sample_rate=61.44e6)

    def one_pass():
        pending, frames, snr = 0, 0, []
        for lo in range(0, L, SLOT):
            hi = min(L, lo + SLOT)
            buf = ring.acquire()                       # the producer (radio driver) writes straight into the pinned slot
            buf[: 2 * (hi - lo)] = x[2 * lo: 2 * hi]
            ring.commit(hi - lo)
            pending += 1
            if pending == 3:
                b = ring.collect(); frames += b["n_frames"]; snr.append(b["features"]["snr_db"]); ring.release(); pending -= 1
        while pending:
            b = ring.collect(); frames += b["n_frames"]; snr.append(b["features"]["snr_db"]); ring.release(); pending -= 1
        return frames, snr

    one_pass()
    ts = []
    for _ in range(3):
        t0 = time.perf_counter()
        frames, snr = one_pass()
        ts.append(time.perf_counter() - t0)
    st = ring.stats()
    dt = min(ts)
    emit(case="C2 as a stream: 61.44 M int16 samples through the pinned ring (4 x 2^22 slots), blocks with u8 rows + Welch + "
              "max-hold + classifier measurements; includes the producer's copy into the slots",
         s_per_stream_second=round(dt, 4), GSps=round(L / dt / 1e9, 2), real_time_margin_x=round(L / dt / 61.44e6, 1),
         frames=frames, blocks=len(snr), h2d_gbs=round(4 * L / dt / 1e9, 1), ring_totals=st)
    ring.close(); pl.close()


if __name__ == "__main__":
    print(json.dumps(nat.device_info(0)))
    which = sys.argv[1:] or ["c1", "c3", "k3"]
    for name in which:
        {"c1": c1, "c3": c3, "k3": k3, "c2ring": c2ring}[name]()
