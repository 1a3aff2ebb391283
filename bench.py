#!/usr/bin/env python3
"""bench.py -- IQ Msamples/s through windowed FFT -> PSD -> waterfall on B200 (BASELINE.json metric).

Workload (N = 1 GPU and per rank for N > 1): BASELINE config 2 -- one second of a 61.44 MS/s int16
IQ stream (61 440 000 samples), 4096-point Hann FFT with 75 % overlap (59 997 frames), producing the
uint8 waterfall rows, the Welch sum, the max-hold and the classifier features of the Welch PSD.
One "step" = one pass of that path over that batch.

  value   : whole-job Msamples/s with the input already resident in HBM (device buffers).
  e2e     : the same step through the host-buffer C-ABI call (pinned host input and outputs;
            H2D of the samples and D2H of rows/PSD inside the timed region).
  roofline: the fused STFT kernel's algorithmic bytes (8 B/sample: 4 in + 4 x 1 B rows) over its
            CUDA-event duration vs the measured HBM peak; FP32 issue figures alongside because this
            shape is instruction-bound, not HBM-bound.
  cpu_baseline: the float64 numpy oracle (a port: the reference has no windowed/overlapped path)
            on a bounded slice, single core, timed on this box.

`--impl reference` times the CPU oracle port with all host cores (multiprocessing over frame
blocks) on the same config and prints the same line with "impl": "reference".
Multi-GPU: config 2 is one ordered stream, so ranks are independent replicas (one stream per GPU,
no data-path collective); torch.distributed is used only for the barrier and the max-over-ranks.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NFFT, HOP, FS, FC = 4096, 1024, 61.44e6, 2.4e9
L_STEP = 61_440_000
VMIN, VMAX = 20.0, 130.0
BYTES_PER_SAMPLE = 4 + (NFFT // HOP) * 1          # int16 IQ in + four u8 rows touched per sample
FLOP_PER_SAMPLE = (NFFT // HOP) * (5 * 12 + 20)
METRIC = "IQ Msamples/s through windowed FFT->PSD->waterfall"


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return json.load(fh), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0}, "fallback"


def synth_ci16(n, seed):
    """QPSK + CW tone + AWGN, 12-bit range int16; a 4 Mi-sample block tiled to n."""
    from sdr_iq_visualizer_b200 import synth
    return synth.tiled_ci16(n, seed)


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region (B200_PROFILING.md clocks line):
    NVML in a thread every 5 ms; nvidia-smi --query-gpu as the fallback."""

    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.sm, self.mask, self.max_mhz = [], 0, None
        self._stop = threading.Event()
        self._thr = None
        self._nvml = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = self.idx
            if vis:
                try:
                    phys = int(vis.split(",")[self.idx])
                except Exception:
                    phys = self.idx
            h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            self._nvml = (pynvml, h)
        except Exception:
            self._nvml = None
        self._thr = threading.Thread(target=self._run, daemon=True)
        self._thr.start()

    def _sample_smi(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        out = subprocess.run(["nvidia-smi", "-i", str(self.idx), f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                             capture_output=True, text=True, timeout=10).stdout.strip().split(",")
        self.sm.append(float(out[0])); self.max_mhz = float(out[1])
        for bit, v in zip((0x8, 0x40, 0x20, 0x4), out[2:6]):
            if v.strip().lower().startswith("active"):
                self.mask |= bit

    def _run(self):
        while not self._stop.is_set():
            try:
                if self._nvml:
                    nv, h = self._nvml
                    self.sm.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                    self.mask |= int(nv.nvmlDeviceGetCurrentClocksEventReasons(h))
                else:
                    self._sample_smi()
            except Exception:
                pass
            self._stop.wait(0.005)

    def stop(self):
        self._stop.set()
        if self._thr:
            self._thr.join(timeout=15)
        reasons = sorted(name for bit, name in self.REASONS.items() if self.mask & bit)
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.max_mhz,
                "reasons": reasons, "samples": len(self.sm), "source": "nvml" if self._nvml else "nvidia-smi"}


def dist_setup(n_gpus):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    return rank, world, local, dist


def barrier_max(dist, local, seconds):
    if dist is None:
        return seconds
    import torch
    t = torch.tensor([seconds], dtype=torch.float64, device=f"cuda:{local}")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def cpu_baseline_single(sample_log2=26):
    from oracle import pipeline_ref
    x = synth_ci16(1 << sample_log2, seed=2)
    best = 1e30
    for _ in range(2):
        t0 = time.perf_counter()
        pipeline_ref.c2_step(x, NFFT, HOP, "hann", FS, FC, VMIN, VMAX)
        best = min(best, time.perf_counter() - t0)
    return {"value": round((1 << sample_log2) / best / 1e6, 3), "unit": "Msamples/s", "cores": 1, "kind": "port",
            "sample": f"2^{sample_log2} int16 IQ samples of the config-2 step (float64 numpy oracle, best of 2), numpy {np.__version__}, "
                      f"{os.cpu_count()} host cores present"}


def run_reference(args):
    """CPU arm: the oracle port on all host cores, same config/metric."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    from oracle import pipeline_ref
    cores = os.cpu_count() or 1
    n = 1 << 23
    x = synth_ci16(n, seed=2)
    ctx = mp.get_context("fork")
    with ctx.Pool(cores) as pool:
        for _ in range(args.warmup):
            pipeline_ref.c2_step_parallel(x, pool, cores, NFFT, HOP)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            acc, mx, F = pipeline_ref.c2_step_parallel(x, pool, cores, NFFT, HOP)
        dt = time.perf_counter() - t0
    v = round(n * args.steps / dt / 1e6, 3)
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "Msamples/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(dt / args.steps * 1e3, 3),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "config2: int16 IQ, 4096-pt Hann FFT, 75% overlap, u8 rows + Welch + max-hold (CPU oracle port)",
                       "nfft": NFFT, "hop": HOP, "samples_per_step": n},
            "cpu_baseline": {"value": v, "unit": "Msamples/s", "cores": cores, "kind": "port",
                             "sample": f"2^23-sample slice of the config-2 second per step, frame blocks over {cores} processes"},
            "e2e": {"value": v, "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--variant", type=int, default=int(os.environ.get("SPX_VARIANT", "-1")))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
        return

    rank, world, local, dist = dist_setup(args.gpus)
    try:  # keep this rank's pinned buffers and threads on the NUMA node its GPU hangs off
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        phys = int(vis.split(",")[local]) if vis else local
        pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(phys))
    except Exception:
        pass
    from sdr_iq_visualizer_b200 import _native as nat
    from sdr_iq_visualizer_b200 import features, spectral as sp
    nat.require_device()
    dev = local
    peaks, peak_src = measured_peaks()

    host_in = nat.pinned_empty(2 * L_STEP, np.int16)
    host_in[:] = synth_ci16(L_STEP, seed=2 + rank)
    variant = args.variant if args.variant >= 0 else sp.DEFAULT_VARIANT.get(NFFT, 0)
    pl = sp.SpectralPlan(NFFT, HOP, "hann", sp.FMT_CI16, device=dev, variant=variant)
    F = pl.frame_count(L_STEP)

    # ---------------- device-resident leg: every launch of a step goes to the plan's compute stream
    d_in = nat.DeviceArray.from_host(host_in, dev)
    d_wf = nat.DeviceArray((F, NFFT), np.uint8, dev)
    d_we = nat.DeviceArray((1, NFFT), np.float64, dev)
    d_mh = nat.DeviceArray((1, NFFT), np.float32, dev)
    d_pxx = nat.DeviceArray((NFFT,), np.float64, dev)
    d_pdb = nat.DeviceArray((NFFT,), np.float64, dev)
    st = pl.stream
    total_timer = nat.DeviceTimer(dev, st)
    kernel_timers = [nat.DeviceTimer(dev, st) for _ in range(args.steps)]

    def device_step(ktimer):
        # fused STFT launch (bracketed by its own CUDA events when timed), Welch finalize, classifier features
        if ktimer is not None:
            ktimer.start()
        res = pl.stft(d_in, wf_rows=d_wf, welch=d_we, maxhold=d_mh, vmin=VMIN, vmax=VMAX)
        if ktimer is not None:
            ktimer.stop()
        pl.welch_finalize(d_we, res.n_frames, FS, pxx=d_pxx, pdb=d_pdb)
        # classifier features of this Welch block: kernel + asynchronous copy of the result struct to pinned memory,
        # everything enqueued on the plan's stream (the host never waits inside a step; results are read after the loop)
        return fq.enqueue(d_pdb, NFFT)

    fq = features.FeatureQueue(1, dev, st, slots=max(args.steps, args.warmup, 1))
    for _ in range(args.warmup):
        device_step(None)
    sampler = ClockSampler(dev)
    if dist is not None:
        dist.barrier()
    nat.device_sync(dev)
    sampler.start()
    total_timer.start()
    for i in range(args.steps):
        last_slot = device_step(kernel_timers[i])
    total_timer.stop()
    dt_dev = total_timer.elapsed_ms() * 1e-3
    nat.device_sync(dev)
    feat = fq.results(last_slot)[0]
    kernel_ms = [t.elapsed_ms() for t in kernel_timers]
    dt_dev = barrier_max(dist, local, dt_dev)

    # ---------------- end-to-end leg: host (pinned) buffers through the C ABI, copies inside the timed region
    h_wf = nat.pinned_empty((F, NFFT), np.uint8)
    h_we = nat.pinned_empty((1, NFFT), np.float64)
    h_mh = nat.pinned_empty((1, NFFT), np.float32)
    e2e_steps = max(3, min(args.steps, 10))

    def e2e_step():
        r = pl.stft(host_in, wf_rows=h_wf, welch=h_we, maxhold=h_mh, vmin=VMIN, vmax=VMAX)
        _, pdb = pl.welch_finalize(h_we[0], r.n_frames, FS)
        return r, features.measure(pdb, device=dev)

    for _ in range(2):
        r, feat_h = e2e_step()
    if dist is not None:
        dist.barrier()
    nat.device_sync(dev)
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        r, feat_h = e2e_step()
    nat.device_sync(dev)
    dt_e2e_local = time.perf_counter() - t0
    clocks = sampler.stop()     # sampled through both timed regions (device-resident leg and end-to-end leg)
    dt_e2e = barrier_max(dist, local, dt_e2e_local)
    h2d = r.h2d_bytes + NFFT * 8
    d2h = r.d2h_bytes + NFFT * 8 + 160

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---------------- context for the roofline: measured FP32 peak, and the HBM-bound headline shape of north_star
    # (cf32 in, hop = N, f32 dB rows: 12 B/sample, 80 nominal flop/sample) timed the same way on 61 440 000 samples
    # (492 MB in + 246 MB out per launch: several times the 126 MB L2, so no explicit flush)
    p32 = nat.fp32_peak_tflops(dev)
    Lh = L_STEP
    plh = sp.SpectralPlan(NFFT, NFFT, "hann", sp.FMT_CF32, device=dev)
    rngh = np.random.default_rng(1)
    d_xh = nat.DeviceArray.from_host(rngh.standard_normal(2 * Lh).astype(np.float32).view(np.complex64), dev)
    d_dbh = nat.DeviceArray((Lh // NFFT, NFFT), np.float32, dev)
    _, ms_h = plh.time_stft(d_xh, warmup=3, iters=10, flush_l2=False, db_rows=d_dbh)
    ms_h = float(np.median(ms_h))
    headline = {"shape": "cf32 in, 4096-pt Hann, hop = N, f32 dB rows (12 B/sample), 61 440 000 samples per launch (738 MB of traffic, "
                         "larger than L2; no flush)",
                "kernel_ms": round(ms_h, 4), "Msamples_per_s": round(Lh / (ms_h * 1e-3) / 1e6, 1),
                "achieved_gbs": round(Lh * 12 / (ms_h * 1e-3) / 1e9, 1),
                "frac_of_hbm_peak": round(Lh * 12 / (ms_h * 1e-3) / 1e9 / peaks["hbm_gbs"], 4)}
    plh.close(); d_xh.free(); d_dbh.free()

    k_ms = float(np.mean(kernel_ms))
    achieved = L_STEP * BYTES_PER_SAMPLE / (k_ms * 1e-3) / 1e9
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as fh:
            traffic = json.load(fh).get("c2_stft_kernel_dram_bytes_per_launch")
    except Exception:
        pass
    line = {
        "metric": METRIC, "value": round(world * L_STEP * args.steps / dt_dev / 1e6, 1), "unit": "Msamples/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(dt_dev / args.steps * 1e3, 4),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "config2: 61.44 MS/s int16 IQ (1 s = 61 440 000 samples per step per GPU), 4096-pt Hann FFT, "
                               "75% overlap -> u8 waterfall rows + Welch + max-hold + classifier features",
                   "nfft": NFFT, "hop": HOP, "frames_per_step": F, "kernel_variant": variant,
                   "l2": "step input (246 MB) + rows (246 MB) exceed the 126 MB L2; no explicit flush",
                   "timing": "CUDA events on the plan's compute stream around the K steps (max over ranks); "
                             "the STFT kernel additionally bracketed per launch",
                   "multi_gpu": "replicas only: one independent stream per GPU, no data-path collective",
                   "real_time_margin_x": round(L_STEP * args.steps / dt_dev / FS, 1)},
        "e2e": {"value": round(world * L_STEP * e2e_steps / dt_e2e / 1e6, 1), "unit": "Msamples/s",
                "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "steps": e2e_steps,
                "ms_per_step": round(dt_e2e / e2e_steps * 1e3, 3),
                "h2d_gbs": round(h2d * e2e_steps / dt_e2e / 1e9, 2), "d2h_gbs": round(d2h * e2e_steps / dt_e2e / 1e9, 2)},
        "gpu_launches": 3 * args.steps,
        "roofline": {"bound": "hbm", "achieved": round(achieved, 1), "peak": peaks["hbm_gbs"], "unit": "GB/s",
                     "frac": round(achieved / peaks["hbm_gbs"], 4), "traffic": traffic, "peak_source": peak_src,
                     "kernel": "stft_kernel<4096,ci16,acc>", "kernel_ms": round(k_ms, 4),
                     "bytes_per_sample": BYTES_PER_SAMPLE, "kernel_share_of_step": round(k_ms * args.steps / (dt_dev * 1e3), 3),
                     "fp32_tflops_nominal": round(L_STEP * FLOP_PER_SAMPLE / (k_ms * 1e-3) / 1e12, 2),
                     "fp32_peak_tflops_measured": round(p32, 2),
                     "frac_of_fp32_peak_nominal_flops": round(L_STEP * FLOP_PER_SAMPLE / (k_ms * 1e-3) / 1e12 / p32, 4),
                     "note": "this shape transforms every sample 4 times (75% overlap): 320 nominal flop/sample vs 8 B/sample, "
                             "so FP32 issue binds before HBM (SURVEY 8d); the HBM-bound shape of north_star is in headline_shape",
                     "headline_shape": headline},
        "clocks": clocks,
        "features": {"snr_db": round(feat["snr_db"], 2), "peak_count": feat["peak_count"]},
    }
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline_single()
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
