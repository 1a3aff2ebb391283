#!/usr/bin/env python3
"""Sharded multi-GPU configs of BASELINE.json, one process per GPU under torchrun:

  c5: one 2^L-sample cf32 capture, 65536-pt Hann STFT, 50 % overlap, frame blocks with (N - hop) halos per rank;
      u8 waterfall rows collected on rank 0, Welch / max-hold reduced.  --collective fused: every rank's kernel
      writes rows and reduces accumulators straight into rank 0's HBM over NVLink (peer memory, system-scope
      atomics); --collective nccl: local outputs, then NCCL all-reduce + gather (the baseline plumbing).
  c4: 64 independent cf32 streams x 2^M samples, 2048-pt Hann, 50 % overlap, streams split across ranks,
      per-stream Welch PSD + classifier features, features all-gathered (no data-path collective).

  python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 --master-port P \
      tools/bench_sharded.py --config c5 --collective fused --steps 5 [--check]

Strong scaling: total work is fixed; value = total samples / max-over-ranks step time (barrier + device sync on
both sides of the timed steps).  --check verifies rank 0's result against the float64 numpy checker on a prefix.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def fill_device(nat, lib, darr, block, device):
    """Tile a host block over a device array with plain H2D copies (synthetic capture without 8 GiB of host RAM)."""
    nb = block.nbytes
    off = 0
    while off < darr.nbytes:
        n = min(nb, darr.nbytes - off)
        nat.check(lib.spx_memcpy_h2d(device, darr.ptr + off, block.ctypes.data, n))
        off += n


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="c5", choices=["c4", "c5"])
    ap.add_argument("--collective", default="fused", choices=["fused", "nccl"])
    ap.add_argument("--log2-samples", type=int, default=None, help="c5: capture length (default 30); c4: per-stream length (default 24)")
    ap.add_argument("--streams", type=int, default=64)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--check", action="store_true")
    args = ap.parse_args()

    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from sdr_iq_visualizer_b200 import _native as nat, dist as sd, features, spectral as sp, synth
    lib = nat.lib()
    dev = local

    def barrier():
        nat.device_sync(dev)
        if world > 1:
            dist.barrier()

    def max_over_ranks(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=f"cuda:{local}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    block = synth.synth_iq(1 << 22, seed=5, tone_cycles_per_sample=20000.37 / 65536).astype(np.complex64)
    out = {"config": args.config, "n_gpus": world, "steps": args.steps, "warmup": args.warmup}

    if args.config == "c5":
        N, hop = 65536, 32768
        L = 1 << (args.log2_samples or 30)
        F = (L - N) // hop + 1
        sh = sd.capture_shard(L, N, hop, rank, world)
        # the rank's slice of the capture (tiled synthetic block, phase-continuous per 2^22 samples): absolute sample
        # s of the capture is block[s mod 2^22], so every rank holds exactly the samples a file reader would give it
        d_in = nat.DeviceArray((sh.n_samples,), np.complex64, dev)
        start = sh.sample0 % block.size
        rolled = np.ascontiguousarray(np.roll(block, -start))
        fill_device(nat, lib, d_in, rolled, dev)
        pl = sp.SpectralPlan(N, hop, "hann", sp.FMT_CF32, device=dev)
        vmin, vmax = -20.0, 110.0
        F_local = sh.f1 - sh.f0
        if args.collective == "fused":
            target = sd.PeerReduceTarget(N, F, rank, world, dev, dst=0, want_rows=True)

            def step():
                target.zero()
                barrier()                       # accumulators are clean before anyone reduces into them
                sd.fused_capture_step(pl, d_in, sh, target, vmin, vmax)
                pl.sync()
                barrier()                       # every rank's remote writes have landed on rank 0
        else:
            welch = torch.zeros((1, N), dtype=torch.float64, device=f"cuda:{local}")
            mh = torch.zeros((1, N), dtype=torch.float32, device=f"cuda:{local}")
            rows = torch.empty((F_local, N), dtype=torch.uint8, device=f"cuda:{local}")
            gathered = [None]

            def step():
                barrier()
                pl.stft(d_in, wf_rows=rows, welch=welch, maxhold=mh, vmin=vmin, vmax=vmax, n_samples=sh.n_samples)
                pl.sync()
                if world > 1:
                    sd.allreduce_partials(welch, mh, F_local)
                    gathered[0] = sd.gather_rows(rows, 0)
                else:
                    gathered[0] = rows
                barrier()

        for _ in range(args.warmup):
            step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step()
        barrier()
        dt = max_over_ranks(time.perf_counter() - t0)
        # compute-only time of this rank's kernels (no collective), for the overlap figure
        _, ms = pl.time_stft(d_in, warmup=1, iters=3, wf_rows=nat.DeviceArray((F_local, N), np.uint8, dev),
                             welch=nat.DeviceArray((1, N), np.float64, dev), maxhold=nat.DeviceArray((1, N), np.float32, dev),
                             vmin=vmin, vmax=vmax, n_samples=sh.n_samples)
        k_ms = max_over_ranks(float(np.median(ms)))
        out.update({"metric": "IQ Msamples/s through windowed FFT->PSD->waterfall", "unit": "Msamples/s", "scaling": "strong",
                    "value": round(L * args.steps / dt / 1e6, 1), "ms_per_step": round(dt / args.steps * 1e3, 3),
                    "kernel_only_ms_max_rank": round(k_ms, 3), "collective": args.collective,
                    "workload": f"config5: 2^{int(np.log2(L))} cf32 samples, 65536-pt Hann, 50% overlap, {F} frames, "
                                f"u8 rows ({F * N / 2**30:.2f} GiB) on rank 0 + Welch/max-hold reduced",
                    "frames_local": F_local, "halo_samples": sh.halo})
        if args.check:
            # rank 0 compares the first frames and the reduced Welch sum of a prefix-sized run with the float64 checker
            from oracle import spectral_ref as sref
            from tests import parity
            ok = True
            if rank == 0:
                if args.collective == "fused":
                    rows0 = target.rows.rows(0, 4)
                    got_rows = np.empty((4, N), np.uint8)
                    nat.check(lib.spx_memcpy_d2h(dev, got_rows.ctypes.data, rows0.ptr, got_rows.nbytes))
                    got_w = target.buffers["welch"].array.to_host()[0]
                    got_m = target.buffers["maxhold"].array.to_host()[0]
                    last_rows = np.empty((2, N), np.uint8)
                    nat.check(lib.spx_memcpy_d2h(dev, last_rows.ctypes.data, target.rows.rows(F - 2, F).ptr, last_rows.nbytes))
                else:
                    got_rows = gathered[0][:4].cpu().numpy()
                    last_rows = gathered[0][F - 2:F].cpu().numpy()
                    got_w, got_m = welch[0].cpu().numpy(), mh[0].cpu().numpy()
                x = np.tile(block, -(-L // block.size))[:L] if L <= (1 << 25) else None
                xs = block if x is None else x
                X = sref.shift_bins(sref.stft(sref.as_complex128(xs[: N + 3 * hop]), N, hop, "hann"))
                parity.check_u8(got_rows, sref.amplitude_db(X), vmin, vmax, what="first rows")
                if x is not None:
                    Xa = sref.shift_bins(sref.stft(sref.as_complex128(x), N, hop, "hann"))
                    P = Xa.real**2 + Xa.imag**2
                    parity.check_power(got_w, P.sum(axis=0), what="reduced welch")
                    parity.check_power(got_m, P.max(axis=0), what="reduced maxhold")
                    parity.check_u8(last_rows, sref.amplitude_db(Xa[-2:]), vmin, vmax, what="last rows (last rank)")
                out["check"] = "ok (rows of first/last rank, reduced Welch and max-hold vs the float64 checker)" if x is not None \
                    else "ok (first rows only: capture too long for the CPU checker)"
    else:
        N, hop = 2048, 1024
        Ls = 1 << (args.log2_samples or 24)
        S = args.streams
        s0, s1 = sd.stream_block(S, rank, world)
        ns = s1 - s0
        d_in = nat.DeviceArray((max(ns, 1) * Ls,), np.complex64, dev)
        # stream s = the synthetic block rotated by s * 4099 samples and scaled: distinct spectra per stream
        for i, s in enumerate(range(s0, s1)):
            blk = np.ascontiguousarray(np.roll(block, -(s * 4099))) * np.float32(1.0 + 0.01 * s)
            view = nat.DeviceView(d_in.ptr + i * Ls * 8, (Ls,), np.complex64, dev)
            view.nbytes = Ls * 8
            fill_device(nat, lib, view, blk.astype(np.complex64), dev)
        pl = sp.SpectralPlan(N, hop, "hann", sp.FMT_CF32, device=dev)
        F = pl.frame_count(Ls)
        d_we = nat.DeviceArray((max(ns, 1), N), np.float64, dev)
        d_pdb = nat.DeviceArray((max(ns, 1), N), np.float64, dev)
        d_pxx = nat.DeviceArray((max(ns, 1), N), np.float64, dev)
        import ctypes as C
        feats = [None]
        FS_BYTES = C.sizeof(nat.spx_features)
        per = -(-S // world)                      # streams per rank, padded so that every rank contributes the same size
        t_mine = torch.zeros((per * FS_BYTES,), dtype=torch.uint8, device=f"cuda:{local}")
        t_all = torch.zeros((world * per * FS_BYTES,), dtype=torch.uint8, device=f"cuda:{local}")
        opts = nat.spx_feature_opts()
        opts.drop_db[0], opts.drop_db[1], opts.drop_db[2] = 3.0, 10.0, 20.0

        ext = torch.cuda.ExternalStream(pl.stream, device=torch.device("cuda", local))   # the plan's compute stream, seen by torch
        ev_feat, ev_gather = torch.cuda.Event(), torch.cuda.Event()
        state = {"gathered": False}

        def step():
            # independent streams: no barrier and no host wait inside the step; the only exchange is the all-gather of
            # the feature structs (device to device over NCCL, a few KiB), ordered against the kernels with events
            if ns:
                pl.stft(d_in, n_streams=ns, welch=d_we, n_samples=Ls)
                pl.welch_finalize(d_we, F, 61.44e6, pxx=d_pxx, pdb=d_pdb, n_streams=ns)
                if state["gathered"]:
                    ext.wait_event(ev_gather)      # the previous all-gather has finished reading t_mine
                nat.check(lib.spx_classify_features_dev(dev, d_pdb.ptr, 1, N, ns, N, int(t_mine.data_ptr()), None, 0,
                                                        C.byref(opts), pl.stream))
            ev_feat.record(ext)
            cur = torch.cuda.current_stream()
            cur.wait_event(ev_feat)
            if world > 1:
                dist.all_gather_into_tensor(t_all, t_mine)
            else:
                t_all.copy_(t_mine)
            ev_gather.record(cur)
            state["gathered"] = True

        def parse():
            raw = t_all.cpu().numpy().tobytes()
            res = []
            for r in range(world):
                a, b = sd.stream_block(S, r, world)
                arr = (nat.spx_features * (b - a)).from_buffer_copy(raw[r * per * FS_BYTES:(r * per + (b - a)) * FS_BYTES])
                res.append([(f.snr_db, f.last_3db - f.first_3db, f.last_10db - f.first_10db, f.last_20db - f.first_20db,
                             f.flatness, f.kurtosis, f.peak_count) for f in arr])
            return res

        for _ in range(args.warmup):
            step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step()
        barrier()
        dt = max_over_ranks(time.perf_counter() - t0)
        out.update({"metric": "IQ Msamples/s through windowed FFT->PSD->waterfall", "unit": "Msamples/s", "scaling": "strong",
                    "value": round(S * Ls * args.steps / dt / 1e6, 1), "ms_per_step": round(dt / args.steps * 1e3, 3),
                    "workload": f"config4: {S} streams x 2^{int(np.log2(Ls))} cf32, 2048-pt Hann, 50% overlap, per-stream Welch PSD "
                                f"+ classifier features; streams {s0}..{s1 - 1} on rank {rank}",
                    "features_gathered": sum(len(f) for f in parse())})
        if args.check and rank == 0 and ns:
            from oracle import classifier_ref as cref, spectral_ref as sref
            from tests import parity
            blk = (np.ascontiguousarray(np.roll(block, -(s0 * 4099))) * np.float32(1.0 + 0.01 * s0)).astype(np.complex64)
            x = np.tile(blk, -(-Ls // blk.size))[:Ls]
            fr, pxx = sref.welch_psd(x, N, hop, "hann", 61.44e6, 0.0)
            got = d_pxx.to_host()[0]
            parity.check_power(got, pxx, what="c4 stream welch")
            f = cref.features(fr, 10 * np.log10(pxx))
            m = features.measure_batch(d_pdb, n=N, batch=ns, device=dev)[0]
            assert abs(m["snr_db"] - f["snr_db"]) < 2e-3, (m["snr_db"], f["snr_db"])
            out["check"] = "ok (stream %d Welch PSD and SNR vs the float64 checker)" % s0
    if rank == 0:
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
