#!/usr/bin/env python3
"""Sharded multi-GPU configs of BASELINE.json, one process per GPU under torchrun.
bench.py imports `run_c5` / `run_c4` and prints their results inside its JSON line (`sharded`), so the driver's
1/2/4/8-GPU scaling record carries the configs that actually shard; this file is also a CLI for single cases.

  c5: one 2^L-sample cf32 capture, 65536-pt Hann STFT, 50 % overlap, frame blocks with (N - hop) halos per rank.
      `rows`:  "sharded" -- uint8 rows stay on the rank that produced them (sharded by frame block), only the Welch
               sum / max-hold are reduced;  "gather" -- every row is also collected on rank 0.
      `collective`: "fused" -- the STFT kernels reduce into rank 0's HBM over NVLink (CUDA-IPC peer memory,
               system-scope atomics) and rows are pushed by the copy engine piece by piece behind the transform;
               "nccl" -- local outputs, then NCCL all-reduce (+ gather): the plain-collective baseline.
  c4: 64 independent cf32 streams x 2^M samples, 2048-pt Hann, 50 % overlap, streams split across ranks, per-stream
      Welch PSD + classifier features, features all-gathered (no data-path collective).

Strong scaling: total work is fixed; value = total samples / max-over-ranks time of K back-to-back steps bracketed by
barrier + device sync.  Every step reduces into its own (pre-zeroed) accumulator target, so no barrier is needed
between steps.  `check=True` verifies the result of the LAST step against the float64 numpy checker at full size:
the synthetic capture is periodic (a 2^22-sample block tiled; 2^22 = 128 hops), so the checker evaluates 128 frames
and every reduced bin / checked row of the 2^30-sample run has an exact float64 counterpart.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "IQ Msamples/s through windowed FFT->PSD->waterfall"
BLOCK_LOG2 = 22


class Ctx:
    """Rank context shared with bench.py: torch.distributed is the plumbing (barriers, handles, tiny tensors)."""

    def __init__(self, rank, world, local, dist, strict=True):
        from sdr_iq_visualizer_b200 import _native as nat
        self.rank, self.world, self.local, self.dist, self.nat = rank, world, local, dist, nat
        self.strict = strict      # a failed parity check aborts (CLI, tests); bench.py records "FAILED" in its line instead
        self.lib = nat.lib()
        self.dev = local

    def barrier(self):
        self.nat.device_sync(self.dev)
        if self.world > 1:
            self.dist.barrier()

    def reduce_max(self, v):
        if self.world == 1:
            return float(v)
        import torch
        t = torch.tensor([float(v)], dtype=torch.float64, device=f"cuda:{self.local}")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def reduce_min(self, v):
        return -self.reduce_max(-float(v))


def synth_block():
    from sdr_iq_visualizer_b200 import synth
    return synth.synth_iq(1 << BLOCK_LOG2, seed=5, tone_cycles_per_sample=20000.37 / 65536).astype(np.complex64)


def fill_device(ctx, darr, block):
    """Tile a host block over a device array with plain H2D copies (synthetic capture without 8 GiB of host RAM)."""
    nb, off = block.nbytes, 0
    while off < darr.nbytes:
        n = min(nb, darr.nbytes - off)
        ctx.nat.check(ctx.lib.spx_memcpy_h2d(ctx.dev, darr.ptr + off, block.ctypes.data, n))
        off += n


def _rows_to_host(ctx, view, r0, r1, N):
    out = np.empty((r1 - r0, N), np.uint8)
    ctx.nat.check(ctx.lib.spx_memcpy_d2h(ctx.dev, out.ctypes.data, view.ptr + r0 * N, out.nbytes))
    return out


class PeriodicChecker:
    """float64 checker of the tiled synthetic capture: frame f of the capture == frame (f mod 128) of one period."""

    def __init__(self, block, N, hop):
        from oracle import spectral_ref as sref
        self.sref = sref
        self.period = block.size // hop
        assert self.period * hop == block.size
        x = np.concatenate([block, block[:N]])
        X = sref.shift_bins(sref.stft(sref.as_complex128(x[: (self.period - 1) * hop + N]), N, hop, "hann"))
        assert X.shape[0] == self.period
        self.X = X
        self.P = X.real ** 2 + X.imag ** 2

    def welch_and_max(self, f0, f1):
        counts = np.bincount(np.arange(f0, f1) % self.period, minlength=self.period).astype(np.float64)
        present = counts > 0
        return counts @ self.P, self.P[present].max(axis=0)

    def db_rows(self, frames):
        return self.sref.amplitude_db(self.X[np.asarray(frames) % self.period])


def run_c5(ctx, collective="fused", rows="gather", log2_samples=30, steps=5, warmup=2, check=True, block=None,
           measure_kernel=True):
    import torch
    from sdr_iq_visualizer_b200 import dist as sd, spectral as sp
    nat, dev, rank, world = ctx.nat, ctx.dev, ctx.rank, ctx.world
    N, hop = 65536, 32768
    L = 1 << log2_samples
    F = (L - N) // hop + 1
    block = synth_block() if block is None else block
    sh = sd.capture_shard(L, N, hop, rank, world)
    F_local = sh.f1 - sh.f0
    d_in = nat.DeviceArray((max(sh.n_samples, 1),), np.complex64, dev)
    # absolute sample s of the capture is block[s mod 2^22]: every rank holds exactly what a file reader would give it
    fill_device(ctx, d_in, np.ascontiguousarray(np.roll(block, -(sh.sample0 % block.size))))
    pl = sp.SpectralPlan(N, hop, "hann", sp.FMT_CF32, device=dev)
    vmin, vmax = -20.0, 110.0
    n_targets = warmup + steps
    gather = rows == "gather"
    out = {"collective": collective, "rows": "gathered on rank 0" if gather else "sharded by frame block (left on the producing rank)",
           "frames": F, "frames_local": F_local, "halo_samples": sh.halo}
    local_rows = None
    if collective == "fused":
        # one accumulator target per step (zeroed up front): steps run back to back without a barrier in between
        targets = [sd.PeerReduceTarget(N, F, rank, world, dev, dst=0, want_rows=gather and k == 0) for k in range(n_targets)]
        rows_view = targets[0].rows
        if not gather:
            local_rows = nat.DeviceArray((max(F_local, 1), N), np.uint8, dev)
        for t in targets:
            t.zero()
        ctx.barrier()

        def step(k):
            t = targets[k]
            if F_local <= 0:
                return
            if gather:
                dst_rows = rows_view.rows(sh.f0, sh.f1)
                mode = 1 if rank == 0 else 3      # owner writes its rows directly; others stage + copy-engine push
            else:
                dst_rows, mode = local_rows, 1    # rows stay here; accumulators go to the owner
            pl.stft(d_in, wf_rows=dst_rows, welch=t.welch, maxhold=t.maxhold, vmin=vmin, vmax=vmax, accumulate=True,
                    n_samples=sh.n_samples, peer_outputs=mode)

        def result_acc():
            t = targets[n_targets - 1]
            return t.buffers["welch"].array.to_host()[0], t.buffers["maxhold"].array.to_host()[0]
    else:
        tdev = torch.device("cuda", ctx.local)
        welch = torch.zeros((1, N), dtype=torch.float64, device=tdev)
        mh = torch.zeros((1, N), dtype=torch.float32, device=tdev)
        local_rows = torch.empty((max(F_local, 1), N), dtype=torch.uint8, device=tdev)
        gathered = [None]

        def step(k):
            if F_local > 0:
                pl.stft(d_in, wf_rows=local_rows[:F_local], welch=welch, maxhold=mh, vmin=vmin, vmax=vmax, n_samples=sh.n_samples)
            pl.sync()
            if world > 1:
                sd.allreduce_partials(welch, mh, F_local)
                if gather:
                    gathered[0] = sd.gather_rows(local_rows[:F_local], 0)
            elif gather:
                gathered[0] = local_rows[:F_local]

        def result_acc():
            return welch[0].cpu().numpy(), mh[0].cpu().numpy()

    for k in range(warmup):
        step(k)
    pl.sync()
    ctx.barrier()
    t0 = time.perf_counter()
    for k in range(steps):
        step(warmup + k)
    pl.sync()
    if collective == "nccl":
        torch.cuda.synchronize()
    ctx.barrier()
    dt = ctx.reduce_max(time.perf_counter() - t0)
    out.update({"value": round(L * steps / dt / 1e6, 1), "unit": "Msamples/s", "ms_per_step": round(dt / steps * 1e3, 3)})

    if measure_kernel:
        # this rank's kernels alone (local outputs, no peer traffic): what the collective has to hide behind
        tw = nat.DeviceArray((1, N), np.float64, dev)
        tm = nat.DeviceArray((1, N), np.float32, dev)
        tr = local_rows if isinstance(local_rows, nat.DeviceArray) else nat.DeviceArray((max(F_local, 1), N), np.uint8, dev)
        if F_local > 0:
            _, ms = pl.time_stft(d_in, warmup=1, iters=3, wf_rows=tr, welch=tw, maxhold=tm, vmin=vmin, vmax=vmax, n_samples=sh.n_samples)
            k_ms = float(np.median(ms))
        else:
            k_ms = 0.0
        out["kernel_ms_max_rank"] = round(ctx.reduce_max(k_ms), 3)
        tw.free(); tm.free()
        if tr is not local_rows:
            tr.free()

    if check:
        from tests import parity
        chk = PeriodicChecker(block, N, hop)
        ok = 1.0
        try:
            if rank == 0:
                got_w, got_m = result_acc()
                want_w, want_m = chk.welch_and_max(0, F)
                parity.check_power(got_w, want_w, what=f"c5 {collective}/{rows} reduced welch")
                parity.check_power(got_m, want_m, what=f"c5 {collective}/{rows} reduced maxhold")
            if gather:
                if rank == 0:
                    # first / last rows and both sides of every shard boundary
                    fr = sorted({0, 1, F - 2, F - 1} | {f for r in range(1, world) for f in
                                                        (sd.frame_block(F, r, world)[0] - 1, sd.frame_block(F, r, world)[0])})
                    fr = [f for f in fr if 0 <= f < F]
                    if collective == "fused":
                        got = np.concatenate([_rows_to_host(ctx, rows_view, f, f + 1, N) for f in fr])
                    else:
                        got = gathered[0][torch.tensor(fr, device=gathered[0].device)].cpu().numpy()
                    parity.check_u8(got, chk.db_rows(fr), vmin, vmax, what=f"c5 {collective} gathered rows at shard boundaries")
            elif F_local > 0:
                fr = sorted({sh.f0, sh.f1 - 1})
                if isinstance(local_rows, nat.DeviceArray):
                    got = np.concatenate([_rows_to_host(ctx, local_rows, f - sh.f0, f - sh.f0 + 1, N) for f in fr])
                else:
                    got = local_rows[torch.tensor([f - sh.f0 for f in fr], device=local_rows.device)].cpu().numpy()
                parity.check_u8(got, chk.db_rows(fr), vmin, vmax, what=f"c5 {collective} local rows of rank {rank}")
        except AssertionError as exc:
            ok = 0.0
            print(f"[rank {rank}] c5 check FAILED: {exc}", file=sys.stderr, flush=True)
        ok = ctx.reduce_min(ok)
        out["check"] = ("ok: reduced Welch / max-hold of all %d frames and rows at every shard boundary vs the float64 checker" % F) \
            if ok > 0.5 else "FAILED"
        if ok < 0.5 and ctx.strict:
            raise SystemExit("c5 parity check failed")

    if gather and collective == "fused":
        # NVLink ingress floor of the gather: the row bytes of ranks 1..G-1 pushed by the copy engines alone, all at once
        # (after the parity check: the probe overwrites the gathered rows with scratch data)
        remote = (F - (sd.frame_block(F, 0, world)[1])) * N
        out["gather_remote_bytes"] = int(remote)
        if world > 1:
            src = nat.DeviceArray((max(F_local, 1) * N,), np.uint8, dev)
            best = 1e30
            for _ in range(3):
                ctx.barrier()
                t1 = time.perf_counter()
                if rank != 0 and F_local > 0:
                    nat.check(ctx.lib.spx_memcpy_d2d_async(dev, rows_view.rows(sh.f0, sh.f1).ptr, src.ptr, F_local * N, pl.stream))
                pl.sync()
                ctx.barrier()
                best = min(best, ctx.reduce_max(time.perf_counter() - t1))
            src.free()
            out["nvlink_ingress_gbs_measured"] = round(remote / best / 1e9, 1)
            out["gather_floor_ms"] = round(best * 1e3, 3)
            # the step can be no faster than the slower of this rank's kernels and the NVLink ingress of rank 0
            bound = max(best * 1e3, out.get("kernel_ms_max_rank", 0.0))
            out["gather_bound_ms"] = round(bound, 3)
            out["gather_vs_bound"] = round(out["ms_per_step"] / bound, 3)
            out["nvlink_gbs_achieved"] = round(remote / (out["ms_per_step"] * 1e-3) / 1e9, 1)

    if collective == "fused":
        for t in targets:
            t.close()
    pl.close()
    d_in.free()
    return out


def run_c4(ctx, log2_samples=24, streams=64, steps=5, warmup=2, check=True, block=None):
    import ctypes as C
    import torch
    from sdr_iq_visualizer_b200 import dist as sd, features, spectral as sp
    nat, dev, rank, world, lib = ctx.nat, ctx.dev, ctx.rank, ctx.world, ctx.lib
    N, hop = 2048, 1024
    Ls, S = 1 << log2_samples, streams
    block = synth_block() if block is None else block
    s0, s1 = sd.stream_block(S, rank, world)
    ns = s1 - s0
    d_in = nat.DeviceArray((max(ns, 1) * Ls,), np.complex64, dev)

    def stream_block_host(s):
        # stream s = the synthetic block rotated by s * 4099 samples and scaled: distinct spectra per stream
        return (np.ascontiguousarray(np.roll(block, -(s * 4099))) * np.float32(1.0 + 0.01 * s)).astype(np.complex64)

    for i, s in enumerate(range(s0, s1)):
        view = nat.DeviceView(d_in.ptr + i * Ls * 8, (Ls,), np.complex64, dev)
        fill_device(ctx, view, stream_block_host(s))
    pl = sp.SpectralPlan(N, hop, "hann", sp.FMT_CF32, device=dev)
    F = pl.frame_count(Ls)
    d_we = nat.DeviceArray((max(ns, 1), N), np.float64, dev)
    d_pdb = nat.DeviceArray((max(ns, 1) * N,), np.float64, dev)
    d_pxx = nat.DeviceArray((max(ns, 1) * N,), np.float64, dev)
    FS_BYTES = C.sizeof(nat.spx_features)
    per = -(-S // world)                      # streams per rank, padded so that every rank contributes the same size
    tdev = torch.device("cuda", ctx.local)
    t_mine = torch.zeros((per * FS_BYTES,), dtype=torch.uint8, device=tdev)
    t_all = torch.zeros((world * per * FS_BYTES,), dtype=torch.uint8, device=tdev)
    opts = nat.spx_feature_opts()
    opts.drop_db[0], opts.drop_db[1], opts.drop_db[2] = 3.0, 10.0, 20.0
    ext = torch.cuda.ExternalStream(pl.stream, device=tdev)   # the plan's compute stream, seen by torch
    ev_feat, ev_gather = torch.cuda.Event(), torch.cuda.Event()
    state = {"gathered": False}
    d_we_v = nat.DeviceView(d_we.ptr, (max(ns, 1) * N,), np.float64, dev)

    def step():
        # independent streams: no barrier and no host wait inside the step; the only exchange is the all-gather of the
        # feature structs (device to device over NCCL, a few KiB), ordered against the kernels with events
        if ns:
            pl.stft(d_in, n_streams=ns, welch=d_we, n_samples=Ls)
            pl.welch_finalize(d_we_v, F, 61.44e6, pxx=d_pxx, pdb=d_pdb, n_streams=ns)
            if state["gathered"]:
                ext.wait_event(ev_gather)      # the previous all-gather has finished reading t_mine
            nat.check(lib.spx_classify_features_dev(dev, d_pdb.ptr, 1, N, ns, N, int(t_mine.data_ptr()), None, 0,
                                                    C.byref(opts), pl.stream))
        ev_feat.record(ext)
        cur = torch.cuda.current_stream()
        cur.wait_event(ev_feat)
        if world > 1:
            ctx.dist.all_gather_into_tensor(t_all, t_mine)
        else:
            t_all.copy_(t_mine)
        ev_gather.record(cur)
        state["gathered"] = True

    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    ctx.barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    torch.cuda.synchronize()
    ctx.barrier()
    dt = ctx.reduce_max(time.perf_counter() - t0)
    raw = t_all.cpu().numpy().tobytes()
    n_feats = 0
    for r in range(world):
        a, b = sd.stream_block(S, r, world)
        n_feats += len((nat.spx_features * (b - a)).from_buffer_copy(raw[r * per * FS_BYTES:(r * per + (b - a)) * FS_BYTES]))
    out = {"value": round(S * Ls * steps / dt / 1e6, 1), "unit": "Msamples/s", "ms_per_step": round(dt / steps * 1e3, 3),
           "streams": S, "streams_local": ns, "features_gathered": n_feats,
           "collective": "none on the data path; all-gather of %d B of feature structs per rank" % (per * FS_BYTES)}
    if check:
        ok = 1.0
        try:
            if rank == 0 and ns:
                from oracle import classifier_ref as cref, spectral_ref as sref
                from tests import parity
                blk = stream_block_host(s0)
                x = np.tile(blk, -(-Ls // blk.size))[:Ls]
                fr, pxx = sref.welch_psd(x, N, hop, "hann", 61.44e6, 0.0)
                got = d_pxx.to_host()[:N]
                parity.check_power(got, pxx, what="c4 stream welch")
                f = cref.features(fr, 10 * np.log10(pxx))
                m = features.measure_batch(d_pdb, n=N, batch=ns, device=dev)[0]
                assert abs(m["snr_db"] - f["snr_db"]) < 2e-3, (m["snr_db"], f["snr_db"])
                assert n_feats == S
        except AssertionError as exc:
            ok = 0.0
            print(f"[rank {rank}] c4 check FAILED: {exc}", file=sys.stderr, flush=True)
        ok = ctx.reduce_min(ok)
        out["check"] = "ok: stream %d Welch PSD and SNR vs the float64 checker, %d feature structs gathered" % (s0, n_feats) \
            if ok > 0.5 else "FAILED"
        if ok < 0.5 and ctx.strict:
            raise SystemExit("c4 parity check failed")
    pl.close()
    d_in.free()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="c5", choices=["c4", "c5"])
    ap.add_argument("--collective", default="fused", choices=["fused", "nccl"])
    ap.add_argument("--rows", default="gather", choices=["gather", "sharded"])
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
    ctx = Ctx(rank, world, local, dist if world > 1 else None)
    if args.config == "c5":
        res = run_c5(ctx, args.collective, args.rows, args.log2_samples or 30, args.steps, args.warmup, args.check)
    else:
        res = run_c4(ctx, args.log2_samples or 24, args.streams, args.steps, args.warmup, args.check)
    res.update({"config": args.config, "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "metric": METRIC, "scaling": "strong"})
    if rank == 0:
        print(json.dumps(res), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
