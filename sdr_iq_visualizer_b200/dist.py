"""Multi-GPU sharding of the spectral path -- one process per GPU.

The reference is single-process; this is new capability.  Only what shards naturally is sharded:
  * independent streams (config 4): stream s goes to rank ``s*world//n_streams`` (contiguous blocks);
    no data-path collective, the per-stream features are all-gathered (a few hundred bytes each);
  * one long capture (config 5): frames [f0, f1) per rank; the rank reads samples
    [f0*hop, (f1-1)*hop + N) itself, so the (N - hop)-sample halo is read twice from the source and
    never exchanged between GPUs.  The partial Welch sums (float64, SUM), max-holds (float32, MAX)
    and frame counts (SUM) are all-reduced and the uint8 waterfall rows are gathered to one rank.
``torch.distributed`` is the plumbing (NCCL over NVLink for CUDA tensors; gloo in the CPU tests).
A single live stream (config 2) does not shard: run replicas (one stream per GPU).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Tuple

import numpy as np


# ----------------------------------------------------------------------------- partitioning (pure host logic)
def stream_block(n_streams: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block [s0, s1) of streams owned by `rank` (sizes differ by at most one)."""
    base, extra = divmod(n_streams, world)
    s0 = rank * base + min(rank, extra)
    return s0, s0 + base + (1 if rank < extra else 0)


def frame_block(n_frames: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block [f0, f1) of frames owned by `rank`."""
    return stream_block(n_frames, rank, world)


@dataclass
class CaptureShard:
    f0: int           # first frame
    f1: int           # one past the last frame
    sample0: int      # first sample the rank must read
    n_samples: int    # samples to read: (f1 - f0 - 1) * hop + nfft, 0 if the rank owns no frame
    halo: int         # samples shared with the next rank (nfft - hop), 0 for the last owner


def capture_shard(n_samples: int, nfft: int, hop: int, rank: int, world: int) -> CaptureShard:
    """Frame-block shard of one long capture with its hop halo (config 5)."""
    F = 0 if n_samples < nfft else (n_samples - nfft) // hop + 1
    f0, f1 = frame_block(F, rank, world)
    if f1 <= f0:
        return CaptureShard(f0, f1, f0 * hop, 0, 0)
    return CaptureShard(f0, f1, f0 * hop, (f1 - f0 - 1) * hop + nfft, (nfft - hop) if f1 < F else 0)


# ----------------------------------------------------------------------------- collectives
def _dist():
    import torch.distributed as dist
    return dist


def allreduce_partials(welch_acc, maxhold, n_frames, group=None):
    """In-place all-reduce of the per-rank partials: welch (float64 tensor, SUM), maxhold (float32
    tensor, MAX); returns the global frame count.  Tensors may be CUDA (NCCL) or CPU (gloo)."""
    import torch
    dist = _dist()
    if welch_acc is not None:
        dist.all_reduce(welch_acc, op=dist.ReduceOp.SUM, group=group)
    if maxhold is not None:
        dist.all_reduce(maxhold, op=dist.ReduceOp.MAX, group=group)
    dev = welch_acc.device if welch_acc is not None else (maxhold.device if maxhold is not None else "cpu")
    cnt = torch.tensor([int(n_frames)], dtype=torch.int64, device=dev)
    dist.all_reduce(cnt, op=dist.ReduceOp.SUM, group=group)
    return int(cnt.item())


def gather_rows(rows, dst: int = 0, group=None):
    """Gather per-rank row blocks ([F_r, N] uint8 or float32 tensors, F_r may differ) to rank `dst`
    in frame order.  Returns the concatenated tensor on `dst`, None elsewhere."""
    import torch
    dist = _dist()
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    n = torch.tensor([rows.shape[0]], dtype=torch.int64, device=rows.device)
    counts = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(counts, n, group=group)
    counts = [int(c.item()) for c in counts]
    fmax = max(counts)
    if fmax == 0:
        return rows if rank == dst else None
    padded = rows
    if rows.shape[0] < fmax:
        padded = torch.zeros((fmax,) + tuple(rows.shape[1:]), dtype=rows.dtype, device=rows.device)
        padded[: rows.shape[0]] = rows
    bufs = [torch.empty_like(padded) for _ in range(world)] if rank == dst else None
    dist.gather(padded.contiguous(), bufs, dst=dst, group=group)
    if rank != dst:
        return None
    return torch.cat([b[:c] for b, c in zip(bufs, counts)], dim=0)


def allgather_objects(obj, group=None) -> list:
    """Small python objects (per-stream feature dicts) from every rank, in rank order."""
    dist = _dist()
    out = [None] * dist.get_world_size(group)
    dist.all_gather_object(out, obj, group=group)
    return out


# ----------------------------------------------------------------------------- GPU drivers
def sharded_capture_stft(plan, load_samples, n_samples: int, rank: int, world: int, device: int, *, wf_rows=True,
                         vmin=-100.0, vmax=0.0, gather_to: Optional[int] = 0, group=None):
    """Config-5 driver.  `load_samples(sample0, count)` returns the rank's slice as a numpy array
    (cf32 or interleaved int16) -- e.g. a memmap of the SigMF data file.  Returns a dict with the
    all-reduced Welch sum / max-hold (torch CUDA tensors), the global frame count and, on
    `gather_to`, the gathered uint8 rows."""
    import torch
    from . import _native as nat
    sh = capture_shard(n_samples, plan.nfft, plan.hop, rank, world)
    N = plan.nfft
    dev = torch.device("cuda", device)
    welch = torch.zeros((1, N), dtype=torch.float64, device=dev)
    mh = torch.zeros((1, N), dtype=torch.float32, device=dev)
    F_local = sh.f1 - sh.f0
    rows = torch.empty((max(F_local, 0), N), dtype=torch.uint8, device=dev) if wf_rows else None
    if F_local > 0:
        x = plan._host_input(load_samples(sh.sample0, sh.n_samples))
        d_in = nat.DeviceArray.from_host(x, device)
        torch.cuda.current_stream(dev).synchronize()
        plan.stft(d_in, wf_rows=rows if wf_rows else False, welch=welch, maxhold=mh, vmin=vmin, vmax=vmax,
                  n_samples=sh.n_samples)
        plan.sync()
        d_in.free()
    total = allreduce_partials(welch, mh, F_local, group)
    gathered = gather_rows(rows, gather_to, group) if (wf_rows and gather_to is not None) else None
    return {"welch_acc": welch, "maxhold": mh, "n_frames": total, "rows": gathered, "local_frames": F_local,
            "shard": sh}


# ----------------------------------------------------------------------------- fused compute + reduction over peer memory
class PeerReduceTarget:
    """The reduction target of a sharded capture, living on rank ``dst`` and mapped into every rank:
    Welch numerator float64 [1, N], max-hold float32 [1, N] and (optionally) the whole uint8 waterfall
    [F, N].  Every rank's STFT kernel reduces / writes straight into it over NVLink
    (``peer_outputs=True``: system-scope atomics, plain remote stores for the rows), so the partial-PSD
    all-reduce and the row gather are not separate collectives any more -- the transfer overlaps the
    transform chunk by chunk.  ``torch.distributed`` only carries the 64-byte IPC handles and barriers."""

    def __init__(self, nfft: int, n_frames: int, rank: int, world: int, device: int, dst: int = 0, want_rows: bool = True,
                 group=None):
        from . import _native as nat
        dist = _dist()
        self.rank, self.world, self.dst, self.device, self.group = rank, world, dst, device, group
        shapes = {"welch": ((1, nfft), np.float64), "maxhold": ((1, nfft), np.float32)}
        if want_rows:
            shapes["rows"] = ((max(n_frames, 1), nfft), np.uint8)
        self.buffers = {}
        handles = [None]
        if rank == dst:
            for k, (shape, dt) in shapes.items():
                self.buffers[k] = nat.PeerBuffer(shape, dt, device)
            handles = [{k: b.handle for k, b in self.buffers.items()}]
        if world > 1:
            dist.broadcast_object_list(handles, src=dst, group=group)
        if rank != dst:
            for k, (shape, dt) in shapes.items():
                self.buffers[k] = nat.PeerBuffer.open(handles[0][k], shape, dt, device)
        self.welch = self.buffers["welch"].view()
        self.maxhold = self.buffers["maxhold"].view()
        self.rows = self.buffers["rows"].view() if want_rows else None

    def zero(self) -> None:
        """Owner clears the accumulators (rows are fully overwritten); call before a barrier."""
        if self.rank == self.dst:
            self.buffers["welch"].array.zero_()
            self.buffers["maxhold"].array.zero_()

    def close(self) -> None:
        for b in self.buffers.values():
            b.close()
        self.buffers = {}


def fused_capture_step(plan, d_in, shard: CaptureShard, target: PeerReduceTarget, vmin=-100.0, vmax=0.0) -> int:
    """One rank's share of a sharded capture with the reduction fused into the kernel: frames
    [shard.f0, shard.f1) of the device-resident slice ``d_in`` are transformed and their Welch / max-hold
    partials and uint8 rows land in ``target`` (possibly on another GPU).  Returns the local frame count;
    the caller synchronises the plan and barriers before the owner reads the result."""
    F_local = shard.f1 - shard.f0
    if F_local <= 0:
        return 0
    rows = target.rows.rows(shard.f0, shard.f1) if target.rows is not None else False
    # the owner writes its own rows directly; every other rank stages them and lets the copy engine push them
    mode = 1 if target.rank == target.dst else 3
    plan.stft(d_in, wf_rows=rows, welch=target.welch, maxhold=target.maxhold, vmin=vmin, vmax=vmax, accumulate=True,
              n_samples=shard.n_samples, peer_outputs=mode)
    return F_local
