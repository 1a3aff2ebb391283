"""CPU tests of the multi-GPU host logic: partitioning with hop halos, and the
all-reduce / gather plumbing on a world_size-2 gloo group, with per-rank partials computed by the
oracle standing in for the GPU kernels."""
import os
import socket

import numpy as np
import pytest

from oracle import spectral_ref as sref
from sdr_iq_visualizer_b200 import dist as sd


def test_stream_blocks_cover_everything():
    for n, w in [(64, 1), (64, 2), (64, 8), (10, 4), (3, 8), (0, 2)]:
        blocks = [sd.stream_block(n, r, w) for r in range(w)]
        assert blocks[0][0] == 0 and blocks[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(blocks, blocks[1:]))
        sizes = [b - a for a, b in blocks]
        assert max(sizes) - min(sizes) <= 1


@pytest.mark.parametrize("L,n,hop,world", [(2**30, 65536, 32768, 8), (2**24, 65536, 32768, 4), (100_000, 4096, 1024, 3),
                                           (5000, 4096, 1024, 4), (100, 4096, 1024, 2)])
def test_capture_shards_tile_frames_with_halos(L, n, hop, world):
    F = sref.frame_count(L, n, hop)
    shards = [sd.capture_shard(L, n, hop, r, world) for r in range(world)]
    assert shards[0].f0 == 0 and shards[-1].f1 == F
    for a, b in zip(shards, shards[1:]):
        assert a.f1 == b.f0
    for s in shards:
        if s.f1 > s.f0:
            assert s.sample0 == s.f0 * hop and s.n_samples == (s.f1 - s.f0 - 1) * hop + n
            assert s.sample0 + s.n_samples <= L
            assert sref.frame_count(s.n_samples, n, hop) == s.f1 - s.f0
    owners = [s for s in shards if s.f1 > s.f0]
    for a, b in zip(owners, owners[1:]):           # neighbours share exactly the (N - hop) halo
        assert a.sample0 + a.n_samples - b.sample0 == n - hop == a.halo
    if L == 2**30:
        assert F == 32767 and all(s.halo in (0, 32768) for s in shards)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, L, n, hop, out_dir):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    x = sref.synth_iq(L, seed=5).astype(np.complex64)          # every rank can read the "file"
    sh = sd.capture_shard(L, n, hop, rank, world)
    mine = x[sh.sample0: sh.sample0 + sh.n_samples]
    P = sref.stft_power_rows(mine, n, hop, "hann")              # oracle stands in for the GPU kernel
    welch = torch.from_numpy(P.sum(axis=0)[None, :].copy())
    mh = torch.from_numpy(P.max(axis=0)[None, :].astype(np.float32)) if P.shape[0] else torch.zeros((1, n), dtype=torch.float32)
    rows = torch.from_numpy(sref.waterfall_u8(sref.power_db10(P), -20.0, 80.0))
    total = sd.allreduce_partials(welch, mh, P.shape[0])
    g = sd.gather_rows(rows, dst=0)
    feats = sd.allgather_objects({"rank": rank, "frames": int(P.shape[0])})
    if rank == 0:
        np.savez(os.path.join(out_dir, "r0.npz"), welch=welch.numpy(), mh=mh.numpy(), total=total, rows=g.numpy(),
                 frames=[f["frames"] for f in feats])
    else:
        assert g is None
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_gloo_allreduce_and_gather_match_single_process(tmp_path, world):
    import torch.multiprocessing as mp
    L, n, hop = 40_000, 1024, 512
    mp.spawn(_worker, args=(world, _free_port(), L, n, hop, str(tmp_path)), nprocs=world, join=True)
    z = np.load(tmp_path / "r0.npz")
    x = sref.synth_iq(L, seed=5).astype(np.complex64)
    P = sref.stft_power_rows(x, n, hop, "hann")
    assert int(z["total"]) == P.shape[0] == sum(z["frames"])
    np.testing.assert_allclose(z["welch"][0], P.sum(axis=0), rtol=1e-12)
    np.testing.assert_array_equal(z["mh"][0], P.max(axis=0).astype(np.float32))
    np.testing.assert_array_equal(z["rows"], sref.waterfall_u8(sref.power_db10(P), -20.0, 80.0))  # frame order kept
