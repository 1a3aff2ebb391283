"""World-size-2 GPU tests of the multi-GPU paths (skipped on boxes with fewer than two GPUs; the host-side sharding logic
is covered on CPU by tests/test_dist_cpu.py with gloo):
  * the fused peer path: CUDA-IPC mapped accumulators / rows on rank 0, system-scope atomics and copy-engine pushes
    over NVLink, checked against the float64 checker inside tools/bench_sharded.py (--check);
  * the C-level NCCL entry points of include/spx.h (spx_nccl_init / spx_allreduce_welch / spx_gather_rows)."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _n_gpus():
    from sdr_iq_visualizer_b200 import _native as nat
    return nat.device_count()


def _torchrun(args, port):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port)] + args
    return subprocess.run(cmd, capture_output=True, text=True, cwd=ROOT, timeout=600)


@pytest.mark.parametrize("rows,collective", [("gather", "fused"), ("sharded", "fused"), ("gather", "nccl")])
def test_sharded_capture_world2_matches_checker(rows, collective):
    if _n_gpus() < 2:
        pytest.skip("needs two GPUs")
    res = _torchrun([os.path.join(ROOT, "tools", "bench_sharded.py"), "--config", "c5", "--log2-samples", "25", "--steps", "2",
                     "--warmup", "1", "--rows", rows, "--collective", collective, "--check"], 29611)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-3000:]
    line = json.loads([l for l in res.stdout.splitlines() if l.startswith("{")][-1])
    assert line["check"].startswith("ok") and line["n_gpus"] == 2 and line["frames"] == (2 ** 25 - 65536) // 32768 + 1


def test_streams_world2_matches_checker():
    if _n_gpus() < 2:
        pytest.skip("needs two GPUs")
    res = _torchrun([os.path.join(ROOT, "tools", "bench_sharded.py"), "--config", "c4", "--log2-samples", "20", "--streams", "6",
                     "--steps", "2", "--warmup", "1", "--check"], 29612)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-3000:]
    line = json.loads([l for l in res.stdout.splitlines() if l.startswith("{")][-1])
    assert line["check"].startswith("ok") and line["features_gathered"] == 6


_NCCL_WORKER = r'''
import ctypes as C, os, sys, time
import numpy as np
sys.path.insert(0, %(root)r)
from sdr_iq_visualizer_b200 import _native as nat
rank, idfile = int(sys.argv[1]), sys.argv[2]
lib = nat.lib()
try:
    import torch, glob
    cands = glob.glob(os.path.join(os.path.dirname(torch.__file__), "..", "nvidia", "nccl", "lib", "libnccl.so.2"))
    if cands:
        os.environ["SPX_NCCL_LIB"] = os.path.abspath(cands[0])
except Exception:
    pass
if rank == 0:
    uid = (C.c_ubyte * 128)()
    nat.check(lib.spx_nccl_unique_id(uid))
    with open(idfile + ".tmp", "wb") as fh:
        fh.write(bytes(uid))
    os.rename(idfile + ".tmp", idfile)
else:
    for _ in range(600):
        if os.path.exists(idfile):
            break
        time.sleep(0.05)
    uid = (C.c_ubyte * 128).from_buffer_copy(open(idfile, "rb").read())
comm = C.c_void_p()
nat.check(lib.spx_nccl_init(C.byref(comm), rank, rank, 2, uid))
N = 4096
w = nat.DeviceArray.from_host(np.full(N, 1.5 + rank, np.float64), rank)
m = nat.DeviceArray.from_host((np.arange(N) %% 7 + 10 * rank).astype(np.float32), rank)
cnt = C.c_int64(100 + rank)
nat.check(lib.spx_allreduce_welch(comm, w.ptr, m.ptr, N, C.byref(cnt), None))
nat.device_sync(rank)
assert cnt.value == 201, cnt.value
assert np.all(w.to_host() == 4.0)
assert np.array_equal(m.to_host(), (np.arange(N) %% 7 + 10).astype(np.float32))
rows_local = nat.DeviceArray.from_host(np.full((3 + rank, N), 7 + rank, np.uint8), rank)
per = (C.c_int64 * 2)(3 * N, 4 * N)
rows_all = nat.DeviceArray((7, N), np.uint8, rank, zero=True) if rank == 0 else None
nat.check(lib.spx_gather_rows(comm, rows_local.ptr, (3 + rank) * N, rows_all.ptr if rows_all else None, per, 0, None))
nat.device_sync(rank)
if rank == 0:
    got = rows_all.to_host()
    assert np.all(got[:3] == 7) and np.all(got[3:] == 8)
nat.check(lib.spx_nccl_destroy(comm))
print("nccl worker", rank, "ok")
'''


def test_c_level_nccl_entry_points_world2(tmp_path):
    if _n_gpus() < 2:
        pytest.skip("needs two GPUs")
    script = tmp_path / "worker.py"
    script.write_text(_NCCL_WORKER % {"root": ROOT})
    idfile = str(tmp_path / "nccl_id.bin")
    procs = [subprocess.Popen([sys.executable, str(script), str(r), idfile], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
             for r in range(2)]
    outs = [p.communicate(timeout=300) for p in procs]
    for p, (o, e) in zip(procs, outs):
        assert p.returncode == 0, o[-1000:] + e[-3000:]
        assert "ok" in o


def test_peer_reduce_and_push_rows_c_entry_points_single_gpu():
    """spx_peer_reduce / spx_peer_push_rows on one GPU (the owner's buffers are plain local memory there)."""
    import ctypes as C
    from sdr_iq_visualizer_b200 import _native as nat
    lib = nat.lib()
    N = 8192
    rng = np.random.default_rng(0)
    wl, ml = rng.random(N), rng.random(N).astype(np.float32)
    wo, mo = rng.random(N), rng.random(N).astype(np.float32)
    d_wl, d_ml = nat.DeviceArray.from_host(wl, 0), nat.DeviceArray.from_host(ml, 0)
    d_wo, d_mo = nat.DeviceArray.from_host(wo, 0), nat.DeviceArray.from_host(mo, 0)
    nat.check(lib.spx_peer_reduce(0, d_wl.ptr, d_ml.ptr, d_wo.ptr, d_mo.ptr, N, None))
    rows = rng.integers(0, 256, (5, N), dtype=np.uint8)
    d_r, d_dst = nat.DeviceArray.from_host(rows, 0), nat.DeviceArray((5, N), np.uint8, 0, zero=True)
    nat.check(lib.spx_peer_push_rows(0, d_dst.ptr, d_r.ptr, rows.nbytes, None))
    nat.device_sync(0)
    np.testing.assert_array_equal(d_wo.to_host(), wo + wl)
    np.testing.assert_array_equal(d_mo.to_host(), np.maximum(mo, ml))
    np.testing.assert_array_equal(d_dst.to_host(), rows)
