"""GPU tests of the pinned ring ingest: continuity across slots, counters, back-pressure."""
import numpy as np
import pytest

from oracle import spectral_ref as sref
from tests import parity

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("chunks", [[1 << 16] * 6, [50_000, 65_536, 4096, 1000, 65_000, 7, 60_001]])
def test_ring_matches_one_shot(chunks):
    from sdr_iq_visualizer_b200 import spectral as sp, _native as nat
    from sdr_iq_visualizer_b200.ring import StreamRing
    n, hop = 4096, 1024
    L = sum(chunks)
    raw = sref.to_ci16(sref.synth_iq(L, seed=21))
    pl = sp.SpectralPlan(n, hop, "hann", sp.FMT_CI16)
    one = pl.stft(raw, wf_rows=True, db_rows=True, welch=True, maxhold=True, vmin=20.0, vmax=130.0)
    ring = StreamRing(pl, n_slots=3, slot_samples=1 << 16, wf_rows=True, db_rows=True, welch=True, maxhold=True,
                      vmin=20.0, vmax=130.0)
    rows, dbs, welch, mh, frames = [], [], np.zeros(n), np.zeros(n, np.float32), 0
    pos, pending = 0, 0
    for c in chunks:
        ring.push(raw[2 * pos: 2 * (pos + c)])
        pos += c
        pending += 1
        if pending == 3:                       # ring full: drain one
            with pytest.raises(nat.SpectralError):
                ring.acquire()
            b = ring.collect()
            assert b["first_frame"] == frames
            rows.append(b["wf_rows"].copy()); dbs.append(b["db_rows"].copy())
            welch += b["welch_acc"]; mh = np.maximum(mh, b["maxhold"]); frames += b["n_frames"]
            ring.release(); pending -= 1
    while pending:
        b = ring.collect()
        rows.append(b["wf_rows"].copy()); dbs.append(b["db_rows"].copy())
        welch += b["welch_acc"]; mh = np.maximum(mh, b["maxhold"]); frames += b["n_frames"]
        ring.release(); pending -= 1
    st = ring.stats()
    assert frames == one.n_frames == st["frames"] and st["samples"] == L
    assert st["h2d_bytes"] == 4 * L and st["in_flight"] == 0           # every sample crosses PCIe exactly once
    np.testing.assert_array_equal(np.concatenate(rows), one.wf_rows)     # same kernel, same inputs: identical
    np.testing.assert_array_equal(np.concatenate(dbs), one.db_rows)
    np.testing.assert_array_equal(mh, one.maxhold[0])
    np.testing.assert_allclose(welch, one.welch_acc[0], rtol=1e-6)
    from tests.test_stft_gpu import oracle_rows
    X = oracle_rows(raw, n, hop, "hann", fmt=1)
    parity.check_power(welch, (X.real**2 + X.imag**2).sum(axis=0), what="ring welch")
    ring.close(); pl.close()
