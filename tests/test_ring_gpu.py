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


def test_streamer_block_mode_matches_one_shot():
    """rx buffers -> pinned ring -> blocks of frames_per_block frames with u8 rows, Welch, running
    max-hold and on-device classifier measurements; identical to one pass over the concatenated stream."""
    import sys
    from unittest.mock import MagicMock
    sys.modules.setdefault("adi", MagicMock())
    from app.processing import classifier as clf
    from app.sdr.streamer import SDRDataStreamer
    from sdr_iq_visualizer_b200 import features, spectral as sp
    n, hop, fpb, nbuf = 4096, 1024, 64, 40
    raw = sref.to_ci16(sref.synth_iq(nbuf * n, seed=33))
    x = raw[0::2].astype(np.float64) + 1j * raw[1::2].astype(np.float64)      # what pyadi-iio's rx() returns
    s = SDRDataStreamer(sample_rate=61_440_000, center_freq=2_400_000_000)
    s.enable_block_mode(nfft=n, overlap=0.75, window="hann", frames_per_block=fpb, n_slots=3, vmin=20.0, vmax=130.0)
    blocks = []
    for b in range(nbuf):
        d = s.process_buffer(x[b * n:(b + 1) * n])
        assert set(d) >= {"time", "samples", "freqs", "power_db", "sample_rate", "center_freq"}
        blk = s.get_latest_block()
        if blk is not None and (not blocks or blk["seq"] != blocks[-1]["seq"]):
            blocks.append(blk)
    s.flush_blocks()
    blk = s.get_latest_block()
    if blk["seq"] != blocks[-1]["seq"]:
        blocks.append(blk)
    st = s.get_status()
    assert st["ring"]["h2d_bytes"] == 4 * nbuf * n and st["blocks_done"] >= len(blocks)
    pl = sp.SpectralPlan(n, hop, "hann", sp.FMT_CI16)
    one = pl.stft(raw, wf_rows=True, welch=True, maxhold=True, vmin=20.0, vmax=130.0)
    # the published blocks are a subsequence of all blocks (only the newest is kept): check each against its frames
    for blk in blocks:
        f0, nf = blk["first_frame"], blk["n_frames"]
        np.testing.assert_array_equal(blk["wf_rows"], one.wf_rows[f0:f0 + nf])
        part = pl.stft(raw[2 * f0 * hop: 2 * ((f0 + nf - 1) * hop + n)], welch=True)
        np.testing.assert_allclose(blk["welch_acc"], part.welch_acc[0], rtol=1e-6)
        _, pdb = pl.welch_finalize(blk["welch_acc"], nf, 61.44e6)
        np.testing.assert_allclose(blk["pxx_db"], pdb, rtol=0, atol=1e-9)
        m = features.measure(blk["pxx_db"])
        for k in ("noise_floor_db", "snr_db", "first_20db", "last_20db", "flatness", "kurtosis", "peak_count"):
            assert blk["features"][k] == m[k], k
        clf._CLASS_HISTORY.clear(); clf._CONF_HISTORY.clear()
        want = clf.classify_signal_advanced(blk["freqs"], blk["pxx_db"])
        assert blk["classification"]["label"] is not None
        got_feats, want_feats = blk["classification"]["features"], want["features"]
        assert got_feats["snr_db"] == want_feats["snr_db"] and got_feats["bandwidth_hz_20db"] == want_feats["bandwidth_hz_20db"]
        assert abs(got_feats["peak_spacing_std_hz"] - want_feats["peak_spacing_std_hz"]) <= 1e-6 * max(1.0, want_feats["peak_spacing_std_hz"])
    last = blocks[-1]
    assert last["first_frame"] + last["n_frames"] == one.n_frames              # flush published the tail block
    np.testing.assert_array_equal(last["maxhold"], one.maxhold[0])              # running max-hold over every block
    clf._CLASS_HISTORY.clear(); clf._CONF_HISTORY.clear()
    s.disable_block_mode()
    pl.close()
