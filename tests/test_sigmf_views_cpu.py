"""CPU tests of the host logic around the hot path: SigMF file conventions and the
dashboard view helpers that need no device (8(f-2))."""
import io
import json
import zipfile

import numpy as np
import pytest

from oracle import spectral_ref as sref
from sdr_iq_visualizer_b200 import sigmf_io, views


def test_sigmf_roundtrip_cf32(tmp_path):
    x = sref.synth_iq(5000, seed=1).astype(np.complex64)
    base = sigmf_io.write_recording(tmp_path / "c1", x, 1e6, 2.4e9, hw="PlutoSDR @ ip:test")
    for p in (base, base + ".sigmf-data", base + ".sigmf-meta"):      # process_sigmf_data.py:35-44
        rec = sigmf_io.fromfile(p)
        assert rec.datatype == "cf32_le" and len(rec) == 5000
        assert rec.sample_rate == 1e6 and rec.center_freq == 2.4e9
        assert np.array_equal(rec.read_samples(), x) and rec.read_samples().dtype == np.complex64
    meta = json.load(open(base + ".sigmf-meta"))
    # the layout the project itself emits (callbacks.py:285-304)
    assert meta["global"]["core:datatype"] == "cf32_le" and meta["global"]["core:sample_rate"] == 1000000
    assert meta["captures"][0]["core:frequency"] == 2400000000 and meta["captures"][0]["core:sample_start"] == 0
    assert meta["annotations"] == [] and meta["captures"][0]["core:datetime"].endswith("Z")
    info = rec.frequency_info()                                        # :125-145
    assert info["start_frequency"] == 2.4e9 - 0.5e6 and info["end_frequency"] == 2.4e9 + 0.5e6 and info["bandwidth"] == 1e6
    assert rec.get_global_field("core:sample_rate") == 1000000 and rec.get_annotations() == []


def test_sigmf_ci16_scaling_and_raw(tmp_path):
    iq = sref.to_ci16(sref.synth_iq(3001, seed=2))
    base = sigmf_io.write_recording(tmp_path / "pluto", iq, 61.44e6, 2.4e9, datatype="ci16_le")
    rec = sigmf_io.fromfile(base)
    assert rec.datatype == "ci16_le" and len(rec) == 3001 and rec.in_scale == 2.0 ** -15 and rec.bytes_per_sample == 4
    assert np.array_equal(rec.raw_slice(), iq)
    got = rec.read_samples(10, 100)                                    # sigmf-python: int16 * 2^-15 -> complex64
    want = (iq[20:220].astype(np.float32) / 32768.0).view(np.complex64)
    assert got.dtype == np.complex64 and np.array_equal(got, want)
    # complex input to a ci16 file is quantised with the inverse scale
    base2 = sigmf_io.write_recording(tmp_path / "q", want.astype(np.complex128), 1e6, 0, datatype="ci16_le")
    assert np.array_equal(sigmf_io.fromfile(base2).raw_slice(), iq[20:220])


def test_sigmf_errors_and_empty(tmp_path):
    base = sigmf_io.write_recording(tmp_path / "e", np.zeros(0, np.complex64), 1e6, 0)
    assert len(sigmf_io.fromfile(base)) == 0
    meta = json.load(open(base + ".sigmf-meta"))
    meta["global"]["core:datatype"] = "ru8"
    json.dump(meta, open(base + ".sigmf-meta", "w"))
    with pytest.raises(ValueError):
        sigmf_io.fromfile(base)
    with pytest.raises(FileNotFoundError):
        sigmf_io.fromfile(tmp_path / "missing")


@pytest.mark.parametrize("L,nfft,hop,maxc", [(10000, 1024, 512, 3000), (4096, 1024, 1024, 1024), (5000, 256, 64, 700),
                                             (1023, 1024, 512, 4096), (70000, 4096, 1024, 1 << 14)])
def test_frame_chunks_cover_every_frame_once(tmp_path, L, nfft, hop, maxc):
    x = (np.arange(L) + 1j * np.arange(L)).astype(np.complex64)
    rec = sigmf_io.fromfile(sigmf_io.write_recording(tmp_path / "r", x, 1e6, 0))
    F = sref.frame_count(L, nfft, hop)
    nxt = 0
    for f0, nf, raw in rec.frame_chunks(nfft, hop, maxc):
        assert f0 == nxt and nf >= 1
        assert raw.size // 2 == (nf - 1) * hop + nfft                 # halo of nfft - hop samples
        assert raw[0] == f0 * hop and raw[-1] == f0 * hop + (nf - 1) * hop + nfft - 1   # frame slicing is bit-exact
        nxt += nf
    assert nxt == F


def test_recording_zip_layout():
    x = sref.synth_iq(256, seed=3)
    blob = sigmf_io.recording_zip(x, 1e6, 2.4e9, "sdr_sample_20250101_000000", hw="PlutoSDR @ ip:x")
    z = zipfile.ZipFile(io.BytesIO(blob))
    assert sorted(z.namelist()) == ["README.txt", "sdr_sample_20250101_000000.sigmf-data", "sdr_sample_20250101_000000.sigmf-meta"]
    assert np.array_equal(np.frombuffer(z.read("sdr_sample_20250101_000000.sigmf-data"), np.complex64), x.astype(np.complex64))
    assert json.loads(z.read("sdr_sample_20250101_000000.sigmf-meta"))["global"]["core:hw"] == "PlutoSDR @ ip:x"


def test_waterfall_block_ring_and_quantisation():
    wf = views.WaterfallBlock(8, depth=4, vmin=-100.0, vmax=0.0)
    rows = [np.full(8, -100.0 + 10 * i) for i in range(7)]
    for i, r in enumerate(rows):
        wf.push_db(r)
        assert len(wf) == min(i + 1, 4)
    got = wf.rows()
    want = sref.waterfall_u8(np.array(rows[-4:]), -100.0, 0.0)          # oldest first, like np.array(deque) (callbacks.py:182)
    assert got.dtype == np.uint8 and np.array_equal(got, want)
    wf.push_db(np.array([np.nan, np.inf, -np.inf, -240.0, 0.0, 1e9, -50.0, -50.0 + 100 / 256]))
    assert list(wf.rows()[-1]) == [0, 255, 0, 0, 255, 255, 128, 129]
    pay = wf.heatmap_payload(np.arange(8.0))
    assert pay["z"].shape == (4, 8) and pay["zmin"] == 0 and pay["zmax"] == 255 and pay["y"] == [0, 1, 2, 3]
    assert wf.rgb().shape == (4, 8, 3) and np.array_equal(wf.rgb()[0, 0], sref.viridis_lut()[got[1, 0]])
    wf.push_rows(np.arange(24, dtype=np.uint8).reshape(3, 8))
    assert np.array_equal(wf.rows()[-3:], np.arange(24, dtype=np.uint8).reshape(3, 8))


def test_streamer_peek_does_not_consume():
    import sys
    from unittest.mock import MagicMock
    sys.modules.setdefault("adi", MagicMock())
    from app.sdr.streamer import SDRDataStreamer
    s = SDRDataStreamer()
    assert s.peek_latest() is None
    s._push({"i": 1}); s._push({"i": 2})
    assert s.peek_latest() == {"i": 2} and s.data_queue.qsize() == 2
    assert s.get_latest_data() == {"i": 1} and s.peek_latest() == {"i": 2}
    assert views.classify_tool_text(None) == {"stats": "No SDR data available yet. Please start streaming.", "include_graph": None}
