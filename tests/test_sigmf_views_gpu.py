"""GPU tests of the rows SURVEY.md 8(f) marks "next": SigMF recordings through the fused STFT path
(BASELINE config 1 shape, file-backed and chunked) and the dashboard / chatbot view helpers."""
import numpy as np
import pytest

from oracle import classifier_ref as cref
from oracle import spectral_ref as sref
from tests import parity

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mods():
    from sdr_iq_visualizer_b200 import _native, sigmf_io, views
    assert _native.device_count() > 0
    return sigmf_io, views


def test_config1_sigmf_file_psd_and_waterfall(mods, tmp_path):
    """C1: 2^20 cf32 samples as a SigMF file, 1024-pt Hann, 50 % overlap -> Welch PSD + u8 waterfall; processed
    in chunks with hop halos -- identical frames to one pass."""
    sigmf_io, _ = mods
    L, n, hop, fs, fc = 1 << 20, 1024, 512, 1e6, 2.4e9
    x = sref.synth_iq(L, seed=1).astype(np.complex64)
    base = sigmf_io.write_recording(tmp_path / "c1", x, fs, fc)
    out = sigmf_io.process_recording(base, n, hop, "hann", waterfall=True, vmin=-60.0, vmax=70.0, hist_r=4.0,
                                     max_chunk_samples=150_000)
    assert out.n_frames == 2047 and out.wf_rows.shape == (2047, 1024)
    f_ref, p_ref = sref.welch_psd(x, n, hop, "hann", fs, fc)
    assert np.array_equal(out.freqs, f_ref) and out.freqs[0] == fc - fs / 2
    parity.check_power(out.pxx, p_ref, what="C1 file welch")
    X = sref.shift_bins(sref.stft(sref.as_complex128(x), n, hop, "hann"))
    P = X.real**2 + X.imag**2
    parity.check_power(out.maxhold, P.max(axis=0), what="C1 file maxhold")
    parity.check_u8(out.wf_rows, sref.amplitude_db(X), -60.0, 70.0, what="C1 file u8")
    assert np.array_equal(out.hist, sref.iq_hist2d(x, 4.0, 256))      # counts bit-exact, every sample once
    assert out.h2d_bytes >= L * 8
    one = sigmf_io.process_recording(base, n, hop, "hann", waterfall=True, vmin=-60.0, vmax=70.0)
    assert np.array_equal(one.wf_rows, out.wf_rows)                    # chunking does not change a single index


def test_ci16_recording_scaled_in_kernel(mods, tmp_path):
    sigmf_io, _ = mods
    iq = sref.to_ci16(sref.synth_iq(200_000, seed=7))
    base = sigmf_io.write_recording(tmp_path / "p", iq, 61.44e6, 2.4e9, datatype="ci16_le")
    pxx, freqs = sigmf_io.psd(base, NFFT=4096, noverlap=3072)
    xs = sref.as_complex128(iq, sref.FMT_CI16, 2.0 ** -15)             # what sigmf-python's read_samples returns
    f_ref, p_ref = sref.welch_psd(xs, 4096, 1024, "hann", 61.44e6, 2.4e9)
    assert np.array_equal(freqs, f_ref)
    parity.check_power(pxx, p_ref, what="ci16 file welch")
    # the reference's 10 000-sample cap (process_sigmf_data.py:148,154) and mlab's default noverlap = 0
    pxx_c, _ = sigmf_io.psd(base, NFFT=1024, max_samples=10_000)
    _, p_c = sref.welch_psd(xs[:10_000], 1024, 1024, "hann", 61.44e6, 2.4e9)
    parity.check_power(pxx_c, p_c, what="capped welch")
    # shorter than one frame: zero-padded to a single frame as mlab does
    short = sigmf_io.write_recording(tmp_path / "s", sref.synth_iq(600, seed=8).astype(np.complex64), 1e6, 0.0)
    pxx_s, _ = sigmf_io.psd(short, NFFT=1024)
    xz = np.zeros(1024, np.complex128); xz[:600] = sref.synth_iq(600, seed=8).astype(np.complex64)
    parity.check_power(pxx_s, sref.welch_psd(xz, 1024, 1024, "hann", 1e6, 0.0)[1], what="short welch", rel_tol=2e-4)


def test_dashboard_tick_views(mods, golden_classifier):
    _, views = mods
    from app.processing import classifier as clf
    fs, fc, n = 61.44e6, 2.4e9, 4096
    rng = np.random.default_rng(11)
    samples = (rng.integers(-2047, 2048, n) + 1j * rng.integers(-2047, 2048, n)).astype(np.complex128)
    freqs, power_db = sref.stream_frame(samples, fs, fc)
    wf = views.WaterfallBlock(n, vmin=0.0, vmax=130.0)
    clf._CLASS_HISTORY.clear(); clf._CONF_HISTORY.clear()
    t = views.dashboard_tick({"time": 0.0, "samples": samples, "freqs": freqs, "power_db": power_db,
                              "sample_rate": fs, "center_freq": fc}, wf)
    # constellation: exact density of all samples (np.histogram2d semantics), displayed transposed
    c = t["constellation"]
    assert c["r"] == 2048.0 and np.array_equal(c["counts"], sref.iq_hist2d(samples, 2048.0, 256)) and c["counts"].sum() == n
    assert np.array_equal(c["z"], c["counts"].T)
    # peak markers: the classifier's rule with the dashboard's spacing
    f = cref.features(freqs, power_db)
    want = cref.greedy_peaks(cref.peak_candidates(power_db, f["adaptive_thr"]), max(5, n // 200))
    assert list(views.peak_markers(power_db)) == want and np.array_equal(t["peak_db"], power_db[want])
    assert np.array_equal(t["waterfall"]["z"], sref.waterfall_u8(power_db[None, :], 0.0, 130.0))
    assert t["time_ms"][1] == 1000.0 / fs and np.array_equal(t["i"], samples.real)
    assert views.dashboard_tick(None, wf) is None
    # classifier line: same text the reference's callback builds (callbacks.py:224-238) from the reference's own result
    z, res = golden_classifier
    for name, want_res in list(res["cases"].items())[:6]:
        clf._CLASS_HISTORY.clear(); clf._CONF_HISTORY.clear()
        r = want_res["advanced"]; ft = r["features"]
        line = (f"Detected: {r['label']} (conf {r['confidence']:.2f}) — OBW20={ft['bandwidth_hz_20db'] / 1e6:.2f}MHz "
                f"SNR={ft['snr_db']:.1f}dB | Flat {ft['spectral_flatness']:.2f} | Kurt {ft['spectral_kurtosis']:.2f} | "
                f"Peaks {ft['peak_count']}\n{r['explanation']}")
        assert views.classification_text(z[name + "_freqs"], z[name + "_power_db"]) == line, name
        clf._CLASS_HISTORY.clear(); clf._CONF_HISTORY.clear()
        tool = views.classify_tool_text({"freqs": z[name + "_freqs"], "power_db": z[name + "_power_db"]})
        reasons = r.get("reasons", [])
        assert tool["include_graph"] == "fd" and tool["stats"] == (
            f"Classification: {r['label']} (conf {r['confidence']:.2f})\nOBW20={ft['bandwidth_hz_20db'] / 1e6:.2f} MHz, "
            f"SNR={ft['snr_db']:.1f} dB" + ("\n- " + "\n- ".join(reasons) if reasons else "")), name
    clf._CLASS_HISTORY.clear(); clf._CONF_HISTORY.clear()


def test_streamer_emits_kernel_quantised_waterfall_row(mods):
    """With waterfall_range set, the frame dict of streamer.py:123-130 also carries the uint8 row produced by the same
    launch, and the dashboard tick takes it as is."""
    import sys
    from unittest.mock import MagicMock
    sys.modules.setdefault("adi", MagicMock())
    _, views = mods
    from app.sdr.streamer import SDRDataStreamer
    rng = np.random.default_rng(5)
    samples = (rng.integers(-2047, 2048, 4096) + 1j * rng.integers(-2047, 2048, 4096)).astype(np.complex128)
    s = SDRDataStreamer(sample_rate=61_440_000)
    plain = s.process_buffer(samples)
    assert set(plain) == {"time", "samples", "freqs", "power_db", "sample_rate", "center_freq"}    # the reference's dict, unchanged
    s.waterfall_range = (0.0, 130.0)
    d = s.process_buffer(samples)
    assert np.array_equal(d["power_db"], plain["power_db"]) and d["wf_row"].dtype == np.uint8
    X = np.fft.fftshift(np.fft.fft(samples))
    parity.check_u8(d["wf_row"][None, :], sref.amplitude_db(X)[None, :], 0.0, 130.0, what="stream wf_row")
    wf = views.WaterfallBlock(4096, vmin=0.0, vmax=130.0)
    t = views.dashboard_tick(d, wf)
    assert np.array_equal(t["waterfall"]["z"][0], d["wf_row"])
