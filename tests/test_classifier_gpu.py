"""GPU tests of the classifier drop-in (kernel K3 + host rules) against the golden results produced
by the reference itself (tests/golden/classifier_cases.*) and against the numpy oracle."""
import numpy as np
import pytest

from oracle import classifier_ref as cref

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def clf():
    from app.processing import classifier  # the drop-in import path of the reference
    from sdr_iq_visualizer_b200 import _native
    assert _native.device_count() > 0
    return classifier


def _clear(clf):
    clf._CLASS_HISTORY.clear()
    clf._CONF_HISTORY.clear()


def test_golden_cases_full_result(clf, golden_classifier):
    """classify_signal_advanced / _simple return exactly what the reference returned."""
    z, res = golden_classifier
    for name, want in res["cases"].items():
        f, p = z[name + "_freqs"], z[name + "_power_db"]
        _clear(clf)
        got = clf.classify_signal_advanced(f, p)
        _clear(clf)
        assert got == want["advanced"], name
        assert clf.classify_signal_simple(f, p) == want["simple"], name
    assert clf.classify_signal_advanced(np.array([]), np.array([])) == res["empty_advanced"]
    assert clf.classify_signal_simple(np.array([]), np.array([])) == res["empty_simple"]


def test_golden_intermediates_bit_exact(clf, golden_classifier):
    z, res = golden_classifier
    for name, want in res["cases"].items():
        f, p = z[name + "_freqs"], z[name + "_power_db"]
        m = clf.measure(f, p)
        assert m["noise_floor_db"] == want["noise_floor_db"], name       # exact order statistics + numpy lerp
        assert m["adaptive_thr"] == want["adaptive_thr"], name
        assert m["peaks"] == want["peaks"], name                          # integer: bit-exact
        assert [m["bw3"], m["bw10"], m["bw20"]] == want["bw"], name       # integer edges -> exact Hz
        assert m["peak_spacing_std_hz"] == want["peak_spacing_std_hz"], name
        np.testing.assert_allclose(m["flatness"], want["flatness"], rtol=1e-12, atol=1e-300, err_msg=name)
        np.testing.assert_allclose(m["kurtosis"], want["kurtosis"], rtol=1e-12, atol=1e-12, err_msg=name)
        # private helpers under the reference's names
        assert clf._estimate_noise_floor(p) == want["noise_floor_db"]
        assert clf._find_peaks(p, want["adaptive_thr"], max(3, len(p) // 300)) == want["peaks"]
        assert [clf._occupied_bandwidth(f, p, d) for d in (3, 10, 20)] == want["bw"]


def test_temporal_smoothing_sequence(clf, golden_classifier):
    """module-global 12-deep history (classifier.py:125-139) reproduces the reference's sequence."""
    z, res = golden_classifier
    _clear(clf)
    for step in res["sequence"]:
        name = step["case"]
        r = clf.classify_signal_advanced(z[name + "_freqs"], z[name + "_power_db"])
        assert (r["label"], r["confidence"]) == (step["label"], step["confidence"]), step
        assert r["reasons"] == step["reasons"] and r["explanation"] == step["explanation"]
    _clear(clf)


@pytest.mark.parametrize("min_dist", [-3, 0, 1, 2, 3, 5])
def test_find_peaks_small_distances_match_reference_loop(clf, min_dist):
    """classifier.py:200-212 with min_distance_bins <= 2 keeps every strict local maximum (up to ~n/2 of them): the
    drop-in helper must return the complete list, not the kernel's default distance and not a truncated buffer."""
    rng = np.random.default_rng(7)
    for n in (3, 50, 1001, 4096):
        p = rng.normal(-80, 3, n)
        p[1::2] += 20.0            # every odd bin is a strict local maximum: the worst case for the list size
        assert clf._find_peaks(p, -1000.0, min_dist) == cref.greedy_peaks(cref.peak_candidates(p, -1000.0), min_dist), (n, min_dist)


def test_reference_test_classifier_runs_unchanged(clf):
    """The reference's own tests/test_classifier.py (vendored byte for byte as tests/golden/reference_test_classifier.py,
    /root/reference/tests/test_classifier.py:1-63) executed against `app.processing.classifier` of this repo."""
    import hashlib
    import importlib.util
    import os
    import unittest
    path = os.path.join(os.path.dirname(__file__), "golden", "reference_test_classifier.py")
    with open(path, "rb") as fh:
        assert hashlib.sha256(fh.read()).hexdigest() == "81b5422a44e9c1b337f15d66b1cca256576cb665007569d4318e57a9e3123678", "vendored reference test was edited"
    spec = importlib.util.spec_from_file_location("reference_test_classifier", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)           # imports app.processing.classifier -> the CUDA-backed drop-in
    import app.processing.classifier as served
    assert mod.classify_signal_advanced is served.classify_signal_advanced
    _clear(clf)
    suite = unittest.defaultTestLoader.loadTestsFromModule(mod)
    assert suite.countTestCases() == 5
    result = unittest.TextTestRunner(verbosity=0).run(suite)
    assert result.wasSuccessful(), (result.failures, result.errors)
    _clear(clf)


def test_reference_unit_tests_restated(clf):
    """The assertions of the reference's tests/test_classifier.py:7-60, on the drop-in."""
    assert clf.classify_signal_simple(np.array([]), np.array([])) == "No Data"
    p = np.zeros(100); p[50] = 50
    assert clf.classify_signal_simple(np.linspace(0, 10e6, 100), p) == "Narrowband"
    assert clf.classify_signal_simple(np.linspace(0, 20e6, 200), np.full(200, 50.0)) == "Wideband"
    assert clf.classify_signal_advanced(np.array([]), np.array([]))["label"] == "No Data"
    freqs = np.linspace(100e6, 101e6, 1024)
    p = np.random.normal(-80, 1, 1024)
    p[512] = -20; p[511] = -30; p[513] = -30
    _clear(clf)
    r = clf.classify_signal_advanced(freqs, p)
    assert {"label", "confidence", "features"} <= set(r)
    assert r["label"] in ["CW Carrier", "Narrowband (voice)", "Unknown", "Low SNR / Noise"]
    _clear(clf)


@pytest.mark.parametrize("n", [1, 2, 3, 5, 100, 299, 300, 1023, 1024, 4096, 5000, 65536, 100000])
def test_random_spectra_vs_oracle(clf, n):
    rng = np.random.default_rng(n)
    for trial in range(3):
        p = rng.normal(-80, 3, n)
        p[rng.integers(0, n, size=max(1, n // 50))] += rng.uniform(5, 50)
        if trial == 1:
            p = np.round(p, 1)       # many exact ties: strict > / >= behaviour, duplicate order statistics
        if trial == 2:
            p = np.round(p)          # heavy duplication
        f = np.linspace(1e9, 1.02e9, n)
        want = cref.features(f, p)
        m = clf.measure(f, p)
        assert m["noise_floor_db"] == want["noise_floor_db"]
        assert m["peak_db"] == want["peak_db"] and m["argmax"] == want["argmax"]
        assert m["adaptive_thr"] == want["adaptive_thr"]
        assert m["min_distance_bins"] == want["min_distance_bins"]
        assert (m["first_3db"], m["last_3db"]) == want["edges"][3]
        assert (m["first_10db"], m["last_10db"]) == want["edges"][10]
        assert (m["first_20db"], m["last_20db"]) == want["edges"][20]
        assert (m["simple_first"], m["simple_last"]) == want["simple_edges"]
        assert m["n_candidates"] == want["n_candidates"]
        assert m["peaks"] == want["peaks"] and m["peak_count"] == want["peak_count"]
        assert m["peak_spacing_std_hz"] == want["peak_spacing_std_hz"]
        np.testing.assert_allclose(m["flatness"], want["flatness"], rtol=1e-11, atol=1e-300)
        np.testing.assert_allclose(m["kurtosis"], want["kurtosis"], rtol=1e-11, atol=1e-11)


def test_float32_and_batch(clf):
    from sdr_iq_visualizer_b200 import features
    rng = np.random.default_rng(3)
    rows = rng.normal(-70, 4, (7, 2048)).astype(np.float32)
    rows[:, 1000] += 40
    got = features.measure_batch(rows)
    f = np.arange(2048.0)
    for b in range(7):
        want = cref.features(f, rows[b].astype(np.float64))
        assert got[b]["noise_floor_db"] == want["noise_floor_db"]
        assert got[b]["peaks"] == want["peaks"]
        assert (got[b]["first_20db"], got[b]["last_20db"]) == want["edges"][20]
    from sdr_iq_visualizer_b200 import _native as nat
    d = nat.DeviceArray.from_host(rows)
    got_d = features.measure_batch(d)
    assert [g["peaks"] for g in got_d] == [g["peaks"] for g in got]
    assert [g["noise_floor_db"] for g in got_d] == [g["noise_floor_db"] for g in got]


def test_feature_queue_matches_synchronous_call(clf):
    """Enqueue-only measurement (device spectra -> pinned result structs, no host sync inside) == measure_batch."""
    from sdr_iq_visualizer_b200 import _native as nat, features
    rng = np.random.default_rng(12)
    p = rng.normal(-80, 3, (5, 4096))
    p[:, 1000] += 40
    want = features.measure_batch(p, want_peaks=False)
    dp = nat.DeviceArray.from_host(p)
    fq = features.FeatureQueue(batch=5, slots=3)
    slots = [fq.enqueue(dp, 4096) for _ in range(4)]          # wraps around the slot ring
    nat.device_sync(0)
    assert slots == [0, 1, 2, 0]
    for s in (0, 1, 2):
        got = fq.results(s)
        for g, w in zip(got, want):
            for k in ("noise_floor_db", "snr_db", "first_20db", "last_20db", "flatness", "kurtosis", "peak_count", "argmax"):
                assert g[k] == w[k], k
