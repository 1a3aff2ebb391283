"""`app.processing.classifier` of the reference, served by the CUDA-backed implementation."""
from sdr_iq_visualizer_b200.classifier import *  # noqa: F401,F403
from sdr_iq_visualizer_b200.classifier import (  # noqa: F401
    _CLASS_HISTORY, _CONF_HISTORY, _estimate_noise_floor, _estimate_snr, _find_peaks, _occupied_bandwidth,
    _peak_spacing_std, _spectral_flatness, _spectral_kurtosis, classify_signal_advanced, classify_signal_simple,
    measure)
