"""B200-native spectral hot path of sdr-iq-visualizer (windowed STFT -> PSD -> waterfall,
classifier features, time-domain views).  CUDA only: importing is cheap, but every compute call
needs csrc/libspx.so and a CUDA device, and raises ``SpectralError`` otherwise."""
from ._native import SpectralError, DeviceArray, pinned_empty, device_count, device_info  # noqa: F401

__all__ = ["SpectralError", "DeviceArray", "pinned_empty", "device_count", "device_info"]
