"""`app.sdr.streamer` of the reference, served by the CUDA-backed implementation.  `adi` is imported
by the implementation module at its top, so `sys.modules['adi'] = MagicMock()` before import still
works (reference tests/test_streamer.py:7-9)."""
from sdr_iq_visualizer_b200.streamer import SDRDataStreamer, sdr_streamer  # noqa: F401
from sdr_iq_visualizer_b200.streamer import adi, queue, threading, time, logging, np  # noqa: F401
