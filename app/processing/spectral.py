"""`app.processing.spectral`: the new spectral module of the drop-in."""
from sdr_iq_visualizer_b200.spectral import *  # noqa: F401,F403
from sdr_iq_visualizer_b200.spectral import (  # noqa: F401
    SpectralPlan, StftResult, freq_axis, get_plan, stream_frame, viridis_lut, waterfall, welch_psd)
