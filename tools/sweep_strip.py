#!/usr/bin/env python3
"""Strip staging (variant 23) vs the default K1v2 (variant 22) on the overlapped N = 4096 shapes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import kernel_sweep as ks
from sdr_iq_visualizer_b200 import spectral as sp
L = 61_440_000
for v in (22, 23, 22, 23):
    ks.run_case("C2 ci16 75% u8+acc", 4096, 1024, "hann", sp.FMT_CI16, L, ["u8", "acc"], v)
    ks.run_case("cf32 75% u8", 4096, 1024, "hann", sp.FMT_CF32, L, ["u8"], v)
    ks.run_case("cf32 50% f32", 4096, 2048, "hann", sp.FMT_CF32, L, ["db"], v)
