#!/usr/bin/env python3
"""N = 65536 on int16 input: two-kernel path (variant 1) vs K2v2 (variant 0)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import kernel_sweep as ks
from sdr_iq_visualizer_b200 import spectral as sp
L = 1 << 28
for v in (1, 0, 1, 0):
    ks.run_case("ci16 N=65536 50% u8+acc L=2^28", 65536, 32768, "hann", sp.FMT_CI16, L, ["u8", "acc"], v, iters=5)
    ks.run_case("ci16 N=65536 hop=N u8 L=2^28", 65536, 65536, "hann", sp.FMT_CI16, L, ["u8"], v, iters=5)
