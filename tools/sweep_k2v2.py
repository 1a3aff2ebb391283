#!/usr/bin/env python3
"""Config-5 shape (cf32, N = 65536): K2v2 single persistent kernel (variant 0) vs the round-1 two-kernel path (variant 1)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import kernel_sweep as ks
from sdr_iq_visualizer_b200 import spectral as sp
for L in (1 << 26, 1 << 28):
    for v in (1, 0):
        ks.run_case(f"C5 shape cf32 N=65536 50% u8+acc L=2^{L.bit_length()-1}", 65536, 32768, "hann", sp.FMT_CF32, L, ["u8", "acc"], v, iters=5)
        ks.run_case(f"cf32 N=65536 hop=N u8 L=2^{L.bit_length()-1}", 65536, 65536, "hann", sp.FMT_CF32, L, ["u8"], v, iters=5)
        ks.run_case(f"cf32 N=65536 50% acc only L=2^{L.bit_length()-1}", 65536, 32768, "hann", sp.FMT_CF32, L, ["acc"], v, iters=5)
