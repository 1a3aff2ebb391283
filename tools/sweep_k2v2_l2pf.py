#!/usr/bin/env python3
"""K2v2 on the config-5 shapes with / without the L2 tensor prefetch (SPX_BIG2_L2PF=0/1 in the environment)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import kernel_sweep as ks
from sdr_iq_visualizer_b200 import spectral as sp
L = 1 << 28
for rep in range(2):
    ks.run_case("C5 shape cf32 N=65536 50% u8+acc L=2^28", 65536, 32768, "hann", sp.FMT_CF32, L, ["u8", "acc"], 0, iters=5)
    ks.run_case("cf32 N=65536 hop=N u8 L=2^28", 65536, 65536, "hann", sp.FMT_CF32, L, ["u8"], 0, iters=5)
    ks.run_case("cf32 N=65536 50% acc only L=2^28", 65536, 32768, "hann", sp.FMT_CF32, L, ["acc"], 0, iters=5)
