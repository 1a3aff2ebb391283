#!/usr/bin/env python3
"""hop >= N shapes of K1v2 with the L2 prefetch distance given by SPX_L2PF_DIST (0 = off): python tools/sweep_l2pf.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import kernel_sweep as ks
from sdr_iq_visualizer_b200 import spectral as sp
L = 61_440_000
for rep in range(2):
    ks.run_case("headline cf32 N=4096 hop=N f32", 4096, 4096, "hann", sp.FMT_CF32, L, ["db"], 0)
    ks.run_case("cf32 N=4096 hop=N u8", 4096, 4096, "hann", sp.FMT_CF32, L, ["u8"], 0)
    ks.run_case("ci16 N=4096 hop=N u8+acc", 4096, 4096, "hann", sp.FMT_CI16, L, ["u8", "acc"], 0)
    ks.run_case("cf32 N=2048 hop=N f32", 2048, 2048, "hann", sp.FMT_CF32, L, ["db"], 0)
    ks.run_case("cf32 N=1024 hop=N f32", 1024, 1024, "hann", sp.FMT_CF32, L, ["db"], 0)
