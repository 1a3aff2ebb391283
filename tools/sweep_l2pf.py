#!/usr/bin/env python3
"""hop >= N shapes of K1v2.  profiles/r02_l2_prefetch_sweep.txt was produced with a run-time knob (SPX_L2PF_DIST = frames ahead,
0 = off) that has since been replaced by the compile-time switch TUNE_L2PF (distance 1, cf32 N = 4096 only: as a run-time branch
it cost the overlapped int16 kernel 1.4 %); today this script just times the shapes with the default kernels."""
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
