#!/usr/bin/env python3
"""A/B of N = 4096 kernel variants on the config-2 and headline shapes, interleaved (variant order repeated) so that box
drift cancels: python tools/sweep_k1_variants.py 0,24 [repeats]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import kernel_sweep as ks
from sdr_iq_visualizer_b200 import spectral as sp
variants = [int(v) for v in (sys.argv[1].split(",") if len(sys.argv) > 1 else "0,24".split(","))]
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
L = 61_440_000
for _ in range(reps):
    for v in variants:
        ks.run_case("C2 ci16 75% u8+acc", 4096, 1024, "hann", sp.FMT_CI16, L, ["u8", "acc"], v)
        ks.run_case("headline cf32 hop=N f32", 4096, 4096, "hann", sp.FMT_CF32, L, ["db"], v)
        ks.run_case("cf32 50% u8+acc", 4096, 2048, "hann", sp.FMT_CF32, L, ["u8", "acc"], v)
