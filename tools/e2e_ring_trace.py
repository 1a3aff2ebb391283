#!/usr/bin/env python3
"""Timeline of the ingest ring on config 2 (SPX_RING_TRACE=1: one stderr line per released slot, CUDA-event times in us):
SPX_RING_TRACE=1 python tools/e2e_ring_trace.py 2> trace.txt"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sdr_iq_visualizer_b200 import _native as nat, ring as ringmod, spectral as sp

N, HOP, FS, SLOT = 4096, 1024, 61.44e6, 1 << 22
pl = sp.SpectralPlan(N, HOP, "hann", sp.FMT_CI16)
rg = ringmod.StreamRing(pl, n_slots=4, slot_samples=SLOT, wf_rows=True, welch=True, maxhold=True, vmin=-20.0, vmax=100.0,
                        features=True, sample_rate=FS)
rng = np.random.default_rng(0)
pending = 0
for k in range(40):
    b = rg.acquire()
    if k < 4:
        b[:] = rng.integers(-2000, 2000, b.size, dtype=np.int16)
    rg.commit(SLOT); pending += 1
    if pending >= 3:
        rg.collect(); rg.release(); pending -= 1
while pending:
    rg.collect(); rg.release(); pending -= 1
rg.close(); pl.close()
