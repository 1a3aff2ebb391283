#!/usr/bin/env python3
"""Where does the ring's end-to-end time go?  config 2 through StreamRing with different ring depths; reports ms per
second-of-stream and the share of host time spent blocked in collect(): python tools/e2e_ring_probe.py"""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sdr_iq_visualizer_b200 import _native as nat, ring as ringmod, spectral as sp

L, N, HOP, FS = 61_440_000, 4096, 1024, 61.44e6
for n_slots, keep, slot_log2 in ((4, 3, 22), (4, 2, 22), (6, 5, 22), (8, 7, 22), (4, 3, 23), (8, 7, 21)):
    pl = sp.SpectralPlan(N, HOP, "hann", sp.FMT_CI16)
    SLOT = 1 << slot_log2
    n_full, tail = divmod(L, SLOT)
    sizes = [SLOT] * n_full + ([tail] if tail else [])
    rg = ringmod.StreamRing(pl, n_slots=n_slots, slot_samples=SLOT, wf_rows=True, welch=True, maxhold=True, vmin=-20.0, vmax=100.0,
                            features=True, sample_rate=FS)
    rng = np.random.default_rng(0)
    for _ in range(n_slots):                       # fill every slot once
        b = rg.acquire(); b[:] = rng.integers(-2000, 2000, b.size, dtype=np.int16); rg.commit(SLOT); rg.collect(); rg.release()
    state = {"pending": 0, "blocked": 0.0}

    def one_pass(drain):
        for n in sizes:
            rg.acquire(); rg.commit(n); state["pending"] += 1
            if state["pending"] >= keep:
                t = time.perf_counter(); rg.collect(); state["blocked"] += time.perf_counter() - t
                rg.release(); state["pending"] -= 1
        while drain and state["pending"]:
            rg.collect(); rg.release(); state["pending"] -= 1

    one_pass(True)
    nat.device_sync(0)
    state["blocked"] = 0.0
    steps = 8
    t0 = time.perf_counter()
    for i in range(steps):
        one_pass(i == steps - 1)
    nat.device_sync(0)
    dt = time.perf_counter() - t0
    print(json.dumps({"n_slots": n_slots, "in_flight": keep, "slot_log2": slot_log2, "ms_per_step": round(dt / steps * 1e3, 3),
                      "GSps": round(L * steps / dt / 1e9, 2), "host_blocked_share": round(state["blocked"] / dt, 3)}), flush=True)
    rg.close(); pl.close()
