#!/usr/bin/env python3
"""End-to-end (host buffers, H2D + D2H inside the call) throughput of the config-2 step vs the H2D piece size of the
three-stream pipeline.  python tools/e2e_sweep.py  -> one JSON line per piece size."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sdr_iq_visualizer_b200 import _native as nat, spectral as sp, synth  # noqa: E402

L, N, HOP = 61_440_000, 4096, 1024
WC = os.environ.get("SPX_E2E_WC", "0") == "1"
host_in = nat.pinned_empty(2 * L, np.int16, write_combined=WC)
host_in[:] = synth.tiled_ci16(L, 2)
F = nat.frame_count(L, N, HOP)
h_wf = nat.pinned_empty((F, N), np.uint8)
h_we = nat.pinned_empty((1, N), np.float64)
h_mh = nat.pinned_empty((1, N), np.float32)
for mib in [float(v) for v in (sys.argv[1:] or ["1", "2", "4", "8", "16", "32"])]:
    os.environ["SPX_H2D_PIECE_BYTES"] = str(int(mib * (1 << 20)))
    pl = sp.SpectralPlan(N, HOP, "hann", sp.FMT_CI16)
    for _ in range(3):
        r = pl.stft(host_in, wf_rows=h_wf, welch=h_we, maxhold=h_mh, vmin=20.0, vmax=130.0)
    nat.device_sync(0)
    ts = []
    for _ in range(10):
        t0 = time.perf_counter()
        r = pl.stft(host_in, wf_rows=h_wf, welch=h_we, maxhold=h_mh, vmin=20.0, vmax=130.0)
        ts.append(time.perf_counter() - t0)
    med = float(np.median(ts))
    print(json.dumps({"wc_input": WC, "piece_mib": mib, "ms_med": round(med * 1e3, 3), "ms_best": round(min(ts) * 1e3, 3),
                      "GSps": round(L / med / 1e9, 2), "h2d_gbs": round(r.h2d_bytes / med / 1e9, 1),
                      "d2h_gbs": round(r.d2h_bytes / med / 1e9, 1)}), flush=True)
    pl.close()
