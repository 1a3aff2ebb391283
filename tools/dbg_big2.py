import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sdr_iq_visualizer_b200 import spectral as sp
from oracle import spectral_ref as sref
n, hop, frames = 65536, 16384, 90
L = n + hop * (frames - 1) + 77
x = sref.synth_iq(L, seed=frames + hop, tone_cycles_per_sample=20000.37 / 65536).astype(np.complex64)
old = sp.SpectralPlan(n, hop, "hann", variant=1)
r0 = old.stft(x, welch=True, maxhold=True)
for rep in range(3):
    pl = sp.SpectralPlan(n, hop, "hann")
    r = pl.stft(x, welch=True, maxhold=True)
    rel = np.abs(r.welch_acc[0] - r0.welch_acc[0]) / np.maximum(r0.welch_acc[0], 1e-30)
    bad = np.flatnonzero(rel > 1e-3)
    print("dbg", os.environ.get("SPX_BIG2_DBG"), "rep", rep, "bad bins", bad.size, "max rel", rel.max(), "first bad", bad[:8], "k1 of bad (mod 256) hist", np.bincount((bad % 256) // 16, minlength=16) if bad.size else None)
    r2 = pl.stft(x, wf_rows=True, welch=True, maxhold=True, vmin=-20, vmax=110)
    rel = np.abs(r2.welch_acc[0] - r0.welch_acc[0]) / np.maximum(r0.welch_acc[0], 1e-30)
    print("   with rows: bad bins", int((rel > 1e-3).sum()))
    pl.close()
