import sys, os, json
sys.path.insert(0, os.getcwd())
sys.path.insert(0, os.path.join(os.getcwd(), "tools"))
import kernel_sweep as ks
from sdr_iq_visualizer_b200 import spectral as sp
L = 61_440_000
for v in (0, 4):
    for n in (1024, 2048):
        ks.run_case(f"cf32 N={n} 50% u8+acc", n, n // 2, "hann", sp.FMT_CF32, L, ["u8", "acc"], v)
        ks.run_case(f"cf32 N={n} 50% acc only", n, n // 2, "hann", sp.FMT_CF32, L, ["acc"], v)
        ks.run_case(f"cf32 N={n} hop=N f32", n, n, "hann", sp.FMT_CF32, L, ["db"], v)
