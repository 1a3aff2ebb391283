import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tools"))
import kernel_sweep as ks
from sdr_iq_visualizer_b200 import spectral as sp
for flush in (False, True):
    ks.run_case(f"headline 2^24 flush={flush}", 4096, 4096, "hann", sp.FMT_CF32, 1 << 24, ["db"], 0, flush=flush)
    ks.run_case(f"C2 shape 2^24 flush={flush}", 4096, 1024, "hann", sp.FMT_CI16, 1 << 24, ["u8", "acc"], 0, flush=flush)
