#!/usr/bin/env python3
"""Summarise an .ncu-rep (read on the CPU box): python tools/ncu_summary.py <rep> [name] -> JSON on stdout.
Per profiled launch: duration, DRAM bytes, pipe utilisation, issue rate, stall reasons per issued instruction."""
import csv
import json
import subprocess
import sys

rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
H, units = rows[0], rows[1]


def num(v):
    try:
        return float(v.replace(",", ""))
    except Exception:
        return v


WANT = {
    "duration": "gpu__time_duration.sum", "dram_read": "dram__bytes_read.sum", "dram_write": "dram__bytes_write.sum",
    "dram_pct": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "regs": "launch__registers_per_thread", "warps_active_pct": "sm__warps_active.avg.pct_of_peak_sustained_active",
    "issue_active_pct": "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm_throughput_pct": "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "warp_inst": "smsp__inst_executed.sum", "sm_ghz": "sm__cycles_elapsed.avg.per_second",
    "pipe_fma_pct": "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "pipe_fma_cycles_pct": "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "pipe_fmaheavy_pct": "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active",
    "pipe_fmalite_pct": "sm__inst_executed_pipe_fmalite.avg.pct_of_peak_sustained_active",
    "pipe_alu_pct": "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "pipe_xu_pct": "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "pipe_lsu_pct": "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "smem_wavefronts": "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "smem_bank_conflicts": "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l2_pct": "lts__t_sectors.avg.pct_of_peak_sustained_elapsed",
}
res = []
for r in rows[2:]:
    d = {"kernel": r[H.index("Kernel Name")], "grid": r[H.index("Grid Size")], "block": r[H.index("Block Size")]}
    for k, m in WANT.items():
        if m in H:
            d[k] = num(r[H.index(m)])
            d[k + "_unit"] = units[H.index(m)]
    for i, h in enumerate(H):
        if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
            v = num(r[i])
            if isinstance(v, float) and v >= 0.03:
                d["stall_" + h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]] = round(v, 3)
    d = {k: v for k, v in d.items() if not (k.endswith("_unit") and v in ("%", ""))}
    res.append(d)
print(json.dumps({"report": rep, "launches": res}, indent=1))
