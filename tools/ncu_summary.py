#!/usr/bin/env python
"""Condense an .ncu-rep (ncu --set full) into the few numbers DESIGN.md / bench.py cite.
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/<name>.txt"""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum.per_second", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "launch__waves_per_multiprocessor", "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.avg.per_second",
    "smsp__warps_eligible.avg.per_cycle_active",
]
STALL = "smsp__average_warps_issue_stalled_"

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
name_col = hdr.index("Kernel Name")
print(f"# {rep}: {len(data)} profiled launch(es), ncu --set full --clock-control none")
for r in data:
    print(f"\n## {r[name_col]}")
    for i, h in enumerate(hdr):
        if h in KEYS:
            print(f"{h:75s} {r[i]:>18s} {units[i]}")
    stalls = [(float(r[i]), h[len(STALL):].replace("_per_issue_active.ratio", "")) for i, h in enumerate(hdr)
              if h.startswith(STALL) and h.endswith("_per_issue_active.ratio") and r[i] not in ("", "n/a")]
    print("warp-stall cycles per issued instruction: " + ", ".join(f"{n} {v:.2f}" for v, n in sorted(stalls, reverse=True)[:8]))
    rd = float(r[hdr.index("dram__bytes_read.sum")]) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[units[hdr.index("dram__bytes_read.sum")]]
    wr = float(r[hdr.index("dram__bytes_write.sum")]) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[units[hdr.index("dram__bytes_write.sum")]]
    print(f"traffic = dram read + write = {rd + wr:.0f} bytes per launch")
