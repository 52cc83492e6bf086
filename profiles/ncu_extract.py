#!/usr/bin/env python
"""Print the key metrics of every kernel in an .ncu-rep (run here, no GPU):
    python profiles/ncu_extract.py gpurun_out/prof.ncu-rep
"""
import csv
import io
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("sm__inst_executed.sum.per_cycle_elapsed", "warp inst / cycle / SM... (sum)"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("launch__registers_per_thread", "registers"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__occupancy_limit_registers", "occ limit regs (CTAs)"),
    ("launch__occupancy_limit_shared_mem", "occ limit smem (CTAs)"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed", "FMA pipe %"),
    ("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed", "ALU pipe %"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_elapsed", "LSU pipe %"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem wavefronts"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem wavefronts % of peak"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
    ("sm__cycles_elapsed.max", "cycles"),
]
STALLS = "smsp__average_warps_issue_stalled_"

out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    u = dict(zip(hdr, units))
    print(f"### {d['Kernel Name'][:100]}\n")
    print("| metric | value |\n|---|---|")
    for k, label in KEYS:
        if k in d:
            print(f"| {label} (`{k}`) | {d[k]} {u[k]} |")
    st = sorted(((float(v.replace(',', '')), k[len(STALLS):].replace('_per_issue_active.ratio', ''))
                 for k, v in d.items() if k.startswith(STALLS) and k.endswith("_per_issue_active.ratio") and v),
                reverse=True)[:8]
    print("| top stall reasons (warps per issue) | " + ", ".join(f"{n} {v:.2f}" for v, n in st) + " |\n")
