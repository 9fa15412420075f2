"""Summarise `ncu --page raw --csv` (full set) per kernel launch into a compact markdown table.
usage: ncu -i X.ncu-rep --page raw --csv > raw.csv; python scripts/ncu_raw_summary.py raw.csv"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}
cols = [
    ("gpu__time_duration.sum", "time"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__registers_per_thread", "regs"),
    ("launch__waves_per_multiprocessor", "waves/SM"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64 pipe %"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "fma pipe %"),
    ("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "alu pipe %"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "xu pipe %"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
    ("dram__bytes_read.sum", "dram rd"),
    ("dram__bytes_write.sum", "dram wr"),
    ("smsp__inst_executed.sum", "warp inst"),
]
cols = [(k, n) for k, n in cols if k in ix]
print("| kernel | " + " | ".join(n for _, n in cols) + " |")
print("|---|" + "---|" * len(cols))
for r in rows[2:]:
    name = r[ix["Kernel Name"]].replace("void ", "").replace("crs::<unnamed>::", "").replace("unnamed>::", "")[:44]
    cells = []
    for k, _ in cols:
        v, u = r[ix[k]], units[ix[k]]
        try:
            fv = float(v)
            v = f"{fv:.4g}"
        except ValueError:
            pass
        cells.append(f"{v} {u}".strip())
    print(f"| {name} | " + " | ".join(cells) + " |")
