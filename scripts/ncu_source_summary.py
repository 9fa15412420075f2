"""Summarise an `ncu --page source --csv` export: instruction mix by opcode and stall reasons.
usage: ncu -i X.ncu-rep --page source --csv --kernel-name regex:NAME > src.csv; python scripts/ncu_source_summary.py src.csv"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
hdr = rows[hi[0]]
ix = {h: i for i, h in enumerate(hdr)}
end = hi[1] - 1 if len(hi) > 1 else len(rows)
first = [r for r in rows[hi[0] + 1:end] if len(r) == len(hdr)]
f = lambda r, k: float(r[ix[k]] or 0)
tot_inst = sum(f(r, "Instructions Executed") for r in first)
tot_samp = sum(f(r, "# Samples") for r in first)
print(rows[0][1][:100])
print(f"SASS rows {len(first)}  warp-instructions {tot_inst:.4g}  samples {tot_samp:.0f}")
cls, samp = collections.Counter(), collections.Counter()
for r in first:
    toks = r[ix["Source"]].split()
    op = toks[1] if toks[0].startswith("@") else toks[0]
    op = op.split(".")[0]
    cls[op] += f(r, "Instructions Executed")
    samp[op] += f(r, "# Samples")
for op, c in cls.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 22):
    print(f"  {op:10s} inst {c / tot_inst * 100:5.1f}%   samples {samp[op] / tot_samp * 100:5.1f}%")
stalls = [h for h in hdr if h.startswith("stall_")]
agg = {s: sum(f(r, s) for r in first) for s in stalls}
tot = sum(agg.values()) or 1.0
print("stall reasons (all samples):", ", ".join(f"{s[6:]} {v / tot * 100:.1f}%" for s, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
