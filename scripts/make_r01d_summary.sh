#!/bin/bash
# Rebuild profiles/r01d_* from the captures scripts/gpu_prof_r01d.sh left in gpurun_out/.
set -e
for f in c2 c4 tree_c3 tree_c4; do ncu -i gpurun_out/prof_$f.ncu-rep --page raw --csv > /tmp/raw_$f.csv 2>/dev/null; done
for f in tree_c3 tree_c4; do ncu -i gpurun_out/prof_$f.ncu-rep --page source --csv > /tmp/src_$f.csv 2>/dev/null; done
{
echo "# r01d — max-tree PCS kernel, in-place row scan, ncu --set full on B200 (clocks uncontrolled; per-launch times are cold-cache/serialised)"
echo
echo "Commands: scripts/gpu_prof_r01d.sh (each ncu pass runs only after the same command exited 0 without ncu); summary by scripts/make_r01d_summary.sh."
echo
echo "## bench.py (c2: 100 arms x 1,024 contexts, fp64) — prepare + posterior grid + fused scan/select"
python scripts/ncu_raw_summary.py /tmp/raw_c2.csv
echo
echo "## bench.py --workload c4 (1,000 arms x 65,536 contexts, fp64) — posterior grid + scan"
python scripts/ncu_raw_summary.py /tmp/raw_c4.csv
echo
echo "## scripts/pcs_tree_run.py 65536 c3 f32 — pcs_tree_kernel<float> (config 3: 256 arms x 8,192 contexts, 2^16 samples)"
python scripts/ncu_raw_summary.py /tmp/raw_tree_c3.csv
echo
tail -1 gpurun_out/plain_tree_c3.log
echo
echo '```'
python scripts/ncu_source_summary.py /tmp/src_tree_c3.csv 16
echo '```'
echo
echo "## scripts/pcs_tree_run.py 65536 c4 f32 16384 — pcs_tree_kernel<float> (config-4 shape: 1,000 arms x 16,384 of the 65,536 contexts)"
python scripts/ncu_raw_summary.py /tmp/raw_tree_c4.csv
echo
tail -1 gpurun_out/plain_tree_c4.log
echo
echo '```'
python scripts/ncu_source_summary.py /tmp/src_tree_c4.csv 16
echo '```'
} > profiles/r01d_ncu_summary.md
for w in c2 c3 c4 c5 loop ref_c2; do cp gpurun_out/bench_$w.json profiles/r01d_bench_$w.json; done
cp gpurun_out/launches_c2.csv profiles/r01d_launches_c2.csv
python - <<'PY'
import csv, json
out = {"_source": "profiles/r01d_ncu_summary.md: dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full launch (fp64 decision kernels, fp32 PCS tables; no L2 flush)"}
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
for tag, key in (("c2", "c2"), ("c4", "c4"), ("tree_c3", "c3_pcs"), ("tree_c4", "c4_pcs_quarter")):
    rows = list(csv.reader(open(f"/tmp/raw_{tag}.csv")))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    d = {}
    for r in rows[2:]:
        name = r[ix["Kernel Name"]].split("<")[0].split("(")[0].split("::")[-1].replace("void ", "").strip()
        name = {"ocba_scan_row_kernel": "ocba_scan_kernel"}.get(name, name)
        b = sum(float(r[ix[m]]) * scale[units[ix[m]]] for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
        d[name] = b
    out[key] = d
json.dump(out, open("profiles/r01d_traffic.json", "w"), indent=1)
print(json.dumps(out, indent=1)[:800])
PY
