#!/bin/bash
# Rebuild profiles/<TAG>_* from the captures scripts/gpu_prof_r02.sh left in gpurun_out/ and from the
# built library (SASS histograms).  usage: bash scripts/make_r02_summary.sh TAG
set -e
TAG=${1:-r02}
ncu -i gpurun_out/${TAG}_headline.ncu-rep --page raw --csv > /tmp/${TAG}_raw.csv 2>/dev/null
{
echo "# ${TAG} — headline step (config 4: 1,000 arms x 65,536 contexts, fp64 decision + 2^16-sample PCS), ncu --set full on one B200"
echo
echo "Commands: scripts/gpu_prof_r02.sh ${TAG} (each ncu pass runs only after the same command exited 0 without ncu; clocks"
echo "uncontrolled; per-launch times are cold-cache and serialised).  Summary by scripts/make_r02_summary.sh."
echo
echo '```'
cat gpurun_out/${TAG}_headline_run.log
echo '```'
echo
python scripts/ncu_raw_summary.py /tmp/${TAG}_raw.csv
echo
for k in pcs_cdf_sample pcs_cdf_build posterior_grid ocba_scan gp_prepare; do
  ncu -i gpurun_out/${TAG}_headline.ncu-rep --page source --csv --kernel-name regex:$k > /tmp/${TAG}_src_$k.csv 2>/dev/null || continue
  echo "## $k — instruction mix and stall reasons (source page)"
  echo '```'
  python scripts/ncu_source_summary.py /tmp/${TAG}_src_$k.csv 14
  echo '```'
  echo
done
echo "## shared-memory wavefronts (bank conflicts) per kernel"
python - <<PY
import csv
rows = list(csv.reader(open("/tmp/${TAG}_raw.csv")))
hdr = rows[0]; ix = {h: i for i, h in enumerate(hdr)}
print("| kernel | shared wavefronts | of which bank conflicts | LSU writeback active % | IPC |")
print("|---|---|---|---|---|")
for r in rows[2:]:
    name = r[ix["Kernel Name"]].replace("void ", "").replace("crs::<unnamed>::", "").replace("unnamed>::", "")[:40]
    g = lambda k: r[ix[k]] if k in ix else "-"
    print(f"| {name} | {g('l1tex__data_pipe_lsu_wavefronts_mem_shared.sum')} | {g('l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum')} | {g('l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed')} | {g('sm__inst_executed.avg.per_cycle_elapsed')} |")
PY
} > profiles/${TAG}_ncu_summary.md
cp gpurun_out/${TAG}_launches.csv profiles/${TAG}_launches_headline.csv
# launch-list shares of the step
python - <<PY > profiles/${TAG}_launch_shares.md
import csv, collections
rows = list(csv.reader(l for l in open("gpurun_out/${TAG}_launches.csv") if l.startswith('"')))
hdr = rows[0]; ix = {h: i for i, h in enumerate(hdr)}
tot = collections.Counter(); cnt = collections.Counter()
for r in rows[1:]:
    if len(r) != len(hdr) or r[ix["Metric Name"]] != "gpu__time_duration.sum": continue
    name = r[ix["Kernel Name"]].split("(")[0].replace("void ", "").split("::")[-1].split("<")[0]
    v = float(r[ix["Metric Value"]].replace(",", "")); u = r[ix["Metric Unit"]]
    v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(u, 1e-6)
    tot[name] += v; cnt[name] += 1
allv = sum(tot.values())
print("# ${TAG}: ncu launch list of \`python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-sub\` (gpu__time_duration.sum, first 400 launches)\n")
print("| kernel | launches | total ms | avg ms | share |\n|---|---|---|---|---|")
for k, v in tot.most_common():
    print(f"| {k} | {cnt[k]} | {v:.3f} | {v / cnt[k]:.4f} | {v / allv * 100:.1f} % |")
PY
# dram traffic per launch
python - <<PY
import csv, json, os
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
rows = list(csv.reader(open("/tmp/${TAG}_raw.csv")))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}
d = {}
for r in rows[2:]:
    name = r[ix["Kernel Name"]].split("<")[0].split("(")[0].split("::")[-1].replace("void ", "").strip()
    name = {"ocba_scan_row_kernel": "ocba_scan_kernel", "gp_prepare_small_kernel": "gp_prepare_kernel"}.get(name, name)
    d[name + ":f64"] = sum(float(r[ix[m]]) * scale[units[ix[m]]] for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
path = "profiles/r02_traffic.json"
out = json.load(open(path)) if os.path.exists(path) else {}
out["_source"] = "dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full launch per kernel (profiles/*_ncu_summary.md; no L2 flush)"
out["headline"] = d
json.dump(out, open(path, "w"), indent=1)
print(json.dumps(out, indent=1))
PY
# SASS evidence: opcode histogram per kernel of the shipped library
python - <<PY > profiles/${TAG}_sass_histogram.md
import collections, re, subprocess
txt = subprocess.run(["cuobjdump", "-sass", "contextual_rs_b200/libcrs_b200.so"], capture_output=True, text=True).stdout
fn = None; hist = {}
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        fn = m.group(1); hist[fn] = collections.Counter(); continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", line)
    if m and fn:
        hist[fn][m.group(1).split(".")[0] + ("." + m.group(1).split(".")[1] if m.group(1).split(".")[0] in ("UBLKCP", "SYNCS", "LDS", "LDG", "STG", "MUFU", "REDUX", "CREDUX") and "." in m.group(1) else "")] += 1
print("# ${TAG}: SASS opcode histogram of contextual_rs_b200/libcrs_b200.so (cuobjdump -sass, sm_100a)\n")
print("Hot-path evidence: DMMA = mma.sync.m8n8k4.f64 (FP64 tensor path of the posterior grid), UBLKCP = cp.async.bulk (1-D TMA),")
print("SYNCS = mbarrier, REDUX/CREDUX = warp reductions of the scan; no UTC*MMA / LDTM / UTMALDG: fp64 has no tcgen05 kind and the")
print("path moves no 2-D tiles.\n")
tot = collections.Counter()
for h in hist.values(): tot.update(h)
print("whole library:", ", ".join(f"{k} {v}" for k, v in tot.most_common(24)), "\n")
print("| kernel (demangled prefix) | instructions | top opcodes |\n|---|---|---|")
want = ("pcs_cdf_sample", "pcs_cdf_build", "posterior_grid_kernelIdLi8ELi24ELb1", "posterior_grid_kernelIdLi4ELi24ELb0", "ocba_scan_row_kernelIdLi32", "ocba_scan_kernelId", "gp_prepare_small_kernelId", "peer_sum", "arm_fold", "pcs_tree_kernelIf", "levi_scan", "gp_mll", "min_zeta_grad")
for w in want:
    for fn, h in hist.items():
        if w in fn:
            print(f"| {fn[:70]} | {sum(h.values())} | " + ", ".join(f"{k} {v}" for k, v in h.most_common(12)) + " |")
            break
PY
echo done
