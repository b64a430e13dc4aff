#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:chol --csv --log-file gpurun_out/r2_chol_launches.csv python tools/cholesky_bench.py 16 32 > gpurun_out/r2_chol_under_ncu.log 2>&1
python - <<'PY'
import csv, collections
rows = list(csv.reader(open("gpurun_out/r2_chol_launches.csv")))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
h = rows[hdr]
ki, vi, gi = h.index("Kernel Name"), h.index("Metric Value"), h.index("Grid Size")
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[hdr + 1:]:
    if len(r) <= vi: continue
    name = r[ki].split("(")[0]
    agg[name][0] += 1; agg[name][1] += float(r[vi].replace(",", "")) / 1e6
for k, (n, ms) in sorted(agg.items(), key=lambda t: -t[1][1]):
    print("%-40s launches %6d total %.2f ms mean %.4f ms" % (k, n, ms, ms / n))
PY
