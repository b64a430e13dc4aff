#!/bin/bash
# Round 2: DMMA panel kernel + rsqrt pivots -- correctness, timing, per-kernel launch list
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_cholesky.py -q -x > gpurun_out/r2_pytest_syrk_v9.log 2>&1; tail -3 gpurun_out/r2_pytest_syrk_v6.log
timeout 400 python tools/cholesky_bench.py 16 32 --groups=2,4 > gpurun_out/r2_cholesky_bench_v11.log 2>&1; cut -c1-250 gpurun_out/r2_cholesky_bench_v11.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:chol --csv --log-file gpurun_out/r2_chol_launches_v7.csv python tools/cholesky_bench.py 16 32 --groups=4 > gpurun_out/r2_chol_under_ncu.log 2>&1
python - <<'PY'
import csv, collections
rows = list(csv.reader(open("gpurun_out/r2_chol_launches_v7.csv")))
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
