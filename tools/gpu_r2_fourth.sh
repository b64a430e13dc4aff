#!/bin/bash
# Round 2, 1 GPU: drop-in rework + Cholesky v2: whole GPU suite, Cholesky timing (16, 32, 64), SYRK counters.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest_gpu.log 2>&1; tail -12 gpurun_out/r2_pytest_gpu.log
timeout 900 python tools/cholesky_bench.py 16 32 64 > gpurun_out/r2_cholesky_bench.log 2>&1; cat gpurun_out/r2_cholesky_bench.log
timeout 600 ncu --set full --import-source on --clock-control none -k regex:cholSyrk -s 60 -c 1 -f -o gpurun_out/r2_syrk_v2 python tools/cholesky_bench.py 32 > gpurun_out/r2_syrk_ncu.log 2>&1
python tools/ncu_summary.py gpurun_out/r2_syrk_v2.ncu-rep 0 > gpurun_out/r2_syrk_v2_metrics.txt 2>&1; head -12 gpurun_out/r2_syrk_v2_metrics.txt; tail -8 gpurun_out/r2_syrk_v2_metrics.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:chol --csv --log-file gpurun_out/r2_chol_launches_v2.csv python tools/cholesky_bench.py 16 > /dev/null 2>&1
python - <<'PY'
import csv, collections
rows = list(csv.reader(open("gpurun_out/r2_chol_launches_v2.csv")))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
h = rows[hdr]; ki, vi = h.index("Kernel Name"), h.index("Metric Value")
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[hdr + 1:]:
    if len(r) <= vi: continue
    name = r[ki].split("(")[0]; agg[name][0] += 1; agg[name][1] += float(r[vi].replace(",", "")) / 1e6
for k, (n, ms) in sorted(agg.items(), key=lambda t: -t[1][1]):
    print("%-40s launches %6d total %.2f ms mean %.4f ms" % (k, n, ms, ms / n))
PY
