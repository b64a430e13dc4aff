#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_cholesky.py tests/test_dropin_cpp.py -x -q > gpurun_out/r2_pytest_new.log 2>&1; tail -6 gpurun_out/r2_pytest_new.log
timeout 600 python tools/cholesky_bench.py 16 32 > gpurun_out/r2_cholesky_bench.log 2>&1; cat gpurun_out/r2_cholesky_bench.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:chol --csv --log-file gpurun_out/r2_chol_launches_v3.csv python tools/cholesky_bench.py 16 > /dev/null 2>&1
python - <<'PY'
import csv, collections
rows = list(csv.reader(open("gpurun_out/r2_chol_launches_v3.csv")))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
h = rows[hdr]; ki, vi = h.index("Kernel Name"), h.index("Metric Value")
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[hdr + 1:]:
    if len(r) <= vi: continue
    name = r[ki].split("(")[0]; agg[name][0] += 1; agg[name][1] += float(r[vi].replace(",", "")) / 1e6
for k, (n, ms) in sorted(agg.items(), key=lambda t: -t[1][1]):
    print("%-40s launches %6d total %.2f ms mean %.4f ms" % (k, n, ms, ms / n))
PY
for w in tt_nside32_lmax96 tt_nside16_lmax47 tqu_nside16_lmax47_masked; do
  timeout 300 python bench.py --workload $w --steps 20 --warmup 5 > gpurun_out/r2_bench_$w.log 2>&1
  python - "$w" <<'PY'
import json, sys
try:
    d = json.loads(open("gpurun_out/r2_bench_%s.log" % sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "ms/step %.4f" % d["ms_per_step"], "frac %.3f" % d["roofline"]["frac"], "e2e ms %.3f" % d["e2e"]["ms_per_step"], "cpu", d["cpu_baseline"] and "%.3g" % d["cpu_baseline"]["value"])
except Exception as e:
    print(sys.argv[1], "failed:", e)
PY
done
