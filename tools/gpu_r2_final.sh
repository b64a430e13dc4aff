#!/bin/bash
# Round 2, final 1-GPU pass of the build: whole GPU suite, smoke, default bench line, reference arm, the Cholesky leg, launch list
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest_gpu.log 2>&1; tail -6 gpurun_out/r2_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; tail -2 gpurun_out/r2_smoke.log
timeout 600 python bench.py > gpurun_out/r2_bench_1gpu.log 2>&1
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_reference.log 2>&1; tail -1 gpurun_out/r2_bench_reference.log | cut -c1-200
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2_bench_1gpu.log").read().strip().splitlines()[-1])
    print("ms/step %.2f" % d["ms_per_step"], "frac %.3f" % d["roofline"]["frac"], "traffic", d["roofline"]["traffic"], "e2e ms %.1f" % d["e2e"]["ms_per_step"], "parity", d.get("parity_max_err"), "launches", d["gpu_launches"], "clocks", d["clocks"])
except Exception as e:
    print("bench failed:", e)
PY
timeout 400 python tools/cholesky_bench.py 16 32 > gpurun_out/r2_cholesky_bench_final.log 2>&1; cut -c1-260 gpurun_out/r2_cholesky_bench_final.log
timeout 600 python bench.py --cholesky --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --spot-check 0 > gpurun_out/r2_bench_1gpu_cholesky.log 2>&1
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2_bench_1gpu_cholesky.log").read().strip().splitlines()[-1])
    print("consumer_cholesky", d["consumer_cholesky"])
except Exception as e:
    print("bench --cholesky failed:", e); print(open("gpurun_out/r2_bench_1gpu_cholesky.log").read()[-1500:])
PY
timeout 300 python bench.py --orbit-mode 3 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2_bench_1gpu_mode3.log 2>&1; tail -1 gpurun_out/r2_bench_1gpu_mode3.log | cut -c1-330
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_tqu_nside64.csv python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --spot-check 0 > /dev/null 2>&1
grep -c tquOrbit gpurun_out/r2_launches_tqu_nside64.csv
