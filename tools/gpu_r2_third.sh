#!/bin/bash
# Round 2, 1 GPU: packed Cholesky + window tests, the whole GPU suite, Cholesky timing next to cuSOLVER, bench with the direct image.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_cholesky.py tests/test_gpu_window.py -x -q > gpurun_out/r2_pytest_new.log 2>&1; tail -15 gpurun_out/r2_pytest_new.log
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest_gpu.log 2>&1; tail -8 gpurun_out/r2_pytest_gpu.log
timeout 600 python tools/cholesky_bench.py 16 32 > gpurun_out/r2_cholesky_bench.log 2>&1; cat gpurun_out/r2_cholesky_bench.log
timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_1gpu_direct.log 2>&1
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2_bench_1gpu_direct.log").read().strip().splitlines()[-1])
    print("ms/step %.2f" % d["ms_per_step"], "frac %.3f" % d["roofline"]["frac"], "e2e ms %.1f" % d["e2e"]["ms_per_step"], "d2h", d["e2e"]["d2h_bytes_per_step"], "parity", d.get("parity_max_err"))
except Exception as e:
    print("bench failed:", e)
PY
