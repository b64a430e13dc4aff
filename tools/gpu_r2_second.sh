#!/bin/bash
# Round 2, second GPU call (1 GPU): the compact-outbox kernels + exchange on one GPU, new host expansion, bench lines.
mkdir -p gpurun_out
df -h /dev/shm /tmp > gpurun_out/r2_shm.log 2>&1; cat gpurun_out/r2_shm.log
timeout 150 tools/bin/orbit_check full > gpurun_out/r2_orbit_check.log 2>&1; tail -12 gpurun_out/r2_orbit_check.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu.log 2>&1; tail -5 gpurun_out/r2_pytest_gpu.log
timeout 400 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench_1gpu.log 2>&1; tail -1 gpurun_out/r2_bench_1gpu.log | cut -c1-300
timeout 300 python bench.py --workload batched_x1024_tqu_nside16_lmax47 --steps 3 --warmup 3 > gpurun_out/r2_bench_batched_1gpu.log 2>&1; tail -1 gpurun_out/r2_bench_batched_1gpu.log | cut -c1-300
python - <<'PY'
import json
for name in ("r2_bench_1gpu", "r2_bench_batched_1gpu"):
    try:
        d = json.loads(open("gpurun_out/%s.log" % name).read().strip().splitlines()[-1])
        print(name, "ms/step %.2f" % d["ms_per_step"], "frac %.3f" % d["roofline"]["frac"], "hbm %.0f GB/s" % d["roofline"]["hbm_write_gbs"],
              "e2e ms %.1f" % d["e2e"]["ms_per_step"], "parity", d.get("parity_max_err"), "cpu", d["cpu_baseline"] and d["cpu_baseline"]["value"])
    except Exception as e:
        print(name, "failed:", e)
PY
