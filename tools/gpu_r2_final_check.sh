#!/bin/bash
# Round 2: last check of the final tree on one GPU -- whole GPU suite, smoke, a short default bench line
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest_gpu.log 2>&1; tail -4 gpurun_out/r2_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; tail -2 gpurun_out/r2_smoke.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r2_bench_1gpu_check.log 2>&1; tail -1 gpurun_out/r2_bench_1gpu_check.log | cut -c1-260
