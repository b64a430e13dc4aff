#!/bin/bash
# Round 2: syrk with swizzled 16-byte fragment loads -- correctness, timing, one ncu capture
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_cholesky.py -q -x > gpurun_out/r2_pytest_syrk_v3.log 2>&1; tail -3 gpurun_out/r2_pytest_syrk_v3.log
timeout 300 python tools/cholesky_bench.py 16 32 > gpurun_out/r2_cholesky_bench_v5.log 2>&1; cut -c1-330 gpurun_out/r2_cholesky_bench_v5.log
timeout 600 ncu --set full --import-source on --clock-control none -k regex:cholSyrk -s 60 -c 1 -f -o gpurun_out/r2_syrk_v3 python tools/cholesky_bench.py 32 > gpurun_out/r2_syrk_v3_ncu.log 2>&1
python tools/ncu_summary.py gpurun_out/r2_syrk_v3.ncu-rep 0 > gpurun_out/r2_syrk_v3_metrics.txt 2>&1; cat gpurun_out/r2_syrk_v3_metrics.txt
