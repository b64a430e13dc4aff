#!/bin/bash
# Round 2: syrk on dense panel planes (16-byte cp.async) + groups of blocks -- correctness, timing per group size, one ncu capture
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_cholesky.py -q -x > gpurun_out/r2_pytest_syrk_v4.log 2>&1; tail -3 gpurun_out/r2_pytest_syrk_v4.log
timeout 400 python tools/cholesky_bench.py 16 32 --groups=1,2,3,4 > gpurun_out/r2_cholesky_bench_v6.log 2>&1; cut -c1-250 gpurun_out/r2_cholesky_bench_v6.log
timeout 300 python tools/sharded_cholesky.py 16 32 --whole > gpurun_out/r2_shchol_1gpu_v2.log 2>&1; tail -2 gpurun_out/r2_shchol_1gpu_v2.log | cut -c1-300
timeout 600 ncu --set full --import-source on --clock-control none -k regex:cholSyrk -s 30 -c 1 -f -o gpurun_out/r2_syrk_v4 python tools/cholesky_bench.py 32 --groups=4 > gpurun_out/r2_syrk_v4_ncu.log 2>&1
python tools/ncu_summary.py gpurun_out/r2_syrk_v4.ncu-rep 0 > gpurun_out/r2_syrk_v4_metrics.txt 2>&1; cat gpurun_out/r2_syrk_v4_metrics.txt
