#!/bin/bash
# Round 2: Cholesky with look-ahead -- correctness, timing with and without, ncu of the main trailing update
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_cholesky.py -q -x > gpurun_out/r2_pytest_syrk_v5.log 2>&1; tail -3 gpurun_out/r2_pytest_syrk_v5.log
timeout 400 python tools/cholesky_bench.py 16 32 --groups=2,4 > gpurun_out/r2_cholesky_bench_v7.log 2>&1; cut -c1-250 gpurun_out/r2_cholesky_bench_v7.log
timeout 400 python tools/cholesky_bench.py 16 32 --groups=4 --no-lookahead > gpurun_out/r2_cholesky_bench_v7_serial.log 2>&1; cut -c1-250 gpurun_out/r2_cholesky_bench_v7_serial.log
timeout 600 ncu --set full --import-source on --clock-control none -k regex:cholSyrk -s 10 -c 1 -f -o gpurun_out/r2_syrk_v5 python tools/cholesky_bench.py 32 --groups=4 > gpurun_out/r2_syrk_v5_ncu.log 2>&1
python tools/ncu_summary.py gpurun_out/r2_syrk_v5.ncu-rep 0 > gpurun_out/r2_syrk_v5_metrics.txt 2>&1; cat gpurun_out/r2_syrk_v5_metrics.txt
