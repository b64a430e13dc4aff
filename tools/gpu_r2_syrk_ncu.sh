#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none -k regex:cholSyrk -s 60 -c 1 -f -o gpurun_out/r2_syrk python tools/cholesky_bench.py 32 > gpurun_out/r2_syrk_ncu.log 2>&1
tail -3 gpurun_out/r2_syrk_ncu.log
python tools/ncu_summary.py gpurun_out/r2_syrk.ncu-rep 0 > gpurun_out/r2_syrk_metrics.txt 2>&1; cat gpurun_out/r2_syrk_metrics.txt
