#!/bin/bash
mkdir -p gpurun_out
timeout 400 ncu --set full --import-source on --clock-control none -k regex:cholDiag -s 20 -c 1 -f -o gpurun_out/r2_chol_diag_v3 python tools/cholesky_bench.py 32 > gpurun_out/r2_chol_diag_ncu.log 2>&1
python tools/ncu_summary.py gpurun_out/r2_chol_diag_v3.ncu-rep 0 > gpurun_out/r2_chol_diag_v3_metrics.txt 2>&1; cat gpurun_out/r2_chol_diag_v3_metrics.txt
timeout 400 ncu --set full --import-source on --clock-control none -k regex:cholPanel -s 8 -c 1 -f -o gpurun_out/r2_chol_panel_v3 python tools/cholesky_bench.py 32 > gpurun_out/r2_chol_panel_ncu.log 2>&1
python tools/ncu_summary.py gpurun_out/r2_chol_panel_v3.ncu-rep 0 > gpurun_out/r2_chol_panel_v3_metrics.txt 2>&1; cat gpurun_out/r2_chol_panel_v3_metrics.txt
