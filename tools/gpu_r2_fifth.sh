#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_cholesky.py tests/test_gpu_orbit.py -x -q > gpurun_out/r2_pytest_new.log 2>&1; tail -6 gpurun_out/r2_pytest_new.log
timeout 600 python tools/cholesky_bench.py 16 32 > gpurun_out/r2_cholesky_bench.log 2>&1; cat gpurun_out/r2_cholesky_bench.log
timeout 300 ncu --set full --import-source on --clock-control none -k regex:cholPanel -s 20 -c 1 -f -o gpurun_out/r2_panel python tools/cholesky_bench.py 16 > /dev/null 2>&1
python tools/ncu_summary.py gpurun_out/r2_panel.ncu-rep 0 > gpurun_out/r2_panel_metrics.txt 2>&1; cat gpurun_out/r2_panel_metrics.txt | head -40
timeout 300 ncu --set full --import-source on --clock-control none -k regex:cholDiag -s 20 -c 1 -f -o gpurun_out/r2_diag python tools/cholesky_bench.py 16 > /dev/null 2>&1
python tools/ncu_summary.py gpurun_out/r2_diag.ncu-rep 0 > gpurun_out/r2_diag_metrics.txt 2>&1; head -3 gpurun_out/r2_diag_metrics.txt; tail -8 gpurun_out/r2_diag_metrics.txt
timeout 300 python tools/orbit_check.py tt > /dev/null 2>&1
timeout 120 tools/bin/orbit_check tt > gpurun_out/r2_orbit_tt_check.log 2>&1; tail -6 gpurun_out/r2_orbit_tt_check.log
