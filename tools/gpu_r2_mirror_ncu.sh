#!/bin/bash
mkdir -p gpurun_out
timeout 500 ncu --replay-mode application --set full --import-source on --clock-control none -k regex:tquOrbit -f -o gpurun_out/r2_orbit_mirror tools/bin/orbit_check prof 3 > gpurun_out/r2_ncu_orbit_mirror.log 2>&1
for k in 0 1 2 3; do python tools/ncu_summary.py gpurun_out/r2_orbit_mirror.ncu-rep $k; echo ----; done > gpurun_out/r2_orbit_mirror_metrics.txt 2>&1
grep -E "Kernel Name|gpu__time_duration|fp64_cycles_active|dram__bytes_write|stall|lts__throughput|l1tex__throughput" gpurun_out/r2_orbit_mirror_metrics.txt
