#!/bin/bash
# Round 2: orbit mode 3 (meridian mirror, single owner) -- parity against the oracle and cmg_tqu, timing at Nside 64
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_orbit.py -q -x > gpurun_out/r2_pytest_mirror.log 2>&1; tail -4 gpurun_out/r2_pytest_mirror.log
timeout 400 tools/bin/orbit_check full > gpurun_out/r2_orbit_check_mirror.log 2>&1; grep -E "mode 3|mode 0|CHECK|cmg_tqu [0-9]" gpurun_out/r2_orbit_check_mirror.log | head -30
