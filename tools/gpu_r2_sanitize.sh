#!/bin/bash
# Round 2: compute-sanitizer over every kernel incl. this round's (orbit mode 3, Cholesky groups / look-ahead / step functions)
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck python tools/sanitize_run.py > gpurun_out/r2_sanitize_memcheck.log 2>&1; tail -4 gpurun_out/r2_sanitize_memcheck.log
timeout 1800 compute-sanitizer --tool racecheck python tools/sanitize_run.py > gpurun_out/r2_sanitize_racecheck.log 2>&1; tail -4 gpurun_out/r2_sanitize_racecheck.log
