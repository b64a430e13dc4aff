#!/bin/bash
mkdir -p gpurun_out
timeout 200 python tools/cholesky_bench.py 16 --groups=1,2,3 > gpurun_out/r2_cholesky_bench_small_groups.log 2>&1; cut -c1-200 gpurun_out/r2_cholesky_bench_small_groups.log
