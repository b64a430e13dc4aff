#!/bin/bash
# Round 2: look-ahead in the sharded Cholesky driver -- one-GPU tests, two ranks against the one-GPU factor, Nside 64 with and without
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_cholesky.py -q -x > gpurun_out/r2_pytest_shchol_la.log 2>&1; tail -3 gpurun_out/r2_pytest_shchol_la.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 tools/sharded_cholesky.py 32 --whole > gpurun_out/r2_shchol_la_2gpu_small.log 2>&1; tail -1 gpurun_out/r2_shchol_la_2gpu_small.log | cut -c1-1200
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29562 tools/sharded_cholesky.py 64 > gpurun_out/r2_shchol_la_2gpu_nside64.log 2>&1; tail -1 gpurun_out/r2_shchol_la_2gpu_nside64.log | cut -c1-700
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29563 tools/sharded_cholesky.py 64 --no-lookahead > gpurun_out/r2_shchol_serial_2gpu_nside64.log 2>&1; tail -1 gpurun_out/r2_shchol_serial_2gpu_nside64.log | cut -c1-700
