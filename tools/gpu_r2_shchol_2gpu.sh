#!/bin/bash
# Round 2: sharded Cholesky -- emulated ranks on one GPU, then two real ranks (Nside 16 / 32 against the whole-matrix factorisation, Nside 64 timed)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_cholesky.py -q -x > gpurun_out/r2_pytest_shchol.log 2>&1; tail -5 gpurun_out/r2_pytest_shchol.log
timeout 300 python tools/cholesky_bench.py 16 32 > gpurun_out/r2_cholesky_bench_v4.log 2>&1; cat gpurun_out/r2_cholesky_bench_v4.log | cut -c1-400
timeout 300 python tools/sharded_cholesky.py 16 32 --whole > gpurun_out/r2_shchol_1gpu.log 2>&1; tail -3 gpurun_out/r2_shchol_1gpu.log | cut -c1-900
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/sharded_cholesky.py 16 32 --whole > gpurun_out/r2_shchol_2gpu_small.log 2>&1; tail -3 gpurun_out/r2_shchol_2gpu_small.log | cut -c1-1200
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 tools/sharded_cholesky.py 64 > gpurun_out/r2_shchol_2gpu_nside64.log 2>&1; tail -2 gpurun_out/r2_shchol_2gpu_nside64.log | cut -c1-1200
