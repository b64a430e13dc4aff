#!/bin/bash
# Round 2: sharded Cholesky on 8 GPUs -- Nside 32 against the whole-matrix factorisation, Nside 64 (BASELINE configs[4], 87 GB) timed
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 tools/sharded_cholesky.py 32 --whole --group=4 > gpurun_out/r2_shchol_8gpu_small.log 2>&1; tail -1 gpurun_out/r2_shchol_8gpu_small.log | cut -c1-1300
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29542 tools/sharded_cholesky.py 64 --group=4 > gpurun_out/r2_shchol_8gpu_nside64.log 2>&1; tail -1 gpurun_out/r2_shchol_8gpu_nside64.log | cut -c1-1300
