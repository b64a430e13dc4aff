#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29571 tools/sharded_cholesky.py 32 64 --whole > gpurun_out/r2_shchol_la_8gpu.log 2>&1; tail -2 gpurun_out/r2_shchol_la_8gpu.log | cut -c1-900
