#!/bin/bash
# Round 2, final 8-GPU pass: sharded Cholesky against the one-GPU factor at Nside 32, then the default bench line with the Cholesky leg
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 tools/sharded_cholesky.py 32 --whole > gpurun_out/r2_shchol_8gpu_small_final.log 2>&1; tail -1 gpurun_out/r2_shchol_8gpu_small_final.log | cut -c1-1300
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29552 bench.py --gpus 8 --steps 10 --warmup 3 --cholesky > gpurun_out/r2_bench_8gpu_final.log 2>&1
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2_bench_8gpu_final.log").read().strip().splitlines()[-1])
    print("ms/step %.3f" % d["ms_per_step"], "frac %.3f" % d["roofline"]["frac"], "e2e ms %.1f" % d["e2e"]["ms_per_step"], "exchange", d["exchange"]["ms"], "gather", d["gather"]["ms"], "parity", d.get("parity_max_err"))
    print("consumer_cholesky", d["consumer_cholesky"])
except Exception as e:
    print("bench failed:", e); print(open("gpurun_out/r2_bench_8gpu_final.log").read()[-2500:])
PY
