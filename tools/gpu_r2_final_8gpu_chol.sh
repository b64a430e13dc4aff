#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29581 bench.py --gpus 8 --steps 5 --warmup 3 --cholesky --no-e2e --no-gather --spot-check 2000 > gpurun_out/r2_bench_8gpu_cholesky.log 2>&1
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2_bench_8gpu_cholesky.log").read().strip().splitlines()[-1])
    print("ms/step %.3f" % d["ms_per_step"], "frac %.3f" % d["roofline"]["frac"], "parity", d.get("parity_max_err"))
    print("consumer_cholesky", d["consumer_cholesky"])
except Exception as e:
    print("bench failed:", e); print(open("gpurun_out/r2_bench_8gpu_cholesky.log").read()[-2500:])
PY
