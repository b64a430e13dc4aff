#!/bin/bash
# Round 2, 8 GPUs, second pass: pull-mode exchange / gather with a rotated peer order next to NCCL.
mkdir -p gpurun_out
run() { # name, ngpu, extra args
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $2 --warmup 3 $3 > gpurun_out/$1.log 2>&1
  tail -1 gpurun_out/$1.log | cut -c1-120
}
run r2_bench_8gpu_pull_rot 8 "--steps 10 --no-e2e"
run r2_bench_8gpu_nccl2 8 "--steps 10 --exchange nccl --no-e2e"
python - <<'PY'
import json
for name in ("r2_bench_8gpu_pull_rot", "r2_bench_8gpu_nccl2"):
    try:
        d = json.loads(open("gpurun_out/%s.log" % name).read().strip().splitlines()[-1])
        ex, ga = d.get("exchange") or {}, d.get("gather") or {}
        print(name, "ms/step %.3f" % d["ms_per_step"], "frac %.3f" % d["roofline"]["frac"], "exchange", ex.get("ms"), ex.get("mode"), ex.get("gbs_in_per_gpu"), "gather", ga.get("ms"), ga.get("gbs_in_per_gpu"), ga.get("parity_max_err"), "parity", d.get("parity_max_err"))
    except Exception as e:
        print(name, "failed:", e)
PY
