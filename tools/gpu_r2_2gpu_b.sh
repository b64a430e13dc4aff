#!/bin/bash
# Round 2, 2 GPUs: pull-mode exchange + gather, TT and batched workloads on two ranks.
mkdir -p gpurun_out
run() { # name, extra args
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --warmup 3 $2 > gpurun_out/$1.log 2>&1
  tail -1 gpurun_out/$1.log | cut -c1-160
}
run r2_bench_2gpu "--steps 5"
run r2_bench_2gpu_tt32 "--steps 20 --workload tt_nside32_lmax96"
run r2_bench_2gpu_batched "--steps 3 --workload batched_x1024_tqu_nside16_lmax47"
python - <<'PY'
import json
for name in ("r2_bench_2gpu", "r2_bench_2gpu_tt32", "r2_bench_2gpu_batched"):
    try:
        d = json.loads(open("gpurun_out/%s.log" % name).read().strip().splitlines()[-1])
        ex, ga = d.get("exchange") or {}, d.get("gather") or {}
        print(name, "ms/step %.3f" % d["ms_per_step"], "frac %.3f" % d["roofline"]["frac"], "exchange", ex.get("ms"), ex.get("mode"), "gather", ga.get("ms"), ga.get("gbs_in_per_gpu"), ga.get("parity_max_err"),
              "e2e ms %.1f" % d["e2e"]["ms_per_step"], "parity", d.get("parity_max_err"))
    except Exception as e:
        print(name, "failed:", e)
PY
