#!/bin/bash
# Round 2, 2-GPU call: orbit shards with the compact outbox, both exchange modes, gather, one shared host matrix.
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2_topo_2gpu.log 2>&1; nproc >> gpurun_out/r2_topo_2gpu.log
run() { # name, extra args
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 $2 > gpurun_out/$1.log 2>&1
  tail -1 gpurun_out/$1.log | cut -c1-200
}
run r2_bench_2gpu_nccl ""
run r2_bench_2gpu_pull "--exchange pull"
python - <<'PY'
import json
for name in ("r2_bench_2gpu_nccl", "r2_bench_2gpu_pull"):
    try:
        d = json.loads(open("gpurun_out/%s.log" % name).read().strip().splitlines()[-1])
        print(name, "ms/step %.2f" % d["ms_per_step"], "frac %.3f" % d["roofline"]["frac"], "exchange", d["exchange"]["ms"], "%.0f GB/s in" % d["exchange"]["gbs_in_per_gpu"],
              "gather", d["gather"].get("ms"), d["gather"].get("gbs_in_per_gpu"), d["gather"].get("parity_max_err"),
              "e2e ms %.1f" % d["e2e"]["ms_per_step"], "d2h", d["e2e"]["d2h_bytes_per_step"], d["e2e"].get("host_matrix_max_abs_diff_vs_device"), "parity", d.get("parity_max_err"))
    except Exception as e:
        print(name, "failed:", e)
PY
