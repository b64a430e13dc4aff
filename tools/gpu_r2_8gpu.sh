#!/bin/bash
# Round 2, 8 GPUs: orbit shards with exchange / gather / one shared host matrix, TT orbit shards (configs[2]), batched (configs[3]).
mkdir -p gpurun_out
{ nproc; free -g | head -2; nvidia-smi topo -m | head -12; } > gpurun_out/r2_topo_8gpu.log 2>&1
run() { # name, ngpu, extra args
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $2 --warmup 3 $3 > gpurun_out/$1.log 2>&1
  tail -1 gpurun_out/$1.log | cut -c1-120
}
run r2_bench_8gpu 8 "--steps 10"
run r2_bench_8gpu_tt32 8 "--steps 50 --workload tt_nside32_lmax96"
run r2_bench_8gpu_batched 8 "--steps 5 --workload batched_x1024_tqu_nside16_lmax47"
run r2_bench_8gpu_nccl 8 "--steps 5 --exchange nccl --no-e2e"
run r2_bench_4gpu 4 "--steps 10"
python - <<'PY'
import json
for name in ("r2_bench_8gpu", "r2_bench_8gpu_tt32", "r2_bench_8gpu_batched", "r2_bench_8gpu_nccl", "r2_bench_4gpu"):
    try:
        d = json.loads(open("gpurun_out/%s.log" % name).read().strip().splitlines()[-1])
        ex, ga, e2 = d.get("exchange") or {}, d.get("gather") or {}, d.get("e2e") or {}
        print(name, "ms/step %.3f" % d["ms_per_step"], "frac %.3f" % d["roofline"]["frac"], "exchange", ex.get("ms"), ex.get("mode"), ex.get("gbs_in_per_gpu"), "gather", ga.get("ms"), ga.get("gbs_in_per_gpu"), ga.get("parity_max_err"),
              "e2e ms", e2.get("ms_per_step"), e2.get("d2h_bytes_per_step"), e2.get("host_matrix_max_abs_diff_vs_device"), "parity", d.get("parity_max_err"))
    except Exception as e:
        print(name, "failed:", e)
PY
