#!/bin/bash
# First GPU call of the next round: everything that was written after round 1's GPU time was spent, cheapest first.
#   /usr/local/graft/bin/gpurun --timeout 600 -- 'bash tools/next_round.sh'        (1 GPU, ~5 min)
# Results land in gpurun_out/next_*.log; see DESIGN.md "Open" for what each one decides.
mkdir -p gpurun_out tools/bin
[ -x tools/bin/orbit_check ] || nvcc -O2 -std=c++17 -I include -o tools/bin/orbit_check tools/orbit_check.cu -L cosmopp_b200/lib -lcosmopp_b200 \
    -Xlinker -rpath="$PWD/cosmopp_b200/lib"
# 1. mode 2 of cmg_tqu_orbit (store destinations precomputed per tile): parity + timing next to modes 0 / 1  (~25 s)
timeout 120 tools/bin/orbit_check full > gpurun_out/next_orbit_check.log 2>&1; tail -4 gpurun_out/next_orbit_check.log
# 2. gated GPU tests + the host-expansion whole call  (~2 min with the torch import)
CMG_TEST_UNVERIFIED=1 timeout 400 python -m pytest tests/test_gpu_orbit.py tests/test_zz_gpu_host_expand.py -x -q > gpurun_out/next_pytest.log 2>&1; tail -3 gpurun_out/next_pytest.log
# 3. e2e with and without the host expansion (27 % of the matrix over PCIe + block copies on the host)
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/next_bench_plain.log 2>&1; tail -1 gpurun_out/next_bench_plain.log | cut -c1-400
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --host-expand "$(nproc)" > gpurun_out/next_bench_expand.log 2>&1
python - <<'PY'
import json
for name in ("plain", "expand"):
    try:
        line = json.loads(open("gpurun_out/next_bench_%s.log" % name).read().strip().splitlines()[-1])
        print(name, "ms/matrix %.2f" % line["ms_per_matrix"], "e2e ms/step %.1f" % line["e2e"]["ms_per_step"], "frac %.3f" % line["roofline"]["frac"])
    except Exception as e:
        print(name, "failed:", e)
PY
# then, on 2 GPUs (charged twice):  gpurun --gpus 2 -- 'python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1
#   --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --gather'   (first multi-GPU run of the orbit shards and of their NCCL gather)
