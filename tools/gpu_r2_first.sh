#!/bin/bash
# Round 2, first GPU call (1 GPU): host facts, what round 1 left unrun (tools/next_round.sh), ncu --set full of the shipped orbit kernels.
mkdir -p gpurun_out
{ nproc; lscpu | grep -E "Model name|Socket|NUMA|Thread|Core|^CPU\(s\)"; free -g | head -2; nvidia-smi topo -m; cat /sys/fs/cgroup/cpu.max 2>/dev/null; numactl -H 2>/dev/null; } > gpurun_out/r2_host.log 2>&1
bash tools/next_round.sh 2>&1 | tee gpurun_out/r2_next_round.log
timeout 200 python tools/d2h_probe.py 8 > gpurun_out/r2_d2h_probe.log 2>&1; cat gpurun_out/r2_d2h_probe.log
timeout 400 ncu --replay-mode application --set full --import-source on --clock-control none -k regex:tquOrbit -f -o gpurun_out/r2_orbit_full tools/bin/orbit_check prof > gpurun_out/r2_ncu_orbit.log 2>&1
tail -3 gpurun_out/r2_ncu_orbit.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_tqu_nside64.csv python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r2_bench_under_ncu.log 2>&1
tail -5 gpurun_out/r2_launches_tqu_nside64.csv | cut -c1-300
