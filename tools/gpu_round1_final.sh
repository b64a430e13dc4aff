#!/bin/bash
# One gpurun call: ncu counters of the orbit kernels, per-rank timings, bench line, GPU tests of the orbit path.
mkdir -p gpurun_out tools/bin
# the probe is a plain C-ABI client (no Python): built here when the snapshot did not bring it along
[ -x tools/bin/orbit_check ] || nvcc -O2 -std=c++17 -I include -o tools/bin/orbit_check tools/orbit_check.cu -L cosmopp_b200/lib -lcosmopp_b200 \
    -Xlinker -rpath="$PWD/cosmopp_b200/lib"
M=gpu__time_duration.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__warps_active.avg.per_cycle_active,launch__registers_per_thread,launch__occupancy_limit_shared_mem,launch__occupancy_limit_registers,launch__grid_size,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum,l1tex__t_requests_pipe_lsu_mem_global_op_st.sum,lts__t_sectors_srcunit_tex_op_write.sum,sm__cycles_elapsed.max,smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio
timeout 70 ncu --replay-mode application --clock-control none --metrics $M -k regex:tquOrbit -f -o gpurun_out/r1_orbit tools/bin/orbit_check prof > gpurun_out/ncu_orbit.log 2>&1
tail -3 gpurun_out/ncu_orbit.log
timeout 40 tools/bin/orbit_check ranks > gpurun_out/orbit_ranks.log 2>&1
grep -v "^  world" gpurun_out/orbit_ranks.log
timeout 150 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_orbit_1gpu.log 2>&1
tail -2 gpurun_out/bench_orbit_1gpu.log | cut -c1-600
timeout 200 python -m pytest tests/test_gpu_orbit.py -x -q > gpurun_out/pytest_orbit.log 2>&1
tail -5 gpurun_out/pytest_orbit.log
