#!/bin/bash
mkdir -p gpurun_out
for s in 0 4000 16000 32000 64000 128000; do CMG_ORBIT_STAGGER=$s timeout 120 tools/bin/orbit_check time; done > gpurun_out/r2_orbit_stagger.log 2>&1
cat gpurun_out/r2_orbit_stagger.log
