#!/bin/bash
mkdir -p gpurun_out
{ echo "streaming stores (st.global.cs, shipped):"; timeout 120 tools/bin/orbit_check time; echo "write-back stores (-DCMG_ORBIT_STORE_WB):"; LD_LIBRARY_PATH=$PWD/tools/bin/libexp timeout 120 tools/bin/orbit_check time; } > gpurun_out/r2_orbit_storepolicy.log 2>&1
cat gpurun_out/r2_orbit_storepolicy.log
