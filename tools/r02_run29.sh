#!/bin/bash
# round-2 GPU call 29: whole-warp refill (32) on the camera launch x AO launch refill; caller-ray launches
mkdir -p gpurun_out
timeout 900 python tools/ao_exp_sweep.py sphere_grid,terrain 00,40,41,42,43,44,45 > gpurun_out/r02_aoexp29.log 2>&1; echo "aoexp rc=$?" >> gpurun_out/r02_aoexp29.log
NRT_SWEEP_SPP=4 timeout 600 python tools/trav_sweep.py 0,83,84,85,90,93,94 sphere_grid,terrain > gpurun_out/r02_sweep29.log 2>&1; echo "sweep rc=$?" >> gpurun_out/r02_sweep29.log
cat gpurun_out/r02_aoexp29.log gpurun_out/r02_sweep29.log
