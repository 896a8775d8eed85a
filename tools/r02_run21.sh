#!/bin/bash
# round-2 GPU call 21: node steps per exit check (unroll 2/3/4)
mkdir -p gpurun_out
NRT_SWEEP_SPP=4 timeout 600 python tools/trav_sweep.py 0,60,80,81,82,71,90,91,92 sphere_grid,terrain > gpurun_out/r02_sweep21.log 2>&1; echo "sweep rc=$?" >> gpurun_out/r02_sweep21.log
timeout 600 python tools/ao_exp_sweep.py sphere_grid,terrain 00,30,01,31 > gpurun_out/r02_aoexp21.log 2>&1; echo "aoexp rc=$?" >> gpurun_out/r02_aoexp21.log
cat gpurun_out/r02_sweep21.log gpurun_out/r02_aoexp21.log
