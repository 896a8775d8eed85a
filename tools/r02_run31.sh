#!/bin/bash
# round-2 GPU call 31: CTAs per SM x unroll on the camera / AO launches
mkdir -p gpurun_out
timeout 900 python tools/ao_exp_sweep.py sphere_grid,terrain 00,60,70,90,50,65,95,69,99 > gpurun_out/r02_aoexp31.log 2>&1; echo "aoexp rc=$?" >> gpurun_out/r02_aoexp31.log
cat gpurun_out/r02_aoexp31.log
