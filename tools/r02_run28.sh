#!/bin/bash
# round-2 GPU call 28: refill threshold of the launches with deferred retire
mkdir -p gpurun_out
timeout 900 python tools/ao_exp_sweep.py sphere_grid,terrain 00,10,30,40,50,12,32,52 > gpurun_out/r02_aoexp28.log 2>&1; echo "aoexp rc=$?" >> gpurun_out/r02_aoexp28.log
timeout 900 python tools/path_exp_sweep.py 0000,0010,0030,0040,0050,0012,0032 > gpurun_out/r02_pathexp28.log 2>&1; echo "pathexp rc=$?" >> gpurun_out/r02_pathexp28.log
cat gpurun_out/r02_aoexp28.log gpurun_out/r02_pathexp28.log
