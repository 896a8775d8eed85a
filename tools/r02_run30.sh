#!/bin/bash
# round-2 GPU call 30: new refill defaults A/B (8 = refill 16) and a sweep of the other knobs around them
mkdir -p gpurun_out
timeout 900 python tools/ao_exp_sweep.py sphere_grid,terrain 00,88,10,20,90,30,40,50,60,70,01,02,03,04,05,06 > gpurun_out/r02_aoexp30.log 2>&1; echo "aoexp rc=$?" >> gpurun_out/r02_aoexp30.log
timeout 900 python tools/path_exp_sweep.py 0000,0088,0010,0020,0040,0050,0060 > gpurun_out/r02_pathexp30.log 2>&1; echo "pathexp rc=$?" >> gpurun_out/r02_pathexp30.log
cat gpurun_out/r02_aoexp30.log gpurun_out/r02_pathexp30.log
