#!/bin/bash
# round-2 GPU call 27: second policy sweep around the new defaults (whole fused passes)
mkdir -p gpurun_out
timeout 900 python tools/ao_exp_sweep.py sphere_grid,terrain 00,10,20,30,40,50,60,70,90,01,02,03,04,05,06,07,09 > gpurun_out/r02_aoexp27.log 2>&1; echo "aoexp rc=$?" >> gpurun_out/r02_aoexp27.log
timeout 900 python tools/path_exp_sweep.py 0000,0010,0030,0040,0050,0090,0001,0003,0004,0009 > gpurun_out/r02_pathexp27.log 2>&1; echo "pathexp rc=$?" >> gpurun_out/r02_pathexp27.log
cat gpurun_out/r02_aoexp27.log gpurun_out/r02_pathexp27.log
