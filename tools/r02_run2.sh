#!/bin/bash
# round-2 GPU call 2: traversal kernel after the pop-loop fix; 64-byte vs 128-byte node A/B; 10 M-triangle scene sweep incl. the TMA / smem-stack variants
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r02_t2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_t2.log
timeout 600 python tools/trav_sweep.py 100,0,1,2,4,5,6,7,9,20,21,22,23 sphere_grid,terrain > gpurun_out/r02_sweep2.log 2>&1
timeout 900 python tools/trav_sweep.py 100,130,140,0,2,20,21 instanced > gpurun_out/r02_sweep2_10m.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:traverse_fast3 -s 2 -c 2 -f -o gpurun_out/r02_trav2 python tools/profile_target.py > gpurun_out/r02_ncu2.log 2>&1
tail -3 gpurun_out/r02_t2.log; cat gpurun_out/r02_sweep2.log gpurun_out/r02_sweep2_10m.log
