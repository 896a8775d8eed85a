#!/bin/bash
# round-2 GPU call 7: full GPU suite (no -x), lane-state statistics, variant sweeps (incl. the 10 M-triangle scene with the
# smem-stack / TMA-treelet variants of the round-1 kernel), ncu --set full on the 10 M scene
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02_t7.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_t7.log
timeout 600 python tools/lane_stats.py sphere_grid,terrain > gpurun_out/r02_lane_stats7.jsonl 2> gpurun_out/r02_lane_stats7.err
timeout 600 python tools/trav_sweep.py 100,0,1,2,4,5,6,7,8,9,10,20,21,23,30 sphere_grid,terrain > gpurun_out/r02_sweep7.log 2>&1
timeout 900 python tools/trav_sweep.py 100,130,140,0,19,20,21,30 instanced > gpurun_out/r02_sweep7_10m.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:traverse_fast3 -s 2 -c 2 -f -o gpurun_out/r02_trav7_10m python tools/profile_target.py instanced > gpurun_out/r02_ncu7.log 2>&1
tail -5 gpurun_out/r02_t7.log; cat gpurun_out/r02_lane_stats7.jsonl | cut -c1-900; cat gpurun_out/r02_sweep7.log gpurun_out/r02_sweep7_10m.log; tail -3 gpurun_out/r02_ncu7.log
