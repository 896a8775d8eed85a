#!/bin/bash
# round-2 GPU call 1: parity of the new traversal kernel, policy sweep vs the round-1 kernel, bench, ncu capture
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r02_smi.txt 2>&1
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r02_t1.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_t1.log
timeout 600 python tools/trav_sweep.py 100,0,1,2,3,4,5,6,7,8,9,10,11 sphere_grid,terrain > gpurun_out/r02_sweep1.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r02_bench1.json 2> gpurun_out/r02_bench1.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:traverse_fast3 -s 2 -c 2 -f -o gpurun_out/r02_trav1 python tools/profile_target.py > gpurun_out/r02_ncu1.log 2>&1
tail -3 gpurun_out/r02_t1.log; cat gpurun_out/r02_sweep1.log; cat gpurun_out/r02_bench1.json
