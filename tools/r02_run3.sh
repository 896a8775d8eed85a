#!/bin/bash
# round-2 GPU call 3: camera payload + fast slot math, LDG.256 node fetch variants
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r02_t3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_t3.log
timeout 600 python tools/trav_sweep.py 100,0,20,30,31,32,33,34 sphere_grid,terrain,instanced > gpurun_out/r02_sweep3.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench3.json 2> gpurun_out/r02_bench3.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:traverse_fast3 -s 2 -c 2 -f -o gpurun_out/r02_trav3 python tools/profile_target.py > gpurun_out/r02_ncu3.log 2>&1
tail -5 gpurun_out/r02_t3.log; cat gpurun_out/r02_sweep3.log; cat gpurun_out/r02_bench3.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['avg_launch_ms'])"
