#!/bin/bash
# round-2 GPU call 9: builder (segmented small blocks, middle phase, 24-bit curve order, 6 CTAs/SM in phase B): full GPU suite, build times,
# launch list, ncu --set full of subtree_kernel and midtree_kernel
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02_t9.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_t9.log
timeout 600 python tools/build_probe.py instanced 4 > gpurun_out/r02_build9.log 2>&1
timeout 300 python tools/build_probe.py terrain 3 >> gpurun_out/r02_build9.log 2>&1
timeout 300 python tools/build_probe.py sphere_grid 3 >> gpurun_out/r02_build9.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_build_launches9.csv python tools/build_profile_target.py > gpurun_out/r02_build_launches9.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"subtree_kernel|midtree_kernel" -f -o gpurun_out/r02_build9 python tools/build_profile_target.py > gpurun_out/r02_ncu9.log 2>&1
tail -5 gpurun_out/r02_t9.log; cat gpurun_out/r02_build9.log; python tools/launch_list.py gpurun_out/r02_build_launches9.csv
