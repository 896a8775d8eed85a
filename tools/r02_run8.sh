#!/bin/bash
# round-2 GPU call 8: fixed tests (node list on the device's own tree, bit-equal camera rays), builder with per-subtree id
# reservation + register-resident node records: tests, 10 M build time, launch list
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_prims.py tests/test_gpu_path.py tests/test_gpu_build.py tests/test_gpu_build_ref.py tests/test_gpu_scene.py tests/test_gpu_render.py tests/test_gpu_dropin.py -q -x > gpurun_out/r02_t8.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_t8.log
timeout 600 python tools/build_probe.py instanced 4 > gpurun_out/r02_build8.log 2>&1
timeout 300 python tools/build_probe.py terrain 3 >> gpurun_out/r02_build8.log 2>&1
timeout 300 python tools/build_probe.py sphere_grid 3 >> gpurun_out/r02_build8.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_build_launches8.csv python tools/build_profile_target.py > gpurun_out/r02_build_launches8.log 2>&1
tail -5 gpurun_out/r02_t8.log; cat gpurun_out/r02_build8.log; python tools/launch_list.py gpurun_out/r02_build_launches8.csv
