#!/bin/bash
# round-2 GPU call 38: nrt_scene_render_ao_device (AO pass over two-level scenes)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_scene.py -q -x > gpurun_out/r02_t38.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_t38.log
tail -30 gpurun_out/r02_t38.log
