#!/bin/bash
# round-2 GPU call 36: the reference's own nanosg.h on top of the facade; drop-in programs
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_dropin.py tests/test_gpu_scene.py -q -x > gpurun_out/r02_t36.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_t36.log
tail -12 gpurun_out/r02_t36.log
