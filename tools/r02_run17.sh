#!/bin/bash
# round-2 GPU call 17: builder stress (random soups, class-boundary sizes, random options) + new middle-phase cases
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_build.py -q -x > gpurun_out/r02_t17.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_t17.log
tail -30 gpurun_out/r02_t17.log
