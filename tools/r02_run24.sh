#!/bin/bash
# round-2 GPU call 24: BVHAccel<double> fast kernel: tests, drop-in programs, throughput probe
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_f64.py tests/test_gpu_dropin.py -q -x > gpurun_out/r02_t24.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_t24.log
tail -25 gpurun_out/r02_t24.log
timeout 900 python tools/f64_probe.py > gpurun_out/r02_f64_probe24.log 2>&1; echo "probe rc=$?" >> gpurun_out/r02_f64_probe24.log
cat gpurun_out/r02_f64_probe24.log
