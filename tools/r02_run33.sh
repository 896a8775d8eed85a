#!/bin/bash
# round-2 GPU call 33: zero-copy small-call path: tests (traverse, edge, drop-in programs incl. per-ray objrender, f64), latency probe
mkdir -p gpurun_out
timeout 1800 python -m pytest tests/test_gpu_traverse.py tests/test_gpu_edge.py tests/test_gpu_dropin.py tests/test_gpu_f64.py tests/test_gpu_prims.py tests/test_gpu_errors.py -q -x -s > gpurun_out/r02_t33.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_t33.log
grep -i "mrays\|passed\|failed\|error" gpurun_out/r02_t33.log | tail -12
timeout 600 python tools/per_ray_probe.py > gpurun_out/r02_per_ray33.log 2>&1; cat gpurun_out/r02_per_ray33.log
