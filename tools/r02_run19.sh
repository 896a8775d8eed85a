#!/bin/bash
# round-2 GPU call 19: new default policies (leaf batching, deferred retire on the camera launch): A/B against the old
# policy in the fused AO pass, path tracer launches swept, traversal / render / path parity tests
mkdir -p gpurun_out
timeout 600 python tools/ao_exp_sweep.py sphere_grid,terrain 00,88,80,08 > gpurun_out/r02_aoexp19.log 2>&1; echo "aoexp rc=$?" >> gpurun_out/r02_aoexp19.log
timeout 900 python tools/path_exp_sweep.py 0000,0088,0080,0008,0010,0020,0050,0060,0002,0062,0068 > gpurun_out/r02_pathexp19.log 2>&1; echo "pathexp rc=$?" >> gpurun_out/r02_pathexp19.log
timeout 1500 python -m pytest tests/test_gpu_traverse.py tests/test_gpu_render.py tests/test_gpu_golden.py tests/test_gpu_edge.py tests/test_gpu_path.py -q -x > gpurun_out/r02_t19.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_t19.log
cat gpurun_out/r02_aoexp19.log gpurun_out/r02_pathexp19.log; tail -15 gpurun_out/r02_t19.log
