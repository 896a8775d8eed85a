#!/bin/bash
# round-2 GPU call 5: path tracer per-bounce parity against the reference's functions; full GPU suite
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_path.py -x -q > gpurun_out/r02_t5_path.log 2>&1; echo "rc=$?" >> gpurun_out/r02_t5_path.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_t5.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_t5.log
tail -30 gpurun_out/r02_t5_path.log; tail -5 gpurun_out/r02_t5.log
