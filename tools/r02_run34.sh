#!/bin/bash
# round-2 GPU call 34: compute-sanitizer over the fp64 fast kernel, its layout kernels and the zero-copy small-call paths
mkdir -p gpurun_out
for tool in memcheck racecheck initcheck; do
  timeout 1200 compute-sanitizer --tool $tool --error-exitcode 3 python tools/sanitize_target.py > gpurun_out/r02_san34_$tool.log 2>&1; echo "$tool rc=$?" >> gpurun_out/r02_san34_$tool.log
  grep -v "^=========     \|^$" gpurun_out/r02_san34_$tool.log | tail -8
done
