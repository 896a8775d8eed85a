#!/bin/bash
# round-2 GPU call 39: compute-sanitizer (memcheck, initcheck) with the scene AO pass in the target
mkdir -p gpurun_out
for tool in memcheck initcheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 3 python tools/sanitize_target.py > gpurun_out/r02_san39_$tool.log 2>&1; echo "$tool rc=$?" >> gpurun_out/r02_san39_$tool.log
  grep -v "^=========     \|^$" gpurun_out/r02_san39_$tool.log | tail -6
done
