#!/bin/bash
# round-2 GPU call 14: compute-sanitizer over every kernel family incl. the round-2 builder kernels, fp64 conformance build, primitive kinds
mkdir -p gpurun_out
python tools/sanitize_target.py > gpurun_out/r02_san_plain.log 2>&1; echo "plain rc=$?" >> gpurun_out/r02_san_plain.log
for tool in memcheck racecheck initcheck; do
  timeout 1500 compute-sanitizer --tool $tool --error-exitcode 3 python tools/sanitize_target.py > gpurun_out/r02_san_$tool.log 2>&1; echo "$tool rc=$?" >> gpurun_out/r02_san_$tool.log
done
tail -4 gpurun_out/r02_san_plain.log; for tool in memcheck racecheck initcheck; do echo "== $tool"; grep -c "========= " gpurun_out/r02_san_$tool.log; tail -4 gpurun_out/r02_san_$tool.log; done
