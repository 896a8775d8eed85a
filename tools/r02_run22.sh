#!/bin/bash
# round-2 GPU call 22: ncu --set full of the two traversal launches (new policies), launch list of the bench step,
# compute-sanitizer over every kernel family incl. the any-hit / ray32 / deferred-retire paths
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:traverse_fast3 -s 2 -c 2 -f -o gpurun_out/r02_trav22 python tools/profile_target.py > gpurun_out/r02_ncu22.log 2>&1
tail -3 gpurun_out/r02_ncu22.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches22.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-configs > gpurun_out/r02_launches22.log 2>&1
tail -2 gpurun_out/r02_launches22.log | cut -c1-300
for tool in memcheck racecheck initcheck; do
  timeout 1200 compute-sanitizer --tool $tool --error-exitcode 3 python tools/sanitize_target.py > gpurun_out/r02_san22_$tool.log 2>&1; echo "$tool rc=$?" >> gpurun_out/r02_san22_$tool.log
  tail -4 gpurun_out/r02_san22_$tool.log
done
