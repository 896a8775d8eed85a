#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/f64_probe.py > gpurun_out/r02_f64_probe25.log 2>&1; echo "probe rc=$?" >> gpurun_out/r02_f64_probe25.log
cat gpurun_out/r02_f64_probe25.log
