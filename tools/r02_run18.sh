#!/bin/bash
# round-2 GPU call 18: leaf-batching / deferred-retire knobs of traverse_fast3_kernel -- kernel-only sweep on exported
# rays (StoreHits epilogue) and the whole fused AO pass per experiment setting
mkdir -p gpurun_out
V="0,40,41,42,43,44,45,46,47,48,49,60,61,62,63,64,65,66,67,68,69,20,50,51,52,53,54,55,56,57,58,59,70,71,72,73,74,75,76,77,78,79"
NRT_SWEEP_SPP=4 timeout 600 python tools/trav_sweep.py $V sphere_grid,terrain > gpurun_out/r02_sweep18.log 2>&1; echo "sweep rc=$?" >> gpurun_out/r02_sweep18.log
timeout 600 python tools/ao_exp_sweep.py > gpurun_out/r02_aoexp18.log 2>&1; echo "aoexp rc=$?" >> gpurun_out/r02_aoexp18.log
tail -50 gpurun_out/r02_sweep18.log; cat gpurun_out/r02_aoexp18.log
