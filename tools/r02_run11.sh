#!/bin/bash
# round-2 GPU call 11: how much does the traversal kernel lose when shared memory takes L1 away (staging-scheme feasibility)
mkdir -p gpurun_out
for pad in 0 6144 12288 20480; do
  echo "== NRT_SMEM_PAD=$pad bytes per 128-thread CTA (10 CTAs per SM)"
  NRT_SMEM_PAD=$pad timeout 300 python tools/trav_sweep.py 0,20 sphere_grid,terrain
done > gpurun_out/r02_smem_pad11.log 2>&1
cat gpurun_out/r02_smem_pad11.log
