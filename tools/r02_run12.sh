#!/bin/bash
# round-2 GPU call 12 (2 GPUs): rehearsal of configs[4] (4096x4096x256 spp, 64x64 tiles, sharded + gathered) at world = 2;
# the C++ fork-per-GPU example's report; per-ray facade Traverse cost (the reference's objrender program, -s output)
mkdir -p gpurun_out
( cd examples && timeout 300 ./bin/multi_gpu_ao 2 1920 1080 64 5 ) > gpurun_out/r02_cpp_multi12.log 2>&1; echo "rc=$?" >> gpurun_out/r02_cpp_multi12.log
NRT_BENCH_C5_MIN_WORLD=2 timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r02_bench12_c5.json 2> gpurun_out/r02_bench12_c5.err; echo "rc=$?" >> gpurun_out/r02_bench12_c5.err
timeout 600 python -m pytest tests/test_gpu_dropin.py -q -s -k "objrender" > gpurun_out/r02_t12_objrender.log 2>&1
timeout 900 python -m pytest tests/test_gpu_path.py -q > gpurun_out/r02_t12_path.log 2>&1; echo "rc=$?" >> gpurun_out/r02_t12_path.log
cat gpurun_out/r02_cpp_multi12.log; tail -3 gpurun_out/r02_bench12_c5.err; grep objrender gpurun_out/r02_t12_objrender.log; tail -3 gpurun_out/r02_t12_path.log
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench12_c5.json').read().strip().splitlines()[-1])
print('value',d['value'],'ms',d['ms_per_step'])
for c in d.get('configs',[]): print('CONFIG', json.dumps(c)[:1500])
PY
