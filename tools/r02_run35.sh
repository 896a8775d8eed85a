#!/bin/bash
# round-2 GPU call 35 (2 GPUs): bench line with the concurrent copy probe / collective-safe compact leg, multi-GPU test, C++ fork-per-GPU example
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -q -x > gpurun_out/r02_t35.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_t35.log; tail -2 gpurun_out/r02_t35.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 2 --no-configs > gpurun_out/r02_bench35_n2.json 2> gpurun_out/r02_bench35_n2.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench35_n2.json').read().strip().splitlines()[-1])
print('value',d['value'],'ms',d['ms_per_step'],d['scaling'],'e2e',d['e2e']['value'],'render',d['e2e']['render_api']['value'])
print('concurrent',d['e2e']['concurrent_pinned_copy_gbs'],'single',d['e2e']['measured_pinned_copy_gbs'])
print('compact',d['e2e']['compact_records'])
print('parity ok',d['parity']['ok'])
PY
timeout 300 examples/bin/multi_gpu_ao 2 2>&1 | tail -4
