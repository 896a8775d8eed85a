#!/bin/bash
# round-2 GPU call 26 (2 GPUs): multi-GPU test + strong-scaling bench line with the new traversal policies
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -q -x > gpurun_out/r02_t26.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_t26.log
tail -5 gpurun_out/r02_t26.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --no-configs > gpurun_out/r02_bench26_n2.json 2> gpurun_out/r02_bench26_n2.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench26_n2.json').read().strip().splitlines()[-1])
print('value',d['value'],'ms',d['ms_per_step'],d['scaling'],'e2e',d['e2e']['value'],'render',d['e2e']['render_api']['value'])
print('parity',d['parity'])
PY
