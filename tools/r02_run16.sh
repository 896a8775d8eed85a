#!/bin/bash
# round-2 GPU call 16 (8 GPUs): bench.py under torchrun at N = 8 (strong scaling of 1920x1080x64 spp + configs[4]), the C++ example on 8 GPUs
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02_smi16.txt 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/r02_bench16_n8.json 2> gpurun_out/r02_bench16_n8.err; echo "rc=$?" >> gpurun_out/r02_bench16_n8.err
( cd examples && timeout 300 ./bin/multi_gpu_ao 8 1920 1080 64 10 ) > gpurun_out/r02_cpp_multi16.log 2>&1; echo "rc=$?" >> gpurun_out/r02_cpp_multi16.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 4 --steps 20 --warmup 3 --no-configs > gpurun_out/r02_bench16_n4.json 2> gpurun_out/r02_bench16_n4.err; echo "rc=$?" >> gpurun_out/r02_bench16_n4.err
tail -3 gpurun_out/r02_bench16_n8.err; cat gpurun_out/r02_cpp_multi16.log | tail -4
python - <<'PY'
import json
for f in ('gpurun_out/r02_bench16_n8.json','gpurun_out/r02_bench16_n4.json'):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f,'value',d['value'],'ms',d['ms_per_step'],'scaling',d['scaling'],'e2e',d['e2e']['value'],'render_api',d['e2e'].get('render_api',{}).get('value'),'parity',d['parity']['ok'])
        for c in d.get('configs',[]): print('  CONFIG', json.dumps(c)[:1400])
    except Exception as e: print(f,'ERR',e)
PY
