#!/bin/bash
# round-2 GPU call 10 (2 GPUs): the N>1 paths -- Python gather test, the C++ fork-per-GPU example over the C-ABI (NCCL),
# bench.py under torchrun (strong scaling = default), the reference arm under torchrun
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02_smi10.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_multi.py -q > gpurun_out/r02_t10.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_t10.log
( cd examples && timeout 300 ./bin/multi_gpu_ao 2 1920 1080 16 5 ) > gpurun_out/r02_cpp_multi10.log 2>&1; echo "rc=$?" >> gpurun_out/r02_cpp_multi10.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r02_bench10_n2.json 2> gpurun_out/r02_bench10_n2.err; echo "bench rc=$?" >> gpurun_out/r02_bench10_n2.err
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 3 --scaling weak --no-cpu-baseline --no-configs > gpurun_out/r02_bench10_n2_weak.json 2> gpurun_out/r02_bench10_n2_weak.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/r02_bench10_n2_ref.json 2> gpurun_out/r02_bench10_n2_ref.err
tail -3 gpurun_out/r02_t10.log; tail -8 gpurun_out/r02_cpp_multi10.log; tail -5 gpurun_out/r02_bench10_n2.err
python - <<'PY'
import json
for f in ('gpurun_out/r02_bench10_n2.json','gpurun_out/r02_bench10_n2_weak.json','gpurun_out/r02_bench10_n2_ref.json'):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'value',d.get('value'),'ms',d.get('ms_per_step'),'scaling',d.get('scaling'),'e2e',(d.get('e2e') or {}).get('value'), 'parity', json.dumps(d.get('parity'))[:600])
        for c in d.get('configs',[]): print('  CONFIG', json.dumps(c)[:700])
    except Exception as e: print(f, 'ERR', e)
PY
