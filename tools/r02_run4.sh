#!/bin/bash
# round-2 GPU call 4: new bench.py (N=1, all configs), GPU tests, objrender timing
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -s -k "objrender" > gpurun_out/r02_t4_objrender.log 2>&1
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r02_t4.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_t4.log
( time timeout 1200 python bench.py --steps 20 --warmup 3 > gpurun_out/r02_bench4.json 2> gpurun_out/r02_bench4.err ) 2> gpurun_out/r02_bench4.time
tail -4 gpurun_out/r02_t4.log; grep objrender gpurun_out/r02_t4_objrender.log; cat gpurun_out/r02_bench4.time; tail -5 gpurun_out/r02_bench4.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench4.json').read().strip().splitlines()[-1])
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'sustained',d['sustained'])
print('parity',json.dumps(d['parity'])[:1500])
print('roofline',json.dumps({k:v for k,v in d['roofline'].items() if k not in ('kernel','algorithmic_note','peak_source')})[:2500])
print('cpu',json.dumps(d['cpu_baseline'])[:1200])
for c in d['configs']: print('CONFIG',json.dumps(c)[:1800])
PY
