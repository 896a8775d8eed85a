#!/bin/bash
# round-2 GPU call 6 (after container re-creation): full GPU suite, bench with all configs, launch list, ncu of the traversal kernel
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r02_smi.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02_t6.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_t6.log
( time timeout 1200 python bench.py --steps 20 --warmup 3 > gpurun_out/r02_bench6.json 2> gpurun_out/r02_bench6.err ) 2> gpurun_out/r02_bench6.time
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02_bench6_ref.json 2> gpurun_out/r02_bench6_ref.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:traverse_fast3 -s 2 -c 2 -f -o gpurun_out/r02_trav6 python tools/profile_target.py > gpurun_out/r02_ncu6.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches6.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-configs > gpurun_out/r02_launches6.log 2>&1
tail -5 gpurun_out/r02_t6.log; cat gpurun_out/r02_bench6.time; tail -5 gpurun_out/r02_bench6.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench6.json').read().strip().splitlines()[-1])
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'sustained',d.get('sustained'))
print('parity',json.dumps(d.get('parity'))[:1500])
print('roofline',json.dumps({k:v for k,v in d['roofline'].items() if k not in ('kernel','algorithmic_note','peak_source')})[:2500])
print('cpu',json.dumps(d['cpu_baseline'])[:1200])
for c in d.get('configs',[]): print('CONFIG',json.dumps(c)[:1800])
PY
