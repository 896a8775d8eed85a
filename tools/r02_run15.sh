#!/bin/bash
# round-2 GPU call 15: validation of the round's state -- full GPU suite, initcheck re-run, bench line (all configs), reference arm,
# launch list of the bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02_t15.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_t15.log
timeout 1500 compute-sanitizer --tool initcheck --error-exitcode 3 python tools/sanitize_target.py > gpurun_out/r02_san_initcheck.log 2>&1; echo "initcheck rc=$?" >> gpurun_out/r02_san_initcheck.log
( time timeout 1500 python bench.py --steps 20 --warmup 3 > gpurun_out/r02_bench15.json 2> gpurun_out/r02_bench15.err ) 2> gpurun_out/r02_bench15.time
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02_bench15_ref.json 2> gpurun_out/r02_bench15_ref.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches15.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-configs > gpurun_out/r02_launches15.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke15.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r02_smoke15.log
tail -4 gpurun_out/r02_t15.log; tail -3 gpurun_out/r02_san_initcheck.log; cat gpurun_out/r02_bench15.time; tail -2 gpurun_out/r02_smoke15.log
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench15.json').read().strip().splitlines()[-1])
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'parity ok',d['parity']['ok'], 'roof', d['roofline']['bound'], d['roofline']['frac'])
for c in d.get('configs',[]): print('CONFIG',c.get('name'),c.get('value'),c.get('ms_per_step'),c.get('build_ms',{}).get('device_best_of_3') if isinstance(c.get('build_ms'),dict) else '', c.get('parity_ok'), c.get('error'))
PY
