#!/bin/bash
# round-2 GPU call 23: full GPU suite, smoke, full bench line (all configs), reference arm
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -x > gpurun_out/r02_t23.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_t23.log
tail -6 gpurun_out/r02_t23.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r02_smoke23.log 2>&1; tail -2 gpurun_out/r02_smoke23.log
timeout 1500 python bench.py > gpurun_out/r02_bench23.json 2> gpurun_out/r02_bench23.err; echo "bench rc=$?"
timeout 900 python bench.py --impl reference > gpurun_out/r02_bench23_ref.json 2> gpurun_out/r02_bench23_ref.err; echo "ref rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench23.json').read().strip().splitlines()[-1])
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'compact',d['e2e']['compact_records'].get('value'),'render',d['e2e']['render_api']['value'])
print('extras',d['extras']['occlusion_any_hit'].get('value'))
print('roofline',d['roofline']['bound'],d['roofline']['frac'],d['roofline']['fractions'])
print('cpu',d['cpu_baseline']['value'],d['cpu_baseline']['cores'])
print('parity',d['parity']['ok'])
for c in d['configs']: print(c.get('name'),c.get('value'),c.get('ms_per_step'),c.get('build_ms',{}).get('device_best_of_3'),c.get('parity_ok'), c.get('error'))
r=json.loads(open('gpurun_out/r02_bench23_ref.json').read().strip().splitlines()[-1])
print('ref',r['value'],r['cpu_baseline'])
PY
