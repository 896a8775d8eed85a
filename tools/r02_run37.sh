#!/bin/bash
# round-2 GPU call 37: end-of-round check at HEAD: full GPU suite, smoke, default bench line, reference arm
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/r02_t37.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_t37.log
tail -4 gpurun_out/r02_t37.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r02_smoke37.log 2>&1; tail -2 gpurun_out/r02_smoke37.log
timeout 1500 python bench.py > gpurun_out/r02_bench37.json 2> gpurun_out/r02_bench37.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench37.json').read().strip().splitlines()[-1])
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'compact',d['e2e']['compact_records'].get('value'),'render',d['e2e']['render_api']['value'])
print('extras',d['extras']['occlusion_any_hit'].get('value'))
print('roofline',d['roofline']['bound'],d['roofline']['frac'],d['roofline']['fractions'], d['roofline']['per_launch_kind'])
print('cpu',d['cpu_baseline']['value'],d['cpu_baseline']['cores'],'parity',d['parity']['ok'],'clocks',d['clocks']['sm_mhz'],d['clocks']['reasons'])
for c in d['configs']: print(c.get('name'),c.get('value'),c.get('ms_per_step'),c.get('build_ms',{}).get('device_best_of_3'),c.get('parity_ok'), c.get('error'))
PY
