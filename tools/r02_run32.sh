#!/bin/bash
# round-2 GPU call 32: final policies: full GPU suite, ncu --set full of the two traversal launches, launch list, bench line
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -x > gpurun_out/r02_t32.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_t32.log
tail -4 gpurun_out/r02_t32.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:traverse_fast3 -s 2 -c 2 -f -o gpurun_out/r02_trav32 python tools/profile_target.py > gpurun_out/r02_ncu32.log 2>&1
tail -2 gpurun_out/r02_ncu32.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches32.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-configs > gpurun_out/r02_launches32.log 2>&1
timeout 1500 python bench.py > gpurun_out/r02_bench32.json 2> gpurun_out/r02_bench32.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench32.json').read().strip().splitlines()[-1])
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'compact',d['e2e']['compact_records'].get('value'),'render',d['e2e']['render_api']['value'])
print('extras',d['extras']['occlusion_any_hit'].get('value'))
print('roofline',d['roofline']['bound'],d['roofline']['frac'],d['roofline']['fractions'])
print('parity',d['parity']['ok'])
for c in d['configs']: print(c.get('name'),c.get('value'),c.get('ms_per_step'),c.get('build_ms',{}).get('device_best_of_3'),c.get('parity_ok'), c.get('error'))
PY
