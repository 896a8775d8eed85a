#!/bin/bash
# round-2 GPU call 20: compact ray records + any-hit occlusion rays: tests, bench line
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_traverse.py tests/test_gpu_render.py tests/test_gpu_path.py -q -x > gpurun_out/r02_t20.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_t20.log
tail -15 gpurun_out/r02_t20.log
timeout 1200 python bench.py --no-configs > gpurun_out/r02_bench20.json 2> gpurun_out/r02_bench20.err; echo "bench rc=$?"
tail -5 gpurun_out/r02_bench20.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench20.json').read().strip().splitlines()[-1])
print('value',d['value'],'ms',d['ms_per_step'])
print('e2e',d['e2e']['value'],d['e2e'].get('compact_records'))
print('extras',d.get('extras'))
print('roofline frac',d['roofline']['frac'],d['roofline']['per_launch_kind'])
print('parity ok',d['parity']['ok'])
PY
