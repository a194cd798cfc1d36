#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r2s_build.log 2>&1
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2s_gpu_tests.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/r2s_gpu_tests.log | cut -c1-300
timeout 120 python __graft_entry__.py smoke > gpurun_out/r2s_smoke.log 2>&1
echo "smoke rc=$?"; tail -1 gpurun_out/r2s_smoke.log | cut -c1-250
timeout 500 python bench.py > gpurun_out/r2s_bench.json 2> gpurun_out/r2s_bench.log
echo "bench rc=$?"; grep -E "host ms per timed|cpu baseline" gpurun_out/r2s_bench.log | cut -c1-300
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2s_bench.json').read().strip().splitlines()[-1])
r = d['roofline']
print('C2 value', round(d['value'], 2), 'ms', round(d['ms_per_step'], 2), 'e2e', round(d['e2e']['value'], 2), 'fwd frac', round(r['frac'], 3),
      'wgrad frac', round(r['wgrad_frac'], 3), 'agg', round(r['aggregate_frac'], 3), 'launches/step', d['gpu_launches'] / d['steps'], 'cpu', d.get('cpu_baseline', {}).get('value'), d.get('cpu_baseline', {}).get('cores'))
PY
