#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r2m_build.log 2>&1
timeout 1200 python -m pytest tests -q -m gpu -s -k "c2_shaped" > gpurun_out/r2m_c2.log 2>&1
echo "c2 rc=$?"; grep -E "C2-shaped|passed|failed" gpurun_out/r2m_c2.log | cut -c1-1300
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r2m_gpu_tests.log 2>&1
echo "pytest rc=$?"; tail -5 gpurun_out/r2m_gpu_tests.log | cut -c1-300
timeout 400 python bench.py --no-cpu-baseline > gpurun_out/r2m_bench.json 2> gpurun_out/r2m_bench.log
echo "bench rc=$?"; grep -E "host ms per timed|gc collections" gpurun_out/r2m_bench.log | cut -c1-400
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2m_bench.json').read().strip().splitlines()[-1])
r = d['roofline']
print('C2 value', round(d['value'], 2), 'ms', round(d['ms_per_step'], 2), 'e2e', round(d['e2e']['value'], 2), 'fwd frac', round(r['frac'], 3),
      'wgrad frac', round(r['wgrad_frac'], 3), 'agg', round(r['aggregate_frac'], 3), 'comp', round(r['compulsory_frac'], 3), 'traffic', r['traffic'],
      'launches/step', d['gpu_launches'] / d['steps'], d['clocks'])
PY
