#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r2e_build.log 2>&1
timeout 300 python tests/conv_tma_child.py --bench > gpurun_out/r2e_conv_tma_cases.jsonl 2> gpurun_out/r2e_conv_tma_cases.log
echo "conv_tma rc=$? ok=$(grep -c '"ok": true' gpurun_out/r2e_conv_tma_cases.jsonl) bad=$(grep -c '"ok": false' gpurun_out/r2e_conv_tma_cases.jsonl)"
grep -E '"ok": false|error' gpurun_out/r2e_conv_tma_cases.jsonl | head; grep bench gpurun_out/r2e_conv_tma_cases.jsonl | cut -c1-330; tail -3 gpurun_out/r2e_conv_tma_cases.log
if grep -q '"ok": false' gpurun_out/r2e_conv_tma_cases.jsonl; then echo "conv_tma failures: stopping"; exit 0; fi
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/r2e_gpu_tests.log 2>&1
echo "pytest rc=$?"; tail -6 gpurun_out/r2e_gpu_tests.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.log
echo "bench rc=$?"; grep "host ms" gpurun_out/r2e_bench.log | cut -c1-200
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2e_bench.json').read().strip().splitlines()[-1])
r = d['roofline']
print('value', round(d['value'], 2), 'ms', round(d['ms_per_step'], 2), 'e2e', round(d['e2e']['value'], 2), 'fwd frac', round(r['frac'], 3),
      'wgrad frac', round(r['wgrad_frac'], 3), 'agg', round(r['aggregate_frac'], 3), 'pass ms', round(r['pass_ms_per_step'], 1),
      'launches/step', d['gpu_launches'] / d['steps'])
PY
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --torch-profile gpurun_out/r2e_torch_profile.txt > /dev/null 2> gpurun_out/r2e_prof.log
head -45 gpurun_out/r2e_torch_profile.txt | cut -c1-70,128-200
