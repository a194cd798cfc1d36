#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r2h_build.log 2>&1
timeout 300 python tests/conv_tma_child.py > gpurun_out/r2h_conv_tma_cases.jsonl 2> gpurun_out/r2h_conv_tma_cases.log
echo "conv_tma rc=$? ok=$(grep -c '"ok": true' gpurun_out/r2h_conv_tma_cases.jsonl) bad=$(grep -c '"ok": false' gpurun_out/r2h_conv_tma_cases.jsonl)"
grep -E 'conv3d|"ok": false|error' gpurun_out/r2h_conv_tma_cases.jsonl | cut -c1-420; tail -3 gpurun_out/r2h_conv_tma_cases.log
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/r2h_gpu_tests.log 2>&1
echo "pytest rc=$?"; tail -6 gpurun_out/r2h_gpu_tests.log | cut -c1-300
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.log
echo "bench rc=$?"; grep "host ms" gpurun_out/r2h_bench.log | cut -c1-260
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2h_bench.json').read().strip().splitlines()[-1])
r = d['roofline']
print('C2 value', round(d['value'], 2), 'ms', round(d['ms_per_step'], 2), 'e2e', round(d['e2e']['value'], 2), 'fwd frac', round(r['frac'], 3),
      'wgrad frac', round(r['wgrad_frac'], 3), 'agg', round(r['aggregate_frac'], 3), 'launches/step', d['gpu_launches'] / d['steps'])
PY
for tag in own lib; do
  case $tag in
    own) envs="";;
    lib) envs="ESB200_CONV3D=lib ESB200_CONV2D=cudnn";;
  esac
  env $envs timeout 400 python bench.py --variant C3 --steps 8 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2h_bench_c3_$tag.json 2> gpurun_out/r2h_bench_c3_$tag.log
  echo "C3 $tag rc=$?"; tail -2 gpurun_out/r2h_bench_c3_$tag.log | cut -c1-250
  python -c "
import json
d=json.loads(open('gpurun_out/r2h_bench_c3_$tag.json').read().strip().splitlines()[-1]); print('C3 $tag', round(d['value'],2), 'scans/s', round(d['ms_per_step'],1), 'ms/step', d['gpu_launches']/d['steps'], 'launches/step')"
done
timeout 400 python bench.py --variant C5 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2h_bench_c5.json 2> gpurun_out/r2h_bench_c5.log
echo "C5 rc=$?"; tail -2 gpurun_out/r2h_bench_c5.log | cut -c1-250
python - <<'PY'
import json
try:
    d = json.loads(open('gpurun_out/r2h_bench_c5.json').read().strip().splitlines()[-1])
    r = d['roofline']
    print('C5 value', round(d['value'], 2), 'ms', round(d['ms_per_step'], 2), 'fwd frac', round(r['frac'], 3), 'wgrad frac', round(r['wgrad_frac'], 3),
          'agg', round(r['aggregate_frac'], 3), 'comp', round(r['compulsory_frac'], 3))
except Exception as e:
    print('C5 failed', e)
PY
