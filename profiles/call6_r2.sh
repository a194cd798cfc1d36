#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r2f_build.log 2>&1
timeout 240 python tests/attn_child.py --bench > gpurun_out/r2f_attn.jsonl 2> gpurun_out/r2f_attn.log
echo "attn rc=$?"; cat gpurun_out/r2f_attn.jsonl | cut -c1-300; tail -3 gpurun_out/r2f_attn.log
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/r2f_gpu_tests.log 2>&1
echo "pytest rc=$?"; tail -5 gpurun_out/r2f_gpu_tests.log | cut -c1-300
for tag in mt2 mt1; do
  case $tag in
    mt2) envs="";;
    mt1) envs="ESB200_SPCONV_MT1=1";;
  esac
  env $envs timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r2f_bench_$tag.json 2> gpurun_out/r2f_bench_$tag.log
  echo "bench $tag rc=$?"; grep "host ms" gpurun_out/r2f_bench_$tag.log | cut -c1-260
done
python - <<'PY'
import json
for tag in ('mt2', 'mt1'):
    try:
        d = json.loads(open(f'gpurun_out/r2f_bench_{tag}.json').read().strip().splitlines()[-1])
        r = d['roofline']
        print(tag, 'value', round(d['value'], 2), 'ms', round(d['ms_per_step'], 2), 'e2e', round(d['e2e']['value'], 2), 'fwd frac', round(r['frac'], 3),
              'wgrad frac', round(r['wgrad_frac'], 3), 'agg', round(r['aggregate_frac'], 3), 'comp', round(r['compulsory_frac'], 3),
              'pass ms', round(r['pass_ms_per_step'], 1), 'launches/step', d['gpu_launches'] / d['steps'])
    except Exception as e:
        print(tag, 'failed:', e)
PY
for tag in own lib; do
  case $tag in
    own) envs="";;
    lib) envs="ESB200_ATTN=lib";;
  esac
  env $envs ESB200_TEXT_RANDOM_INIT=1 timeout 300 python bench.py --variant C4 --steps 5 --warmup 2 --no-cpu-baseline --no-e2e > gpurun_out/r2f_bench_c4_$tag.json 2> gpurun_out/r2f_bench_c4_$tag.log
  echo "C4 $tag rc=$?"; tail -2 gpurun_out/r2f_bench_c4_$tag.log | cut -c1-200
  python -c "
import json
d=json.loads(open('gpurun_out/r2f_bench_c4_$tag.json').read().strip().splitlines()[-1]); print('C4 $tag', round(d['value'],2), 'scans/s', round(d['ms_per_step'],1), 'ms/step')"
done
