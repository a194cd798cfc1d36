#!/bin/bash
# Third gpurun call of round 2: first runs of the stem tcgen05 kernel and the TMA-fed sparse conv, the CUDA-graphed 2D branch,
# bench A/B and a time-only launch list of one step.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r2c_build.log 2>&1
timeout 200 python tests/conv_tma_child.py 0 2 > gpurun_out/r2c_conv_tma_cases.jsonl 2> gpurun_out/r2c_conv_tma_cases.log
echo "conv_tma+stem rc=$?"; grep -E 'stem|error' gpurun_out/r2c_conv_tma_cases.jsonl; tail -2 gpurun_out/r2c_conv_tma_cases.log
timeout 300 python tests/spconv_tma_child.py --bench > gpurun_out/r2c_spconv_tma.jsonl 2> gpurun_out/r2c_spconv_tma.log
echo "spconv_tma rc=$?"; cat gpurun_out/r2c_spconv_tma.jsonl | cut -c1-260; tail -3 gpurun_out/r2c_spconv_tma.log
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/r2c_gpu_tests.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/r2c_gpu_tests.log
for tag in graph nograph tma; do
  case $tag in
    graph) envs="";;
    nograph) envs="ESB200_GRAPH2D=0";;
    tma) envs="ESB200_SPCONV=tma";;
  esac
  env $envs timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2c_bench_$tag.json 2> gpurun_out/r2c_bench_$tag.log
  echo "bench $tag rc=$?"; tail -2 gpurun_out/r2c_bench_$tag.log | cut -c1-300
done
python - <<'PY'
import json
for tag in ('graph', 'nograph', 'tma'):
    try:
        d = json.loads(open(f'gpurun_out/r2c_bench_{tag}.json').read().strip().splitlines()[-1])
        r = d['roofline']
        print(tag, 'value', round(d['value'], 2), 'ms', round(d['ms_per_step'], 2), 'e2e', round(d['e2e']['value'], 2), 'fwd frac', round(r['frac'], 3),
              'wgrad frac', round(r['wgrad_frac'], 3), 'agg', round(r['aggregate_frac'], 3), 'comp', round(r['compulsory_frac'], 3),
              'pass ms', round(r['pass_ms_per_step'], 1), 'launches/step', d['gpu_launches'] / d['steps'])
    except Exception as e:
        print(tag, 'failed:', e)
PY
export ESB_CUDA_PROFILER_RANGE=1
timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/r2c_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2c_ncu.log 2>&1
python profiles/summarize_launches.py gpurun_out/r2c_launches.csv 45 | cut -c1-150
