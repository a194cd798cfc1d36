#!/bin/bash
# Second gpurun call of round 2: full -m gpu suite on the own 2D kernels, the launch list and the per-kernel metric table
# (DRAM bytes, tensor pipe, L2 hit) of one C2 training step.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r2b_build.log 2>&1
timeout 900 python -m pytest tests -q -m gpu --junitxml gpurun_out/r2b_gpu_tests.xml > gpurun_out/r2b_gpu_tests.log 2>&1
echo "pytest rc=$?"; tail -8 gpurun_out/r2b_gpu_tests.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2b_bench_own.json 2> gpurun_out/r2b_bench_own.log
tail -c 1500 gpurun_out/r2b_bench_own.json; echo
export ESB_CUDA_PROFILER_RANGE=1
timeout 600 ncu --profile-from-start off --clock-control none --csv --log-file gpurun_out/r2b_metrics_own.csv \
  --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,lts__t_sector_hit_rate.pct,sm__warps_active.avg.pct_of_peak_sustained_active \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2b_ncu_metrics.log 2>&1
python profiles/summarize_metrics.py gpurun_out/r2b_metrics_own.csv 70
