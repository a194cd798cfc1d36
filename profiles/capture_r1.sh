#!/bin/bash
# Round-1 ncu captures (run under gpurun from the repo root). Numbers printed under ncu are never bench values.
mkdir -p gpurun_out
export ESB_CUDA_PROFILER_RANGE=1
BENCH="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e"
# full metric set for the dominant kernel: head-level sparse conv, forward (MN-major B) and dgrad (K-major B)
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on --kernel-name-base demangled \
  -k 'regex:spconv_tc_fwd_kernel<.int.128' -c 4 -f -o gpurun_out/prof_spconv_tc_fwd_r1 $BENCH > gpurun_out/ncu_full_fwd.log 2>&1
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on --kernel-name-base demangled \
  -k 'regex:spconv_tc_wgrad_kernel<.int.128' -s 2 -c 3 -f -o gpurun_out/prof_spconv_tc_wgrad_r1 $BENCH > gpurun_out/ncu_full_wgrad.log 2>&1
ls -la gpurun_out/*.ncu-rep
