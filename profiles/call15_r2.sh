#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r2q_build.log 2>&1
export ESB_CUDA_PROFILER_RANGE=1 ESB200_GRAPH2D=0
timeout 420 ncu --profile-from-start off --set full --clock-control none --import-source on \
  -k 'regex:spconv_tc_fwd_kernel|spconv_tc_wgrad_kernel' -s 6 -c 14 \
  -o gpurun_out/r2_spconv_kernels python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2q_A.log 2>&1
echo "spconv rc=$?"
ncu -i gpurun_out/r2_spconv_kernels.ncu-rep --page raw --csv 2>/dev/null | python profiles/ncu_raw_pick.py > gpurun_out/r2_spconv_kernels_summary.txt; cat gpurun_out/r2_spconv_kernels_summary.txt | cut -c1-330
unset ESB_CUDA_PROFILER_RANGE
timeout 200 ncu --set full --clock-control none --import-source on -k 'regex:attn_fwd_kernel|attn_bwd_kernel' -s 10 -c 4 \
  -o gpurun_out/r2_attn_kernels python tests/attn_child.py --bench > gpurun_out/r2q_B.log 2>&1
echo "attn rc=$?"
ncu -i gpurun_out/r2_attn_kernels.ncu-rep --page raw --csv 2>/dev/null | python profiles/ncu_raw_pick.py > gpurun_out/r2_attn_kernels_summary.txt; cat gpurun_out/r2_attn_kernels_summary.txt | cut -c1-330
