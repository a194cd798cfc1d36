#!/bin/bash
# ncu evidence of round 2 (one B200): A. full-set captures of the tensor-core kernels (.ncu-rep, read with `ncu -i ... --page raw`)
# B. DRAM bytes / tensor pipe / L2 hit of EVERY hand-written kernel of one C2 training step (4 scans), C. the kernels that are
# not in the training step (attention, NMS, unprojection, 9-DoF IoU) from the predict path and the attention child.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r2p_build.log 2>&1
OWN='regex:^(void )?(\(anonymous namespace\)::)?(spconv|conv_tma|conv2d|stem7x7|attn_|paint|hash_|kernel_map|bn_|norm_|focal|seg_|assign|count_|best_level|topk|segmented|nms|unproject|depth_|adamw|gather2|img_normalize|maxpool|act_|bias_act|interp|pair_|tile_mask|voxelize|compact|flag_|inverse|cast_|sumsq|clip|bbox_cd|generative|iou|box3d|hungarian)'
METRICS=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,lts__t_sector_hit_rate.pct,sm__warps_active.avg.pct_of_peak_sustained_active
export ESB_CUDA_PROFILER_RANGE=1
# the 2D branch runs eagerly here: ncu's multi-pass kernel replay fails (LaunchFailed) on the kernel nodes of the replayed CUDA graphs
export ESB200_GRAPH2D=0
# A
timeout 500 ncu --profile-from-start off --set full --clock-control none --import-source on \
  -k 'regex:spconv_tc_fwd_kernel|spconv_tc_wgrad_kernel|conv_tma_kernel|conv_tma_wgrad_kernel|stem7x7_tc_kernel' -c 40 \
  -o gpurun_out/r2_top_kernels python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2p_A.log 2>&1
echo "A rc=$?"; ls -la gpurun_out/r2_top_kernels.ncu-rep
ncu -i gpurun_out/r2_top_kernels.ncu-rep --page raw --csv 2>/dev/null | python profiles/ncu_raw_pick.py > gpurun_out/r2_top_kernels_summary.txt; head -30 gpurun_out/r2_top_kernels_summary.txt | cut -c1-220
# B
timeout 900 ncu --profile-from-start off --clock-control none --csv --log-file gpurun_out/r2_step_metrics.csv -k "$OWN" --metrics $METRICS \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2p_B.log 2>&1
echo "B rc=$?"; python profiles/summarize_metrics.py gpurun_out/r2_step_metrics.csv 90 > gpurun_out/r2_step_kernel_metrics.txt; head -60 gpurun_out/r2_step_kernel_metrics.txt | cut -c1-200
# C
unset ESB_CUDA_PROFILER_RANGE
if [ "$ESB_CAPTURE_SKIP_C" = "1" ]; then exit 0; fi
timeout 400 ncu --clock-control none --csv --log-file gpurun_out/r2_attn_metrics.csv -k "$OWN" --metrics $METRICS python tests/attn_child.py --bench > gpurun_out/r2p_C1.log 2>&1
python profiles/summarize_metrics.py gpurun_out/r2_attn_metrics.csv 10 | cut -c1-200
timeout 400 ncu --clock-control none --csv --log-file gpurun_out/r2_predict_metrics.csv -k "$OWN" --metrics $METRICS python profiles/predict_once.py > gpurun_out/r2p_C2.log 2>&1
python profiles/summarize_metrics.py gpurun_out/r2_predict_metrics.csv 60 | grep -E "nms|unproject|depth_|iou|box3d|topk|paint|interp|voxelize|img_normalize|# " | cut -c1-200
