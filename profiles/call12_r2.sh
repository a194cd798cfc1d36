#!/bin/bash
mkdir -p gpurun_out
ESB_CAPTURE_SKIP_C=1 bash profiles/capture_r2.sh
timeout 600 python -m pytest tests/test_model_gpu.py -q -s -k "c2_shaped" > gpurun_out/r2l_c2_parity.log 2>&1
echo "c2 parity rc=$?"; grep -E "C2-shaped|passed|failed|Error" gpurun_out/r2l_c2_parity.log | cut -c1-1200
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2l_smoke.log 2>&1
echo "smoke rc=$?"; tail -2 gpurun_out/r2l_smoke.log | cut -c1-300
