#!/bin/bash
# cycle breakdown of k_acc_tc (DDK_TC_TRACE build): dense step at 80 poses
mkdir -p gpurun_out
DDK_NVCC_EXTRA=-DDDK_TC_TRACE=1 python -m disco_diffdock_b200.build --force > /dev/null || exit 1
timeout 300 python tools/profile_step.py --complexes 2 --rev-steps 1 2>&1 | grep -E "tc_trace|edges" | tee gpurun_out/tc_trace.txt
python -m disco_diffdock_b200.build --force > /dev/null
