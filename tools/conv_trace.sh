#!/bin/bash
# Per-CTA timeline summary of k_conv_fused (DDK_CONV_TRACE build): how long the CTAs of a launch are resident, how much of
# that is work / weight-slice reloads / task claiming, at 40 and 400 poses per launch, dense (t = 1) and sparse (step 14).
# Rebuilds the library with the trace counters, runs, then restores the product build.
mkdir -p gpurun_out
DDK_NVCC_EXTRA=-DDDK_CONV_TRACE=1 python -m disco_diffdock_b200.build --force > /dev/null || exit 1
for cx in 1 10; do
  for st in 0 14; do
    echo "== complexes $cx start-step $st" >> gpurun_out/conv_trace.txt
    DDK_CONV_TRACE=1 timeout 300 python tools/profile_step.py --complexes $cx --rev-steps 1 --start-step $st 2>&1 | grep -v "^$" >> gpurun_out/conv_trace.txt
  done
done
python -m disco_diffdock_b200.build --force > /dev/null
cat gpurun_out/conv_trace.txt
