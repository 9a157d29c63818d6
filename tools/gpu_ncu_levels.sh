#!/bin/bash
# ncu full capture of the four fused conv levels on a mid-trajectory (sparse) step of 80 poses.
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_conv_fused' -c 5 \
    -o gpurun_out/prof_levels -f python tools/profile_step.py --complexes 2 --rev-steps 1 --start-step 12 > gpurun_out/prof_levels.log 2>&1
tail -n 3 gpurun_out/prof_levels.log
