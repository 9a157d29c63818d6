#!/bin/bash
# ncu full capture of the fused conv kernel (dense first step and a sparse late step) + launch list.
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_conv_fused' -s 3 -c 2 \
    -o gpurun_out/prof_fused_dense -f python tools/profile_step.py --complexes 2 --rev-steps 1 > gpurun_out/prof_dense.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_conv_fused' -s 3 -c 2 \
    -o gpurun_out/prof_fused_sparse -f python tools/profile_step.py --complexes 2 --rev-steps 1 --start-step 14 > gpurun_out/prof_sparse.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_fused.csv \
    python tools/profile_step.py --complexes 2 --rev-steps 2 > gpurun_out/launches_fused.log 2>&1
tail -3 gpurun_out/prof_dense.log gpurun_out/prof_sparse.log
ls -la gpurun_out | tail -8
