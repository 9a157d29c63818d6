#!/bin/bash
# per-warp-role cycle trace of k_conv_tcr (trace build on the GPU box), then an ncu full capture of one dense level-3 launch
mkdir -p gpurun_out
cp disco_diffdock_b200/libddk.so /tmp/libddk_keep.so
DDK_NVCC_EXTRA=-DDDK_TCR_TRACE=1 python -m disco_diffdock_b200.build --force > /dev/null 2>&1
timeout 300 python tools/profile_step.py --complexes 2 --rev-steps 1 --pocket 2> gpurun_out/tcr_trace.txt | tail -1
cat gpurun_out/tcr_trace.txt | cut -c1-700
cp /tmp/libddk_keep.so disco_diffdock_b200/libddk.so
if [ "$1" == "ncu" ]; then
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_conv_tcr' -s 3 -c 1 \
    -o gpurun_out/prof_tcr_dense -f python tools/profile_step.py --complexes 2 --rev-steps 1 --pocket > gpurun_out/prof_tcr_dense.log 2>&1
tail -2 gpurun_out/prof_tcr_dense.log
fi
