#!/bin/bash
# One gpurun call: GPU parity tests, bench lines (ours + reference arm), ncu launch list and one full capture of the fused
# conv kernel on a dense (t = 1) and a sparse (late) step.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
tail -c 3000 gpurun_out/bench.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches.csv \
    python tools/profile_step.py --complexes 2 --rev-steps 2 > gpurun_out/launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_conv_fused|k_acc_tc|k_edge_hidden' -s 6 -c 4 \
    -o gpurun_out/prof_fused_dense -f python tools/profile_step.py --complexes 2 --rev-steps 1 > gpurun_out/prof_dense.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_conv_fused|k_acc_tc|k_edge_hidden' -s 6 -c 4 \
    -o gpurun_out/prof_fused_sparse -f python tools/profile_step.py --complexes 2 --rev-steps 1 --start-step 14 > gpurun_out/prof_sparse.log 2>&1
ls -la gpurun_out
