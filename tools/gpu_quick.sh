#!/bin/bash
# Quick GPU check: parity tests + one bench line (no CPU baseline) + optional extras given as arguments:
#   ncu   : full capture of k_conv_fused on a sparse step     micro : FFMA / FFMA2 issue-rate micro-benchmark
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
tail -c 2500 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
for a in "$@"; do
  case $a in
    micro) ./tools/microbench/ffma2_bench > gpurun_out/ffma2_bench.txt 2>&1; cat gpurun_out/ffma2_bench.txt ;;
    ncu) timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_conv_fused|k_acc_tc|k_edge_hidden' -s 6 -c 4 \
           -o gpurun_out/prof_fused_sparse -f python tools/profile_step.py --complexes 2 --rev-steps 1 --start-step 14 > gpurun_out/prof_sparse.log 2>&1
         tail -2 gpurun_out/prof_sparse.log ;;
    launches) timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches.csv \
           python tools/profile_step.py --complexes 2 --rev-steps 2 > gpurun_out/launches.log 2>&1 ;;
  esac
done
