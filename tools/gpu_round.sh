#!/bin/bash
# parity tests, bench (ours + reference arm), then the ncu launch list of bench.py itself with DRAM bytes per launch
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader > gpurun_out/gpu.txt
[ -x tools/microbench/umma_pacing ] && timeout 120 tools/microbench/umma_pacing > gpurun_out/umma_pacing.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -25 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
tail -c 4500 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
if [ "$1" != "quick" ]; then
  timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err
  cat gpurun_out/bench_ref.json
  timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 2400 --csv \
      --log-file gpurun_out/launches_bench.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-sparse > gpurun_out/launches_bench.log 2>&1
  python tools/launch_summary.py gpurun_out/launches_bench.csv --traffic-json gpurun_out/conv_lv3_traffic.json | tee gpurun_out/launches_bench_summary.txt
fi
