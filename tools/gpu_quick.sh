#!/bin/bash
# parity tests of the conv-kernel paths, then (if green) the bench; every step under its own timeout
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 240 --timeout-method thread -k "conv_kernel_paths or forward_stages" > gpurun_out/pytest_quick.log 2>&1; rc=$?
echo "quick rc=$rc" >> gpurun_out/pytest_quick.log; tail -30 gpurun_out/pytest_quick.log
if [ $rc -eq 0 ]; then
  timeout 1500 python -m pytest tests -m gpu -q --timeout 300 --timeout-method thread > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
  tail -30 gpurun_out/pytest_gpu.log
  timeout 900 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
  tail -c 3000 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
fi
