#!/bin/bash
# Quick GPU check: parity tests + one bench line (no ncu).
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
tail -c 2500 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
