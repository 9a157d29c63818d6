#!/bin/bash
# k_conv_tcr (DDK_TC=2): parity tests of the conv paths, cycle trace, dense bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 240 --timeout-method thread -k "conv_kernel_paths or forward_stages or truncation or large_receptor" > gpurun_out/pytest_quick.log 2>&1; rc=$?
echo "quick rc=$rc" >> gpurun_out/pytest_quick.log; tail -4 gpurun_out/pytest_quick.log
if [ $rc -eq 0 ]; then
  export DDK_TC=2
  bash tools/tcr_trace.sh $1 2>&1 | grep -E "layer 3|Report" | cut -c1-1100
  timeout 900 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-sparse > gpurun_out/bench_tc2.json 2> gpurun_out/bench.err; echo "bench rc=$?"
  python -c "
import json; d=json.load(open('gpurun_out/bench_tc2.json')); print('TC2 value', d['value'], 'e2e', d['e2e']['value'], d['e2e']['per_complex_calls']['value']); print(d['roofline']['kernel_ms']); print('lv3 launch ms', d['roofline']['launch_ms'])"
fi
