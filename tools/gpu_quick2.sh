#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 240 --timeout-method thread -k "conv_kernel_paths or forward_stages" > gpurun_out/pytest_quick.log 2>&1; rc=$?
echo "quick rc=$rc" >> gpurun_out/pytest_quick.log; tail -5 gpurun_out/pytest_quick.log
if [ $rc -eq 0 ]; then
  bash tools/tcr_trace.sh $1
  timeout 900 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-sparse > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
  python -c "
import json; d=json.load(open('gpurun_out/bench.json')); print('value', d['value'], 'e2e', d['e2e']['value'], d['e2e']['per_complex_calls']['value']); print(d['roofline']['kernel_ms']); print('lv3 launch ms', d['roofline']['launch_ms'])"
  tail -3 gpurun_out/bench.err
fi
