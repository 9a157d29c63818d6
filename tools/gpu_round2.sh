#!/bin/bash
# full GPU suite, bench in the default mode (k_conv_tcr) and in mode 1 (k_conv_fused + k_acc_tc), launch list with DRAM bytes
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 --timeout-method thread > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -12 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/bench.json')); print('value', d['value'], 'e2e', d['e2e']['value'], d['e2e']['per_complex_calls']['value'], 'sparse', d.get('sparse',{}).get('value')); print(d['roofline']['kernel_ms']); print('lv3 launch ms', d['roofline']['launch_ms'], 'frac', d['roofline']['frac'], d['roofline']['bound'], 'cpu', d.get('cpu_baseline'))"
DDK_TC=1 timeout 900 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-sparse > gpurun_out/bench_tc1.json 2>> gpurun_out/bench.err
python -c "
import json; d=json.load(open('gpurun_out/bench_tc1.json')); print('TC1 value', d['value'], 'e2e', d['e2e']['value']); print(d['roofline']['kernel_ms'])"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err
python -c "
import json; d=json.load(open('gpurun_out/bench_ref.json')); print('REF', d['value'], d['work'])"
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 2400 --csv \
    --log-file gpurun_out/launches_bench.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-sparse > gpurun_out/launches_bench.log 2>&1
python tools/launch_summary.py gpurun_out/launches_bench.csv --traffic-json gpurun_out/conv_lv3_traffic.json | tee gpurun_out/launches_bench_summary.txt | head -14
