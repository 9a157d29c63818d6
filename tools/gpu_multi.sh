#!/bin/bash
# N-GPU runs on one box: weak-scaling bench line and the strong-scaled BASELINE configs 3/4/5 through the product entry point
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus_n$N.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29501 \
    bench.py --gpus $N --steps 2 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/bench_n$N.json')); print('N=$N value', d['value'], 'e2e', d['e2e']['value'], 'sparse', d.get('sparse',{}).get('value')); print(d['work']['per_rank'])"
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    tools/bench_configs.py --configs ${2:-3 4 5} > gpurun_out/configs_n$N.jsonl 2> gpurun_out/configs_n$N.err; echo "configs rc=$?"
cat gpurun_out/configs_n$N.jsonl; tail -3 gpurun_out/configs_n$N.err
