#!/bin/bash
set -u
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l); echo "gpus: $N"
timeout 360 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n${N}_default.json 2> gpurun_out/bench_n${N}_default.err; echo "bench rc=$?"; wc -l gpurun_out/bench_n${N}_default.json; python -c "
import json; d=json.loads(open('gpurun_out/bench_n${N}_default.json').read().strip().splitlines()[-1]); print('value',d['value'],'e2e', d['e2e']['value'], 'ms', d['ms_per_step'],'frac', d['roofline']['frac'], d['tp_exchange'], d['prefill'])"; tail -3 gpurun_out/bench_n${N}_default.err | cut -c1-300
