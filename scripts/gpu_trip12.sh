#!/bin/bash
set -u
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l); echo "gpus: $N"
echo "== bench n=$N"; timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench rc=$?"; python -c "
import json,sys; d=json.load(open('gpurun_out/bench_n$N.json')); print('value',d['value'],'e2e', d['e2e']['value'], 'ms', d['ms_per_step'],'frac', d['roofline']['frac'], 'pa_ms', d['roofline']['mean_launch_ms'], d['prefill'])"; tail -3 gpurun_out/bench_n$N.err | cut -c1-300
