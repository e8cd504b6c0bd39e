#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest tp world=4"; timeout 300 python -m pytest tests/test_tp_gpu.py -m gpu -q -p no:cacheprovider -x -k "4" > gpurun_out/pytest_tp4.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_tp4.log | cut -c1-300
echo "== bench n=4 fused"; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 4 --steps 10 --warmup 3 --fused-allreduce > gpurun_out/bench_n4_fused.json 2> gpurun_out/bench_n4_fused.err; echo "bench rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/bench_n4_fused.json').read().strip().splitlines()[-1]); print('value',d['value'],'e2e', d['e2e']['value'], 'ms', d['ms_per_step'],'frac', d['roofline']['frac'], d['prefill'])"; tail -3 gpurun_out/bench_n4_fused.err | cut -c1-300
