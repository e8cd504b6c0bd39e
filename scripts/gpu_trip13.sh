#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest tp (nccl + fused)"; timeout 400 python -m pytest tests/test_tp_gpu.py -m gpu -q -p no:cacheprovider -x > gpurun_out/pytest_tp.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_tp.log | cut -c1-300
echo "== bench n=2 fused"; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 10 --warmup 3 --fused-allreduce --no-prefill > gpurun_out/bench_n2_fused.json 2> gpurun_out/bench_n2_fused.err; echo "bench rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/bench_n2_fused.json').read().strip().splitlines()[-1]); print('value',d['value'],'e2e', d['e2e']['value'], 'ms', d['ms_per_step'],'frac', d['roofline']['frac'])"; tail -4 gpurun_out/bench_n2_fused.err | cut -c1-300
echo "== bench n=2 nccl"; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --steps 10 --warmup 3 --no-prefill > gpurun_out/bench_n2_nccl.json 2> gpurun_out/bench_n2_nccl.err; echo "bench rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/bench_n2_nccl.json').read().strip().splitlines()[-1]); print('value',d['value'],'e2e', d['e2e']['value'], 'ms', d['ms_per_step'],'frac', d['roofline']['frac'])"
