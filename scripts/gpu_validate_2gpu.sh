#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest tp"; timeout 300 python -m pytest tests/test_tp_gpu.py -m gpu -q -p no:cacheprovider -x > gpurun_out/pytest_tp.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_tp.log | cut -c1-200
echo "== bench n=2 default"; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 2 > gpurun_out/bench_n2_final.json 2> gpurun_out/bench_n2_final.err; echo "bench rc=$?"; wc -l gpurun_out/bench_n2_final.json; python -c "
import json; d=json.loads(open('gpurun_out/bench_n2_final.json').read().strip().splitlines()[-1]); print('value',d['value'],'e2e', d['e2e']['value'], 'ms', d['ms_per_step'],'frac', d['roofline']['frac'], d['tp_exchange'], d['prefill'])"
echo "== bench n=2 reference arm"; timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29572 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/bench_ref_n2.json 2> gpurun_out/bench_ref_n2.err; echo "ref rc=$?"; wc -l gpurun_out/bench_ref_n2.json; cut -c1-200 gpurun_out/bench_ref_n2.json
