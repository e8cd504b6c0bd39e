#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest tp"; timeout 300 python -m pytest tests/test_tp_gpu.py -m gpu -q -p no:cacheprovider -x > gpurun_out/pytest_tp.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_tp.log | cut -c1-250
echo "== bench n=2"; timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench rc=$?"; cat gpurun_out/bench_n2.json; tail -3 gpurun_out/bench_n2.err
echo "== bench n=2 reference arm"; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/bench_ref_n2.json 2> gpurun_out/bench_ref_n2.err; echo "ref rc=$?"; cut -c1-400 gpurun_out/bench_ref_n2.json
