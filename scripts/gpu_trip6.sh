#!/bin/bash
set -u
mkdir -p gpurun_out
nvidia-smi -L | head -3
echo "== pytest tp"; timeout 900 python -m pytest tests/test_tp_gpu.py -m gpu -q -p no:cacheprovider > gpurun_out/pytest_tp.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/pytest_tp.log
echo "== bench n=2"; timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench rc=$?"; cat gpurun_out/bench_n2.json; tail -5 gpurun_out/bench_n2.err
