#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== tma_bw"; timeout 300 ./probes/build/tma_bw > gpurun_out/tma_bw.log 2>&1; echo "rc=$?"; cat gpurun_out/tma_bw.log
echo "== ref triton"; REF_STEPS=5 timeout 900 python scripts/ref_triton_bench.py > gpurun_out/ref_triton.json 2> gpurun_out/ref_triton.err; echo "rc=$?"; cat gpurun_out/ref_triton.json; tail -5 gpurun_out/ref_triton.err
