#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== elementwise bench"; timeout 300 python scripts/elementwise_bench.py > gpurun_out/elementwise_bench.jsonl 2> gpurun_out/elementwise_bench.err; echo "rc=$?"; cat gpurun_out/elementwise_bench.jsonl | cut -c1-260; tail -3 gpurun_out/elementwise_bench.err
echo "== ncu elementwise"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:"rmsnorm_kernel|rotary_kernel|silu_and_mul_kernel|store_kv_prefill_kernel" -s 12 -c 16 -o gpurun_out/elementwise_r1 -f python scripts/elementwise_bench.py > gpurun_out/ncu_elementwise.log 2>&1; echo "ncu rc=$?"
