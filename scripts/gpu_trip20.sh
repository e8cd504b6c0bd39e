#!/bin/bash
set -u
mkdir -p gpurun_out
EW_ONCE=1 timeout 400 ncu --set full --clock-control none -k regex:"rmsnorm_kernel|rotary_kernel|silu_and_mul_kernel|store_kv_prefill_kernel" -c 5 -o gpurun_out/elementwise_r1 -f python scripts/elementwise_bench.py > gpurun_out/ncu_elementwise.log 2>&1; echo "ncu rc=$?"
