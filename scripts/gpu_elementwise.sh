#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest"; timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py -m gpu -q -p no:cacheprovider -k "rmsnorm or silu or model or graph" --maxfail=6 > gpurun_out/pytest_ew.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_ew.log | cut -c1-300
echo "== elementwise bench"; timeout 300 python scripts/elementwise_bench.py > gpurun_out/elementwise_bench.jsonl 2> gpurun_out/elementwise_bench.err; echo "rc=$?"; python -c "
import json
for l in open('gpurun_out/elementwise_bench.jsonl'):
    d=json.loads(l); print(d['kernel'][:52], round(d['ms'],4), round(d['achieved_GBps']), round(d['frac_of_measured_peak'],3))"; tail -3 gpurun_out/elementwise_bench.err
echo "== ncu elementwise"; timeout 600 ncu --set full --clock-control none -k regex:"rmsnorm_kernel|rotary_kernel|silu_and_mul_kernel|store_kv_prefill_kernel" -s 60 -c 5 -o gpurun_out/elementwise_r1 -f python scripts/elementwise_bench.py > gpurun_out/ncu_elementwise.log 2>&1; echo "ncu rc=$?"
