#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest prefill+model"; timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py -m gpu -q -p no:cacheprovider -k "prefill or model or fused_qkv" --maxfail=6 > gpurun_out/pytest_prefill.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_prefill.log | cut -c1-300
echo "== prefill bench"; timeout 300 python scripts/prefill_bench.py > gpurun_out/prefill_bench.jsonl 2> gpurun_out/prefill_bench.err; echo "rc=$?"; cat gpurun_out/prefill_bench.jsonl | grep gen2; tail -3 gpurun_out/prefill_bench.err
echo "== ncu prefill"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:prefill_attn_tc_kernel -s 3 -c 1 -o gpurun_out/prefill_tc_r1c -f python scripts/prefill_bench.py > gpurun_out/ncu_prefill.log 2>&1; echo "ncu rc=$?"
