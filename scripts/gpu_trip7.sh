#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest prefill"; timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -p no:cacheprovider -k "prefill" --maxfail=6 > gpurun_out/pytest_prefill.log 2>&1; echo "pytest rc=$?"; tail -30 gpurun_out/pytest_prefill.log | cut -c1-300
echo "== prefill bench"; timeout 300 python scripts/prefill_bench.py > gpurun_out/prefill_bench.jsonl 2> gpurun_out/prefill_bench.err; echo "rc=$?"; cat gpurun_out/prefill_bench.jsonl; tail -5 gpurun_out/prefill_bench.err
