#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest paged"; timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py -m gpu -q -p no:cacheprovider -k "paged or model or graph or profile" --maxfail=10 > gpurun_out/pytest_gen2b.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_gen2b.log
echo "== bench gen2"; timeout 1200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_gen2b.json 2> gpurun_out/bench_gen2b.err; echo "bench rc=$?"; cat gpurun_out/bench_gen2b.json; tail -5 gpurun_out/bench_gen2b.err
echo "== ncu full gen2"; timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:paged_attn_tc_kernel -c 1 -o gpurun_out/paged_attn_tc_r1b -f python bench.py --profile-range 1 --no-cpu-baseline --no-prefill > gpurun_out/ncu_full_tcb.log 2>&1; echo "ncu rc=$?"
