#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest paged gen2 + model"; timeout 1200 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py -m gpu -q -p no:cacheprovider -k "paged or model or graph or profile or prefill_attention_long" --maxfail=10 > gpurun_out/pytest_gen2.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/pytest_gen2.log
echo "== bench gen2"; timeout 1200 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_gen2.json 2> gpurun_out/bench_gen2.err; echo "bench rc=$?"; cat gpurun_out/bench_gen2.json; tail -5 gpurun_out/bench_gen2.err
echo "== bench gen1"; SLLM_PAGED_ATTN_GEN=1 timeout 1200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_gen1.json 2> gpurun_out/bench_gen1.err; echo "bench rc=$?"; cat gpurun_out/bench_gen1.json; tail -5 gpurun_out/bench_gen1.err
echo "== ncu full gen2"; timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:paged_attn_tc_kernel -c 2 -o gpurun_out/paged_attn_tc_r1 -f python bench.py --profile-range 1 --no-cpu-baseline --no-prefill > gpurun_out/ncu_full_tc.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out
