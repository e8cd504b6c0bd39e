#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest all gpu"; timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --maxfail=10 > gpurun_out/pytest_all.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_all.log | cut -c1-300
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
echo "== bench"; timeout 900 python bench.py > gpurun_out/bench_final_n1.json 2> gpurun_out/bench_final_n1.err; echo "bench rc=$?"; cat gpurun_out/bench_final_n1.json; tail -3 gpurun_out/bench_final_n1.err
echo "== ncu prefill"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:prefill_attn_tc_kernel -s 3 -c 1 -o gpurun_out/prefill_tc_r1 -f python scripts/prefill_bench.py > gpurun_out/ncu_prefill.log 2>&1; echo "ncu rc=$?"
