#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== probe"; timeout 120 ./probes/build/umma_probe > gpurun_out/umma_probe.log 2>&1; echo "probe rc=$?"; cat gpurun_out/umma_probe.log
echo "== pytest model"; timeout 900 python -m pytest tests/test_model_gpu.py -m gpu -q -p no:cacheprovider > gpurun_out/pytest_model.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_model.log
echo "== smoke"; timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
echo "== bench"; timeout 1200 python bench.py --steps 10 --warmup 3 > gpurun_out/bench1.json 2> gpurun_out/bench1.err; echo "bench rc=$?"; cat gpurun_out/bench1.json; tail -5 gpurun_out/bench1.err
echo "== ncu launches"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_r1.csv python bench.py --profile-range 1 --no-cpu-baseline --no-prefill > gpurun_out/ncu_launch.log 2>&1; echo "ncu1 rc=$?"
echo "== ncu full"; timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:paged_attn_kernel -c 2 -o gpurun_out/paged_attn_r1 -f python bench.py --profile-range 1 --no-cpu-baseline --no-prefill > gpurun_out/ncu_full.log 2>&1; echo "ncu2 rc=$?"
ls -la gpurun_out
