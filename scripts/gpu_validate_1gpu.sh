#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest all gpu"; timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider --maxfail=10 > gpurun_out/pytest_all.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/pytest_all.log | cut -c1-300
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
echo "== bench"; timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench_n1.json')); print('value',d['value'],'e2e', d['e2e']['value'], 'ms', d['ms_per_step'], d['e2e']['ms_per_step'],'frac', d['roofline']['frac'], d['clocks'], d['prefill'])"; tail -3 gpurun_out/bench_n1.err
echo "== ncu launches"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_r1b.csv python bench.py --profile-range 1 --no-cpu-baseline --no-prefill > gpurun_out/ncu_launch.log 2>&1; echo "ncu rc=$?"
