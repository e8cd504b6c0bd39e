#!/bin/bash
# 1-GPU session E: (background) build the Triton JIT cache of the unmodified reference for fp16 + the bf16-patched copy at batch 64
# (~8 min of ptxas, see scripts/ref_triton_bench.py) and bring it back COMPRESSED (gpurun merges at most 64 MiB);
# (foreground, meanwhile) the GPU test suite, an ncu source-level capture of the tcgen05 prefill kernel; (afterwards) the full
# default bench line with the warmed cache and the ncu time / DRAM bytes of the reference's _fwd_paged_attention_phase1.
#   gpurun --timeout 1500 -- 'bash scripts/gpu_r2_session_e.sh'
#   afterwards here:  cp gpurun_out/triton_cache_*.tar.gz baseline/_ref/
set -u
mkdir -p gpurun_out/triton_cache
for dt in fp16 bf16; do
  REF_DTYPE=$dt REF_BATCH=64 REF_STEPS=2 REF_WARMUP=1 REF_TRITON_CACHE_DIR=$PWD/gpurun_out/triton_cache/$dt \
    timeout 1200 python scripts/ref_triton_bench.py > gpurun_out/ref_warm_$dt.json 2> gpurun_out/ref_warm_$dt.err &
done
echo "== (while the reference compiles) GPU test suite (the KV-profile test is run later: the reference processes hold 72 GB now)"
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider --maxfail=5 --deselect tests/test_model_gpu.py::test_profile_num_blocks_and_exhaustion > gpurun_out/pytest_all_r2e.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/pytest_all_r2e.log | cut -c1-300
grep -h "golden_trace_vs_exact\|golden_trace_summary" gpurun_out/parity_log.jsonl | tail -8 | cut -c1-250
echo "== ncu source-level capture of prefill_attn_tc_kernel (8 x 4096)"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:prefill_attn_tc_kernel --launch-skip 2 --launch-count 1 -f -o gpurun_out/r2_prefill_tc python scripts/prefill_one.py > gpurun_out/ncu_prefill.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/ncu_prefill.log | cut -c1-200
ncu -i gpurun_out/r2_prefill_tc.ncu-rep --page raw --csv > gpurun_out/r2_prefill_tc_raw.csv 2>/dev/null
ncu -i gpurun_out/r2_prefill_tc.ncu-rep --page source --csv > gpurun_out/r2_prefill_tc_source.csv 2>/dev/null
ls -la gpurun_out/r2_prefill_tc* | cut -c1-200
echo "== waiting for the reference's Triton compiles"
wait
for dt in fp16 bf16; do echo "$dt: $(cut -c1-400 gpurun_out/ref_warm_$dt.json)"; tail -2 gpurun_out/ref_warm_$dt.err | cut -c1-300; done
for dt in fp16 bf16; do tar czf gpurun_out/triton_cache_$dt.tar.gz -C gpurun_out/triton_cache $dt; done
ls -la gpurun_out/triton_cache_*.tar.gz
mkdir -p baseline/_ref && cp gpurun_out/triton_cache_*.tar.gz baseline/_ref/ && rm -rf gpurun_out/triton_cache baseline/_ref/triton_cache
echo "== the KV-profile test (needs the whole GPU)"
timeout 300 python -m pytest tests/test_model_gpu.py::test_profile_num_blocks_and_exhaustion -m gpu -q -p no:cacheprovider 2>&1 | tail -2 | cut -c1-300
echo "== bench.py default (N=1) with the warmed cache"
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_n1_full.json 2> gpurun_out/bench_n1_full.err; echo "rc=$?"
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench_n1_full.json').read().strip().splitlines()[-1])
    print('  value',round(d['value']),'e2e',round(d['e2e']['value']),'ms',round(d['ms_per_step'],3),'frac',round(d['roofline']['frac'],3),'traffic',d['roofline']['traffic'])
    print('  traffic_source', str(d['roofline']['traffic_source'])[:200])
    rt=d.get('reference_triton') or {}
    for k,v in (rt.get('runs') or {}).items(): print('  ref_triton',k,{x:v.get(x) for x in ('value','ms_per_step','paged_attention_ms_per_layer','first_forward_s','error','wall_s')})
    print('  ratios',{k:v for k,v in rt.items() if k.startswith('e2e_over')})
    p=d.get('parity_at_bench_shape') or {}
    print('  parity ok',p.get('ok'),'attn worst',p.get('attention_worst_rel_err'))
    for s in p.get('sequences') or []: print('   ', s)
    print('  cpu', (d.get('cpu_baseline') or {}).get('value'), 'prefill', d.get('prefill'))
except Exception as e: print('  no line', e)
PY
grep "^\[bench" gpurun_out/bench_n1_full.err | cut -c1-200; grep -v "^\[bench" gpurun_out/bench_n1_full.err | tail -4 | cut -c1-300
echo "== ncu: the reference's phase-1 kernel (time, DRAM bytes), fp16 as shipped, warm cache"
REF_STEPS=1 REF_WARMUP=0 timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv -k regex:_fwd_paged_attention_phase --launch-skip 70 --launch-count 6 python scripts/ref_triton_bench.py > gpurun_out/ncu_ref_phase1.csv 2> gpurun_out/ncu_ref_phase1.err; echo "rc=$?"
grep -E "paged_attention" gpurun_out/ncu_ref_phase1.csv | cut -d, -f5,12- | head -20 | cut -c1-200
gzip -f gpurun_out/r2_prefill_tc_source.csv gpurun_out/r2_prefill_tc_raw.csv 2>/dev/null
sz=$(du -sm gpurun_out | cut -f1); echo "gpurun_out: ${sz} MiB"
if [ "$sz" -gt 58 ]; then echo "dropping the .ncu-rep (csv exports kept)"; rm -f gpurun_out/r2_prefill_tc.ncu-rep; du -sm gpurun_out; fi
