#!/bin/bash
# 1-GPU evidence session of round 2: the full default bench line (reference Triton arm fp16 + bf16, oracle parity at the
# benchmarked shape, live ncu DRAM traffic), the TP-shard launch lists, prefill kernel vs the reference's flash_attn call,
# compute-sanitizer over the tcgen05 / mbarrier kernels.
#   gpurun --timeout 1500 -- 'bash scripts/gpu_r2_1gpu_evidence.sh'
set -u
mkdir -p gpurun_out
echo "== bench.py default (N=1)"
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_n1_full.json 2> gpurun_out/bench_n1_full.err; echo "rc=$?"
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench_n1_full.json').read().strip().splitlines()[-1])
    print('  value',round(d['value']),'e2e',round(d['e2e']['value']),'ms',round(d['ms_per_step'],3),'frac',round(d['roofline']['frac'],3),'traffic',d['roofline']['traffic'])
    print('  traffic_source', str(d['roofline']['traffic_source'])[:200])
    rt=d.get('reference_triton') or {}
    for k,v in (rt.get('runs') or {}).items(): print('  ref_triton',k,{x:v.get(x) for x in ('value','ms_per_step','paged_attention_ms_per_layer','error','wall_s')})
    print('  ratios',{k:v for k,v in rt.items() if k.startswith('e2e_over')})
    p=d.get('parity_at_bench_shape') or {}
    print('  parity ok',p.get('ok'),'attn worst',p.get('attention_worst_rel_err'),'seqs',p.get('sequences'))
    print('  cpu', (d.get('cpu_baseline') or {}).get('value'), 'prefill', d.get('prefill'))
except Exception as e: print('  no line', e)
PY
tail -5 gpurun_out/bench_n1_full.err | cut -c1-300
echo "== TP shard launch lists on one GPU (exchange absent)"
for tp in 8 4; do
  timeout 300 python scripts/shard_profile.py --tp $tp > gpurun_out/shard_tp$tp.json 2> gpurun_out/shard_tp$tp.err; echo "tp$tp rc=$?"; cat gpurun_out/shard_tp$tp.json | cut -c1-400
  head -14 gpurun_out/shard_launches_tp${tp}_b256_s4096.csv | cut -c1-220
done
timeout 300 python scripts/shard_profile.py --tp 8 --fuse-rotary-store --tag _frs > gpurun_out/shard_tp8_frs.json 2> gpurun_out/shard_tp8_frs.err; echo "tp8 frs rc=$?"; cat gpurun_out/shard_tp8_frs.json | cut -c1-300
echo "== prefill attention kernels vs flash_attn_varlen_func (the reference's call)"
timeout 300 python scripts/prefill_bench.py > gpurun_out/prefill_kernel_bench.jsonl 2> gpurun_out/prefill_kernel_bench.err; echo "rc=$?"; cut -c1-300 gpurun_out/prefill_kernel_bench.jsonl; tail -2 gpurun_out/prefill_kernel_bench.err | cut -c1-200
echo "== compute-sanitizer memcheck (paged decode gen 2, prefill gen 2 packed + paged, fused rotary/store)"
export SLLM_PAGED_ATTN_GEN=0 SLLM_PREFILL_ATTN_GEN=0
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 --launch-timeout 120 python -m pytest -q -p no:cacheprovider -x \
  "tests/test_kernels_gpu.py::test_paged_attention_golden" "tests/test_kernels_gpu.py::test_prefill_attention_golden" \
  "tests/test_chunked_prefill_gpu.py" -k "golden or store or sarathi or prefix0 or 1-token or ragged" > gpurun_out/sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"
grep -E "ERROR SUMMARY|passed|failed|Invalid|error" gpurun_out/sanitizer_memcheck.log | tail -6 | cut -c1-300
echo "== compute-sanitizer racecheck (shared-memory hazards; mbarrier/TMA traffic is invisible to it, see the log header)"
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 --launch-timeout 120 python -m pytest -q -p no:cacheprovider -x \
  "tests/test_kernels_gpu.py::test_paged_attention_golden" "tests/test_kernels_gpu.py::test_prefill_attention_golden" > gpurun_out/sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?"
grep -E "RACECHECK SUMMARY|passed|failed|hazard" gpurun_out/sanitizer_racecheck.log | tail -6 | cut -c1-300
