#!/bin/bash
# 1-GPU: final validation of the shipped build + the reference arm with the relocated Triton cache.
#   gpurun --timeout 600 -- 'bash scripts/gpu_r2_final.sh'
set -u
mkdir -p gpurun_out
echo "== GPU suite, shipped library"
timeout 300 python -m pytest tests -m gpu -q -p no:cacheprovider --maxfail=5 > gpurun_out/pytest_final.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/pytest_final.log | cut -c1-300
echo "== reference Triton arm with the packed cache (fp16, then bf16)"
for dt in fp16 bf16; do
  REF_DTYPE=$dt REF_STEPS=10 REF_WARMUP=2 timeout 110 python scripts/ref_triton_bench.py > gpurun_out/ref_triton_$dt.json 2> gpurun_out/ref_triton_$dt.err; echo "$dt rc=$?"
  python -c "
import json
try:
    d=json.loads(open('gpurun_out/ref_triton_$dt.json').read().strip().splitlines()[-1]); print('  ', {k:d.get(k) for k in ('value','ms_per_step','paged_attention_ms_per_layer','first_forward_s','dtype')})
except Exception as e: print('   no line', e)"
done
echo "== prefill kernel, shipped library"
timeout 200 python scripts/prefill_bench.py > gpurun_out/prefill_bench_shipped.jsonl 2>/dev/null; echo "rc=$?"
# (the run of this script in round 2 also executed the suite, the prefill bench and a one-CTA timeline with the "ping-pong" scheduling
#  variant of the prefill kernel - commit 5f0c0d1^..; it was 4 % slower and was removed: profiles/r2_prefill_pingpong_trace.txt,
#  profiles/r2_prefill_bench_pingpong_variant.jsonl)
echo "== bench.py (N=1, quick: no reference arm / CPU baseline / prefill; parity on 1 sequence)"
timeout 300 python bench.py --steps 20 --warmup 3 --no-ref-triton --no-cpu-baseline --no-prefill --parity-seqs 1 > gpurun_out/bench_n1_quick.json 2> gpurun_out/bench_n1_quick.err; echo "rc=$?"
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench_n1_quick.json').read().strip().splitlines()[-1])
    print('  value',round(d['value']),'e2e',round(d['e2e']['value']),'ms',round(d['ms_per_step'],3),'frac',round(d['roofline']['frac'],3),'launches/step',d['gpu_launches']//d['steps'],'parity',(d.get('parity_at_bench_shape') or {}).get('ok'))
except Exception as e: print('  no line', e)
PY
grep -v "^\[bench" gpurun_out/bench_n1_quick.err | tail -3 | cut -c1-300
