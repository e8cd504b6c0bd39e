#!/bin/bash
# 1-GPU: final validation of the shipped build + the prefill scheduling experiment (variant libraries) + the reference arm with the
# relocated Triton cache.
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
echo "== GPU suite + prefill kernel + one-CTA timeline, ping-pong variant"
SLLM_LIB_PATH=$PWD/swiftllm_b200/libswiftllm_b200_pp.so timeout 300 python -m pytest tests -m gpu -q -p no:cacheprovider --maxfail=5 --deselect tests/test_cabi.py > gpurun_out/pytest_pp.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/pytest_pp.log | cut -c1-300
SLLM_LIB_PATH=$PWD/swiftllm_b200/libswiftllm_b200_pp.so timeout 200 python scripts/prefill_bench.py > gpurun_out/prefill_bench_pp.jsonl 2>/dev/null; echo "rc=$?"
python - <<'PY'
import json
for tag in ("shipped","pp"):
    try:
        for l in open(f'gpurun_out/prefill_bench_{tag}.jsonl'):
            d=json.loads(l)
            if 'tflops' in d and not d['gen'].startswith('gen1'): print('  ', tag.ljust(8), d['gen'][:30].ljust(30), d.get('Bp'), d.get('L'), round(d['ms'],3), 'ms', round(d['tflops']), 'TF/s', 'x FA2', round(d.get('speedup_over_reference_flash_attn',0),2))
    except Exception as e: print('   ', tag, 'no data', e)
PY
timeout 120 python scripts/prefill_trace.py > gpurun_out/prefill_trace_pp.txt 2>&1; echo "trace rc=$?"; head -30 gpurun_out/prefill_trace_pp.txt | cut -c1-160
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
