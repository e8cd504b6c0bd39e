#!/bin/bash
# Builds the Triton JIT cache of the UNMODIFIED reference's kernels (fp16 as shipped + the bf16-patched copy) on a B200 box, so
# that bench.py's reference Triton arm does not spend ~8 minutes per dtype in ptxas (the reference's phase-1 kernel unrolls 128
# pages and is specialised three times on cur_layer).  Both dtypes compile concurrently at batch 64 (same constexprs as batch
# 256: seq_block_size 2048, 264 blocks per sequence; 36 GB each), while the GPU runs the test suite and the TP-shard profile;
# then the full default bench line is produced with the warmed cache.
#   gpurun --timeout 1500 -- 'bash scripts/gpu_r2_warm_triton_cache.sh'
#   afterwards here:  mkdir -p baseline/_ref/triton_cache && cp -r gpurun_out/triton_cache/* baseline/_ref/triton_cache/
set -u
mkdir -p gpurun_out/triton_cache
for dt in fp16 bf16; do
  REF_DTYPE=$dt REF_BATCH=64 REF_STEPS=2 REF_WARMUP=1 REF_TRITON_CACHE_DIR=$PWD/gpurun_out/triton_cache/$dt \
    timeout 1200 python scripts/ref_triton_bench.py > gpurun_out/ref_warm_$dt.json 2> gpurun_out/ref_warm_$dt.err &
done
echo "== (while the reference compiles) GPU test suite"
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider --maxfail=5 > gpurun_out/pytest_all_r2b.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/pytest_all_r2b.log | cut -c1-300
echo "== TP=8 / TP=4 shard on one GPU with the in-kernel split merge"
for tp in 8 4; do
  timeout 300 python scripts/shard_profile.py --tp $tp --fuse-rotary-store --tag _merge > gpurun_out/shard_tp${tp}_merge.json 2> gpurun_out/shard_tp${tp}_merge.err; echo "tp$tp rc=$?"; cut -c1-300 gpurun_out/shard_tp${tp}_merge.json
  head -6 gpurun_out/shard_launches_tp${tp}_b256_s4096_merge.csv | cut -c1-200
done
SLLM_PAGED_ATTN_FUSED_MERGE=0 timeout 300 python scripts/shard_profile.py --tp 8 --fuse-rotary-store --tag _nomerge > gpurun_out/shard_tp8_nomerge.json 2> gpurun_out/shard_tp8_nomerge.err; echo "tp8 separate merge rc=$?"; cut -c1-300 gpurun_out/shard_tp8_nomerge.json
echo "== waiting for the reference's Triton compiles"
wait
for dt in fp16 bf16; do echo "$dt: $(cut -c1-500 gpurun_out/ref_warm_$dt.json)"; tail -2 gpurun_out/ref_warm_$dt.err | cut -c1-300; done
du -sh gpurun_out/triton_cache/*
mkdir -p baseline/_ref/triton_cache && cp -r gpurun_out/triton_cache/* baseline/_ref/triton_cache/
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
    print('  parity ok',p.get('ok'),'attn worst',p.get('attention_worst_rel_err'),'seqs',p.get('sequences'))
    print('  cpu', (d.get('cpu_baseline') or {}).get('value'), 'prefill', d.get('prefill'))
except Exception as e: print('  no line', e)
PY
grep "^\[bench" gpurun_out/bench_n1_full.err | cut -c1-200; tail -3 gpurun_out/bench_n1_full.err | cut -c1-300
