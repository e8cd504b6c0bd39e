#!/bin/bash
# First GPU run of everything marked `pending_gpu` (written when no GPU time was left), then the A/B numbers they are for.
#   1 GPU : gpurun --timeout 1500 -- 'bash scripts/gpu_validate_pending.sh'
#   N GPUs: gpurun --gpus N --timeout 1500 -- 'bash scripts/gpu_validate_pending.sh'
set -u
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l); echo "gpus: $N"
export SLLM_RUN_PENDING=1
# SLLM_DEBUG_SYNC=1 makes every native call synchronise and name itself when a kernel traps (use when a pending test fails)
echo "== pending 1-GPU tests (chunked prefill: store, paged prefill attention gen1+gen2, model; fused rotary+store)"
timeout 900 python -m pytest tests/test_chunked_prefill_gpu.py tests/test_decode_fusion_gpu.py tests/test_swap_device_gpu.py -m gpu -q -p no:cacheprovider --maxfail=8 > gpurun_out/pytest_pending_1gpu.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/pytest_pending_1gpu.log | cut -c1-300
echo "== validated suite still green"
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider --maxfail=5 > gpurun_out/pytest_all.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/pytest_all.log | cut -c1-300
echo "== SARATHI bench (configs[2])"
timeout 600 python scripts/sarathi_bench.py > gpurun_out/sarathi_bench.json 2> gpurun_out/sarathi_bench.err; echo "rc=$?"; cat gpurun_out/sarathi_bench.json; tail -3 gpurun_out/sarathi_bench.err | cut -c1-300
echo "== reference Triton path (baseline/_ref, scripts/install_reference.sh)"
timeout 900 python scripts/ref_triton_bench.py > gpurun_out/ref_triton_bench.json 2> gpurun_out/ref_triton_bench.err; echo "rc=$?"; cat gpurun_out/ref_triton_bench.json; tail -3 gpurun_out/ref_triton_bench.err | cut -c1-300
[ "$N" -eq 1 ] && echo "== N=1 bench A/B: fused rotary + store"
for variant in "" "--fuse-rotary-store"; do
  [ "$N" -eq 1 ] || break
  tag=$(echo "n1default$variant" | tr -d ' ' | tr -- '-' '_')
  timeout 600 python bench.py --steps 20 --warmup 3 --no-prefill --no-cpu-baseline $variant > gpurun_out/bench_${tag}.json 2> gpurun_out/bench_${tag}.err; echo "bench [$variant] rc=$?"
  python -c "
import json
try:
    d=json.loads(open('gpurun_out/bench_${tag}.json').read().strip().splitlines()[-1]); print('  value',round(d['value']),'e2e',round(d['e2e']['value']),'ms',round(d['ms_per_step'],3),'launches/step',d['gpu_launches']//d['steps'])
except Exception as e: print('  no line', e)"
done
if [ "$N" -ge 2 ]; then
  echo "== NVLink probe (flag round trip, peer read / write bandwidth)"
  mkdir -p probes/build; [ -x probes/build/p2p_latency ] || nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o probes/build/p2p_latency probes/p2p_latency.cu
  timeout 120 probes/build/p2p_latency > gpurun_out/p2p_latency.log 2>&1; echo "rc=$?"; cat gpurun_out/p2p_latency.log
  echo "== pending TP tests (two-shot exchange, vocab-sharded lm_head)"
  timeout 900 python -m pytest tests/test_tp_gpu.py -m gpu -q -p no:cacheprovider --maxfail=4 > gpurun_out/pytest_pending_tp.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/pytest_pending_tp.log | cut -c1-300
  for variant in "" "--shard-lm-head" "--two-shot-allreduce" "--nvls-allreduce" "--nvls-allreduce --shard-lm-head --fuse-rotary-store"; do
    tag=$(echo "default$variant" | tr -d ' ' | tr -- '-' '_')
    timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29581 bench.py --gpus $N --steps 10 --warmup 3 --no-prefill $variant > gpurun_out/bench_n${N}_${tag}.json 2> gpurun_out/bench_n${N}_${tag}.err; echo "bench [$variant] rc=$?"
    python -c "
import json,sys
try:
    d=json.loads(open('gpurun_out/bench_n${N}_${tag}.json').read().strip().splitlines()[-1]); print('  value',round(d['value']),'e2e',round(d['e2e']['value']),'ms',round(d['ms_per_step'],3),d['tp_exchange'],'|',d['lm_head'])
except Exception as e: print('  no line', e)"
  done
fi
