#!/bin/bash
# Multi-GPU session (charged N x box time: everything that can share a process group does).
#   gpurun --gpus 2 --timeout 900  -- 'bash scripts/gpu_tp_n.sh'        # pre-flight: correctness of every exchange variant
#   gpurun --gpus 8 --timeout 1500 -- 'bash scripts/gpu_tp_n.sh'        # the TP 8 numbers (configs[1], [3], [4])
set -u
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l); echo "gpus: $N"
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
echo "== NVLink probe (flag round trip, peer read / write bandwidth)"
timeout 120 probes/build/p2p_latency > gpurun_out/p2p_latency_n$N.log 2>&1; echo "rc=$?"; tail -12 gpurun_out/p2p_latency_n$N.log
echo "== TP parity tests (every world size <= $N; all exchange variants, sharded lm_head, CUDA graph)"
timeout 900 python -m pytest tests/test_tp_gpu.py -m gpu -q -s -p no:cacheprovider --maxfail=6 > gpurun_out/pytest_tp_n$N.log 2>&1; echo "rc=$?"
grep -E "TP parity|passed|failed|Error|error" gpurun_out/pytest_tp_n$N.log | tail -15 | cut -c1-400
echo "== sweep, configs[1] (batch 256, seq 4096), TP=$N"
timeout 600 $TR --master-port 29581 scripts/tp_sweep.py --gpus $N --profile nccl,nvls+lmhead+rs,two_shot+lmhead+rs > gpurun_out/tp_sweep_n${N}_b256.jsonl 2> gpurun_out/tp_sweep_n${N}_b256.err; echo "rc=$?"
cat gpurun_out/tp_sweep_n${N}_b256.jsonl | cut -c1-400; tail -3 gpurun_out/tp_sweep_n${N}_b256.err | cut -c1-300
if [ "$N" -ge 4 ]; then
  echo "== sweep, configs[3] (batch 1024, seq 4096), TP=$N"
  timeout 600 $TR --master-port 29582 scripts/tp_sweep.py --gpus $N --batch 1024 --variants nccl,two_shot+lmhead+rs,nvls+lmhead+rs --profile nvls+lmhead+rs > gpurun_out/tp_sweep_n${N}_b1024.jsonl 2> gpurun_out/tp_sweep_n${N}_b1024.err; echo "rc=$?"
  cat gpurun_out/tp_sweep_n${N}_b1024.jsonl | cut -c1-400; tail -3 gpurun_out/tp_sweep_n${N}_b1024.err | cut -c1-300
fi
if [ "$N" -eq 8 ]; then
  echo "== configs[4]: Llama-3-70B TP=8, mixed prefill/decode at seq_len 8192 (512-token chunk + 64 decodes per step)"
  timeout 600 $TR --master-port 29583 scripts/sarathi_bench.py --gpus 8 --model llama3-70b --prompt 8192 --seqlen 8192 --prompts 2 > gpurun_out/sarathi_70b_tp8.json 2> gpurun_out/sarathi_70b_tp8.err; echo "rc=$?"
  cat gpurun_out/sarathi_70b_tp8.json | cut -c1-600; tail -3 gpurun_out/sarathi_70b_tp8.err | cut -c1-300
  echo "== configs[4] decode leg: Llama-3-70B TP=8 pure decode, batch 256 at seq_len 8192"
  timeout 600 $TR --master-port 29584 scripts/tp_sweep.py --gpus 8 --model llama3-70b --batch 256 --seqlen 8192 --variants nccl,nvls+lmhead+rs > gpurun_out/tp_sweep_70b_n8.jsonl 2> gpurun_out/tp_sweep_70b_n8.err; echo "rc=$?"
  cat gpurun_out/tp_sweep_70b_n8.jsonl | cut -c1-400; tail -3 gpurun_out/tp_sweep_70b_n8.err | cut -c1-300
fi
echo "== bench.py default at N=$N (what the driver's scaling run will launch)"
timeout 600 $TR --master-port 29585 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_n${N}_default.json 2> gpurun_out/bench_n${N}_default.err; echo "rc=$?"
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_n${N}_default.json').read().strip().splitlines()[-1]); print('  value',round(d['value']),'e2e',round(d['e2e']['value']),'ms',round(d['ms_per_step'],3),d['tp_exchange'],'|',d['lm_head'],'| attn frac',round(d['roofline']['frac'],3))
except Exception as e: print('  no line', e)
PY
tail -3 gpurun_out/bench_n${N}_default.err | cut -c1-300
