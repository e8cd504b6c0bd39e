#!/bin/bash
# Multi-GPU session (charged N x box time: everything that can share a process group does; most important results first).
#   gpurun --gpus 2 --timeout 900  -- 'bash scripts/gpu_tp_n.sh'        # pre-flight: correctness of every exchange variant
#   gpurun --gpus 8 --timeout 1200 -- 'bash scripts/gpu_tp_n.sh'        # the TP 8 numbers (configs[1], [3], [4])
set -u
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l); echo "gpus: $N"
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
if [ "$N" -le 2 ]; then
  echo "== NVLink probe (flag round trip, peer read / write bandwidth)"
  timeout 120 probes/build/p2p_latency > gpurun_out/p2p_latency_n$N.log 2>&1; echo "rc=$?"; tail -12 gpurun_out/p2p_latency_n$N.log
fi
echo "== TP parity tests (all exchange variants, sharded lm_head, CUDA graph)"
if [ "$N" -ge 8 ]; then K="world8"; else K="not world8"; fi
timeout 600 python -m pytest tests/test_tp_gpu.py -m gpu -q -s -p no:cacheprovider --maxfail=6 -k "$K" > gpurun_out/pytest_tp_n$N.log 2>&1; echo "rc=$?"
grep -E "TP parity|passed|failed|Error|error|watchdog" gpurun_out/pytest_tp_n$N.log | tail -15 | cut -c1-600
echo "== sweep, configs[1] (batch 256, seq 4096), TP=$N"
if [ "$N" -ge 8 ]; then V="nccl,two_shot,nvls,ll,ll_nvls,nccl+lmhead+rs,ll+lmhead+rs,ll_nvls+lmhead+rs"; P="nccl,ll+lmhead+rs,ll_nvls+lmhead+rs,nvls";
else V="nccl,one_shot,two_shot,nvls,ll,ll_nvls,one_shot+lmhead+rs,ll+lmhead+rs,ll_nvls+lmhead+rs"; P="one_shot,ll,ll_nvls"; fi
timeout 600 $TR --master-port 29581 scripts/tp_sweep.py --gpus $N --variants $V --profile $P > gpurun_out/tp_sweep_n${N}_b256.jsonl 2> gpurun_out/tp_sweep_n${N}_b256.err; echo "rc=$?"
cut -c1-330 gpurun_out/tp_sweep_n${N}_b256.jsonl; grep -v "^\*\|^$\|Warning\|warn" gpurun_out/tp_sweep_n${N}_b256.err | tail -3 | cut -c1-300
if [ "$N" -ge 4 ]; then
  echo "== bench.py at N=$N with the candidate defaults (exchange = LL push, sharded lm_head, fused rotary+store)"
  for X in "--ll-nvls-allreduce" "--ll-allreduce"; do
    tag=$(echo "$X" | tr -d ' -')
    timeout 400 $TR --master-port 29585 bench.py --gpus $N --steps 20 --warmup 3 $X --shard-lm-head --fuse-rotary-store > gpurun_out/bench_n${N}_${tag}.json 2> gpurun_out/bench_n${N}_${tag}.err; echo "bench [$X] rc=$?"
    python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_n${N}_${tag}.json').read().strip().splitlines()[-1]); print('  value',round(d['value']),'e2e',round(d['e2e']['value']),'ms',round(d['ms_per_step'],3),d['tp_exchange'][:60],'|',d['lm_head'][:20],'| attn frac',round(d['roofline']['frac'],3))
except Exception as e: print('  no line', e)
PY
  done
  echo "== sweep, configs[3] (batch 1024, seq 4096), TP=$N"
  timeout 600 $TR --master-port 29582 scripts/tp_sweep.py --gpus $N --batch 1024 --variants nccl,ll+lmhead+rs,ll_nvls+lmhead+rs,nvls+lmhead+rs --profile ll_nvls+lmhead+rs > gpurun_out/tp_sweep_n${N}_b1024.jsonl 2> gpurun_out/tp_sweep_n${N}_b1024.err; echo "rc=$?"
  cut -c1-330 gpurun_out/tp_sweep_n${N}_b1024.jsonl; grep -v "^\*\|^$\|Warning\|warn" gpurun_out/tp_sweep_n${N}_b1024.err | tail -3 | cut -c1-300
fi
if [ "$N" -eq 8 ]; then
  echo "== configs[4]: Llama-3-70B TP=8, mixed prefill/decode at seq_len 8192 (512-token chunk + 64 decodes per step)"
  timeout 500 $TR --master-port 29583 scripts/sarathi_bench.py --gpus 8 --model llama3-70b --prompt 8192 --seqlen 8192 --prompts 2 --exchange ll_nvls --shard-lm-head > gpurun_out/sarathi_70b_tp8.json 2> gpurun_out/sarathi_70b_tp8.err; echo "rc=$?"
  cut -c1-700 gpurun_out/sarathi_70b_tp8.json; grep -v "^\*\|^$\|Warning\|warn" gpurun_out/sarathi_70b_tp8.err | tail -3 | cut -c1-300
  echo "== configs[4] decode leg: Llama-3-70B TP=8 pure decode, batch 256 at seq_len 8192"
  timeout 500 $TR --master-port 29584 scripts/tp_sweep.py --gpus 8 --model llama3-70b --batch 256 --seqlen 8192 --variants nccl,ll_nvls+lmhead+rs --profile ll_nvls+lmhead+rs > gpurun_out/tp_sweep_70b_n8.jsonl 2> gpurun_out/tp_sweep_70b_n8.err; echo "rc=$?"
  cut -c1-330 gpurun_out/tp_sweep_70b_n8.jsonl; grep -v "^\*\|^$\|Warning\|warn" gpurun_out/tp_sweep_70b_n8.err | tail -3 | cut -c1-300
fi
if [ "$N" -le 2 ]; then
  echo "== bench.py smoke at N=$N on the tiny model (TP parity path of the bench line: oracle vs gathered shards)"
  timeout 300 $TR --master-port 29586 bench.py --gpus $N --model tiny --batch 16 --seqlen 512 --steps 3 --warmup 3 --no-prefill --no-live-traffic > gpurun_out/bench_n${N}_tiny.json 2> gpurun_out/bench_n${N}_tiny.err; echo "rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_n${N}_tiny.json').read().strip().splitlines()[-1]); print('  tiny: value',round(d['value']),'parity',json.dumps(d.get('parity_at_bench_shape'))[:600])
except Exception as e: print('  no line', e)
PY
  grep -v "^\*\|^$\|Warning\|warn" gpurun_out/bench_n${N}_tiny.err | tail -4 | cut -c1-300
  echo "== bench.py default at N=$N (what the driver's scaling run will launch)"
  timeout 600 $TR --master-port 29585 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_n${N}_default.json 2> gpurun_out/bench_n${N}_default.err; echo "rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_n${N}_default.json').read().strip().splitlines()[-1]); print('  value',round(d['value']),'e2e',round(d['e2e']['value']),'ms',round(d['ms_per_step'],3),d['tp_exchange'],'|',d['lm_head'],'| attn frac',round(d['roofline']['frac'],3))
except Exception as e: print('  no line', e)
PY
  grep "^\[bench" gpurun_out/bench_n${N}_default.err | cut -c1-200 | tail -12
fi
