#!/bin/bash
# The 8-GPU session (charged 8x: most important first; N=4 jobs run two at a time on disjoint GPU halves).
#   gpurun --gpus 8 --timeout 1000 -- 'bash scripts/gpu_tp8.sh'
set -u
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l); echo "gpus: $N"; [ "$N" -ge 8 ] || { echo "needs 8 GPUs"; exit 1; }
TR8="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
TR4="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
clean() { grep -v "^\*\|^$\|Warning\|warn\|OMP_NUM" "$1" | tail -3 | cut -c1-300; }
echo "== world-8 TP parity: every exchange / sharding mode vs TP=1 (eager + CUDA graph), one process group"
timeout 400 python -m pytest tests/test_tp_gpu.py -m gpu -q -s -p no:cacheprovider -k world8 > gpurun_out/pytest_tp_n8.log 2>&1; echo "rc=$?"
grep -E "TP parity|passed|failed|Error|error|watchdog" gpurun_out/pytest_tp_n8.log | tail -8 | cut -c1-900
echo "== sweep, configs[1] (batch 256, seq 4096), TP=8"
timeout 400 $TR8 --master-port 29581 scripts/tp_sweep.py --gpus 8 --variants nccl,two_shot,nvls,ll,ll_nvls,nvls+lmhead+rs,ll+lmhead+rs,ll_nvls+lmhead+rs \
  --profile nccl,nvls+lmhead+rs,ll+lmhead+rs,ll_nvls+lmhead+rs > gpurun_out/tp_sweep_n8_b256.jsonl 2> gpurun_out/tp_sweep_n8_b256.err; echo "rc=$?"
cut -c1-330 gpurun_out/tp_sweep_n8_b256.jsonl; clean gpurun_out/tp_sweep_n8_b256.err
echo "== sweep, configs[3] (batch 1024, seq 4096), TP=8"
timeout 400 $TR8 --master-port 29582 scripts/tp_sweep.py --gpus 8 --batch 1024 --variants nccl,nvls+lmhead+rs,ll+lmhead+rs,ll_nvls+lmhead+rs --profile ll_nvls+lmhead+rs \
  > gpurun_out/tp_sweep_n8_b1024.jsonl 2> gpurun_out/tp_sweep_n8_b1024.err; echo "rc=$?"
cut -c1-330 gpurun_out/tp_sweep_n8_b1024.jsonl; clean gpurun_out/tp_sweep_n8_b1024.err
echo "== configs[4]: Llama-3-70B TP=8, mixed prefill/decode at seq_len 8192 (512-token chunk + 64 decodes per step)"
timeout 400 $TR8 --master-port 29583 scripts/sarathi_bench.py --gpus 8 --model llama3-70b --prompt 8192 --seqlen 8192 --prompts 2 --exchange ll_nvls --shard-lm-head \
  > gpurun_out/sarathi_70b_tp8.json 2> gpurun_out/sarathi_70b_tp8.err; echo "rc=$?"
cut -c1-700 gpurun_out/sarathi_70b_tp8.json; clean gpurun_out/sarathi_70b_tp8.err
echo "== configs[4] decode leg: Llama-3-70B TP=8 pure decode, batch 256 at seq_len 8192"
timeout 400 $TR8 --master-port 29584 scripts/tp_sweep.py --gpus 8 --model llama3-70b --batch 256 --seqlen 8192 --variants nccl,ll_nvls+lmhead+rs --profile ll_nvls+lmhead+rs \
  > gpurun_out/tp_sweep_70b_n8.jsonl 2> gpurun_out/tp_sweep_70b_n8.err; echo "rc=$?"
cut -c1-330 gpurun_out/tp_sweep_70b_n8.jsonl; clean gpurun_out/tp_sweep_70b_n8.err
echo "== TP=4, two jobs at once on GPUs 0-3 / 4-7: configs[1] sweep and configs[3] (batch 1024) sweep"
CUDA_VISIBLE_DEVICES=0,1,2,3 timeout 400 $TR4 --master-port 29591 scripts/tp_sweep.py --gpus 4 --variants nccl,one_shot,ll,ll_nvls,one_shot+lmhead+rs,ll+lmhead+rs,ll_nvls+lmhead+rs \
  --profile one_shot+lmhead+rs,ll+lmhead+rs > gpurun_out/tp_sweep_n4_b256.jsonl 2> gpurun_out/tp_sweep_n4_b256.err &
CUDA_VISIBLE_DEVICES=4,5,6,7 timeout 400 $TR4 --master-port 29592 scripts/tp_sweep.py --gpus 4 --batch 1024 --variants nccl,one_shot+lmhead+rs,ll+lmhead+rs,ll_nvls+lmhead+rs \
  > gpurun_out/tp_sweep_n4_b1024.jsonl 2> gpurun_out/tp_sweep_n4_b1024.err &
wait
cut -c1-330 gpurun_out/tp_sweep_n4_b256.jsonl; clean gpurun_out/tp_sweep_n4_b256.err
cut -c1-330 gpurun_out/tp_sweep_n4_b1024.jsonl; clean gpurun_out/tp_sweep_n4_b1024.err
du -sm gpurun_out
