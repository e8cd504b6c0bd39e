#!/usr/bin/env python
"""BASELINE.json configs[2]: Llama-3-8B bf16, SARATHI-style piggybacking - every step carries one 512-token CHUNK of a
4096-token prompt plus 64 decoding sequences at seq_len 4096 (chunked prefill through the paged cache: SURVEY.md §8 f-1,
`LlamaModel.forward(..., prefill_prefix_lens_list=...)`).  Not the headline bench (that is bench.py / configs[1]); one JSON
line with the mean step time over whole prompts (8 chunk positions each), tokens/s = (chunk + decodes) / step, and the
same schedule with the chunk and the decodes issued as two separate calls (what the reference would have to do).

    python scripts/sarathi_bench.py [--chunk 512] [--decodes 64] [--prompt 4096] [--prompts 3]
    torchrun --nproc-per-node 8 scripts/sarathi_bench.py --gpus 8 --model llama3-70b --prompt 8192 --seqlen 8192
        (BASELINE.json configs[4]: Llama-3-70B, TP = 8, mixed prefill/decode at seq_len 8192)
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--chunk", type=int, default=512)
    ap.add_argument("--decodes", type=int, default=64)
    ap.add_argument("--prompt", type=int, default=4096)
    ap.add_argument("--seqlen", type=int, default=4096)
    ap.add_argument("--prompts", type=int, default=3, help="timed prompts (each = prompt/chunk steps); one more runs as warm-up")
    ap.add_argument("--layers", type=int, default=0)
    ap.add_argument("--gpus", type=int, default=1, help="tensor-parallel degree (launch with torchrun --nproc-per-node N)")
    ap.add_argument("--model", type=str, default="llama3-8b", choices=["llama3-8b", "llama3-70b"])
    ap.add_argument("--exchange", type=str, default="auto", choices=["auto", "nccl", "one_shot", "two_shot", "two_shot_nvls", "ll", "ll_nvls"],
                    help="TP exchange (EngineConfig.fused_allreduce); auto = the library default for this TP degree")
    ap.add_argument("--shard-lm-head", action="store_true")
    args = ap.parse_args()
    import torch.distributed as dist
    import swiftllm_b200
    from swiftllm_b200.model_config import LLAMA3_8B, LLAMA3_70B
    from swiftllm_b200.worker.weight import synthetic_getter
    cfg = dict(LLAMA3_8B if args.model == "llama3-8b" else LLAMA3_70B)
    if args.layers:
        cfg["num_hidden_layers"] = args.layers
    mc = swiftllm_b200.LlamaModelConfig(cfg)
    n, rank, local = args.gpus, int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    assert int(os.environ.get("WORLD_SIZE", "1")) == n, "launch with torchrun --nproc-per-node <--gpus>"
    dev = torch.device("cuda", local)
    torch.cuda.set_device(local)
    if n > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    Bd, S, bs = args.decodes, args.seqlen, 16
    bps = (max(S, args.prompt) + bs - 1) // bs + 2
    ec = swiftllm_b200.EngineConfig(model_path="", use_dummy=False, block_size=bs, gpu_mem_utilization=0.97, num_cpu_blocks=0,
                                    max_seqs_in_block_table=Bd + 1, max_blocks_per_seq=bps, max_batch_size=Bd + 1,
                                    max_tokens_in_batch=args.chunk + Bd, dtype="bfloat16", tp_size=n, tp_rank=rank,
                                    fused_allreduce={"auto": None, "nccl": False, "one_shot": True}.get(args.exchange, args.exchange),
                                    shard_lm_head=args.shard_lm_head)
    with torch.inference_mode():
        m = swiftllm_b200.LlamaModel(ec, mc)
        m.load_weights(synthetic_getter(seed=0, std=0.02, device=dev))
        nblk = (Bd + 1) * bps
        m.init_kvcache_and_swap(nblk)
        g = torch.Generator(device=dev); g.manual_seed(3)
        for s in range(0, nblk, max(1, nblk // 16)):
            m.k_cache[s:s + max(1, nblk // 16)].normal_(generator=g); m.v_cache[s:s + max(1, nblk // 16)].normal_(generator=g)
        gen = torch.Generator().manual_seed(5)
        dec_ids = [[int(t)] for t in torch.randint(0, mc.vocab_size, (Bd,), generator=gen)]
        dec_sids = list(range(1, Bd + 1))
        prompt = torch.randint(0, mc.vocab_size, (args.prompt,), generator=gen).tolist()
        nchunks = (args.prompt + args.chunk - 1) // args.chunk

        def one_prompt(piggyback: bool):
            for c in range(nchunks):
                ids = prompt[c * args.chunk:(c + 1) * args.chunk]
                if piggyback:
                    m.forward([ids] + dec_ids, [0] + dec_sids, [S] * Bd, prefill_prefix_lens_list=[c * args.chunk])
                else:
                    m.forward([ids], [0], [], prefill_prefix_lens_list=[c * args.chunk])
                    m.forward(dec_ids, dec_sids, [S] * Bd)
            m.free_seqs_resources([0])

        def barrier():
            if n > 1:
                dist.barrier()
            torch.cuda.synchronize()

        def timed(piggyback):
            one_prompt(piggyback)
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.prompts):
                one_prompt(piggyback)
            e1.record(); barrier()
            ms = e0.elapsed_time(e1)
            if n > 1:                                             # device time, max over ranks
                t = torch.tensor([ms], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX); ms = float(t.item())
            return ms / (args.prompts * nchunks)

        m.forward(dec_ids, dec_sids, [S] * Bd)                    # allocate the decoding sequences' blocks once
        ms_pig, ms_sep = timed(True), timed(False)
    tok = args.chunk + Bd
    if rank != 0:
        torch.cuda.synchronize(); dist.barrier(); os._exit(0)
    print(json.dumps({"metric": "sarathi_step_tokens_per_s", "config": {"workload": f"{args.model} bf16, {args.chunk}-token chunk of a "
                      f"{args.prompt}-token prompt + {Bd} decodes at seq_len {S} per step (BASELINE.json configs[2] / [4])", "layers": cfg["num_hidden_layers"]},
                      "piggybacked": {"ms_per_step": ms_pig, "tokens_per_s": tok / (ms_pig * 1e-3)},
                      "separate_calls": {"ms_per_step": ms_sep, "tokens_per_s": tok / (ms_sep * 1e-3)},
                      "steps_timed": args.prompts * nchunks, "n_gpus": n, "parallelism": f"tp{n}", "data": "synthetic",
                      "exchange": args.exchange if m.comm is None or args.exchange != "auto" else ("ll" if getattr(m.comm, "ll", False) else "two_shot" if m.comm.two_shot else "one_shot"),
                      "lm_head_sharded": bool(args.shard_lm_head)}), flush=True)
    if n > 1:
        torch.cuda.synchronize(); dist.barrier(); os._exit(0)


if __name__ == "__main__":
    main()
