"""Kernel-level prefill attention throughput (TFLOP/s) of both generations, Llama-3-8B head geometry, bf16.
FLOPs = 4 * nq * D * sum_i L_i (L_i + 1) / 2 (causal).  Prints one JSON line per (generation, shape)."""
import json, os, sys, types
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from swiftllm_b200.worker.kernels.prefill_attn import prefill_attention

def run(gen, Bp, L, nq=32, nkv=8, D=128, iters=10):
    os.environ["SLLM_PREFILL_ATTN_GEN"] = gen
    T = Bp * L
    g = torch.Generator(device="cuda"); g.manual_seed(0)
    q = torch.randn(T, nq, D, device="cuda", dtype=torch.bfloat16, generator=g)
    k = torch.randn(T, nkv, D, device="cuda", dtype=torch.bfloat16, generator=g)
    v = torch.randn(T, nkv, D, device="cuda", dtype=torch.bfloat16, generator=g)
    o = torch.empty_like(q)
    st = types.SimpleNamespace(num_prefill_seqs=Bp, prefill_seq_start_locs=torch.arange(Bp, device="cuda", dtype=torch.int32) * L,
                               prefill_seq_lens=torch.full((Bp,), L, device="cuda", dtype=torch.int32), max_prefill_len=L,
                               softmax_scale=D ** -0.5)
    for _ in range(3):
        prefill_attention(q, k, v, o, None, None, st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        prefill_attention(q, k, v, o, None, None, st)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    flops = 4 * nq * D * Bp * L * (L + 1) / 2
    return {"gen": "gen1-mma.sync" if gen == "1" else "gen2-tcgen05", "Bp": Bp, "L": L, "ms": ms, "tflops": flops / (ms * 1e-3) / 1e12}, o

peak = 1673.7
for Bp, L in ((8, 4096), (2, 16384), (32, 1024)):
    r1, o1 = run("1", Bp, L)
    r2, o2 = run("0", Bp, L)
    r2["max_abs_diff_vs_gen1"] = float((o1.float() - o2.float()).abs().max())
    for r in (r1, r2):
        r["frac_of_measured_bf16_peak"] = r["tflops"] / peak
        print(json.dumps(r), flush=True)
