"""HBM roofline of the elementwise / copy kernels at prefill-size T (Llama-3-8B shapes, bf16): achieved GB/s =
algorithmic bytes / CUDA-event time, against the measured copy peak.  One JSON line per kernel."""
import json, os, sys, types
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from swiftllm_b200.worker.kernels.rmsnorm import rmsnorm_inplace, fused_add_rmsnorm_inplace
from swiftllm_b200.worker.kernels.rotary_emb import rotary_embedding_inplace
from swiftllm_b200.worker.kernels.silu_and_mul import silu_and_mul_inplace
from swiftllm_b200.worker.kernels.kvcache_mgmt import store_kvcache

T = int(os.environ.get("EW_T", 32768)); H, nq, nkv, D, F, L, bs = 4096, 32, 8, 128, 14336, 2, 16
dt, dev = torch.bfloat16, "cuda"
peak = 6575.4
try:
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
g = torch.Generator(device=dev); g.manual_seed(0)
rn = lambda *s: torch.randn(*s, device=dev, dtype=dt, generator=g)

ONCE = os.environ.get("EW_ONCE", "0") == "1"      # profiling aid: one launch per kernel (ncu -c 5 captures each once)


def timed(fn, iters=10):
    if ONCE:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1)
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters

def report(name, nbytes, ms):
    gbs = nbytes / (ms * 1e-3) / 1e9
    print(json.dumps({"kernel": name, "T": T, "algorithmic_bytes": nbytes, "ms": ms, "achieved_GBps": gbs, "peak_GBps": peak,
                      "frac_of_measured_peak": gbs / peak}), flush=True)

s = 2
x, r, w = rn(T, H), rn(T, H), rn(H)
report("rmsnorm_kernel<bf16,0> (rmsnorm_inplace)", 2 * T * H * s + H * s, timed(lambda: rmsnorm_inplace(x, w, 1e-5)))
report("rmsnorm_kernel<bf16,1> (fused_add_rmsnorm_inplace)", 4 * T * H * s + H * s, timed(lambda: fused_add_rmsnorm_inplace(x, r, w, 1e-5)))
q, k = rn(T, nq, D), rn(T, nkv, D)
st = types.SimpleNamespace(position_cos=rn(T, D // 2), position_sin=rn(T, D // 2))
report("rotary_kernel<bf16>", 2 * T * (nq + nkv) * D * s + 2 * T * (D // 2) * s, timed(lambda: rotary_embedding_inplace(q, k, st)))
ug = rn(T, 2 * F)
report("silu_and_mul_kernel<bf16>", 3 * T * F * s, timed(lambda: silu_and_mul_inplace(ug)))
Bp, Lp = T // 4096, 4096
nblk = Bp * (Lp // bs)
kc = torch.zeros(nblk, L, nkv, bs, D, device=dev, dtype=dt); vc = torch.zeros_like(kc)
bt = torch.arange(nblk, device=dev, dtype=torch.int32).view(Bp, Lp // bs)
v = rn(T, nkv, D)
i32 = lambda a: torch.tensor(a, dtype=torch.int32, device=dev)
sk = types.SimpleNamespace(seq_ids=i32(list(range(Bp))), num_prefill_seqs=Bp, num_decoding_seqs=0, num_prefill_tokens=T,
                           prefill_seq_start_locs=i32([i * Lp for i in range(Bp)]), prefill_seq_lens=i32([Lp] * Bp), max_prefill_len=Lp,
                           decoding_seq_lens=i32([]))
report("store_kv_prefill_kernel<bf16>", 4 * T * nkv * D * s, timed(lambda: store_kvcache(k, v, kc, vc, bt, None, None, sk, 1)))
