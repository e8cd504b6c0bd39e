"""Timeline of ONE CTA of the tcgen05 prefill kernel (development tool; needs the trace variant of the library):

    SLLM_BUILD_VARIANT=pptrace SLLM_NVCC_EXTRA="-DSLLM_PT_PINGPONG -DSLLM_PT_TRACE" python -m swiftllm_b200.build   # -> libswiftllm_b200_pptrace.so
    python scripts/prefill_trace.py

The traced CTA (middle query block of sequence 1, head 3) stamps clock64() whenever one role hands over to another:
  MMA issuer   1/2 = S_A(j) issued (begin/end), 7/8 = S_B(j), 3/4 = P.V of tile A issued, 5/6 = P.V of tile B
  softmax A/B  9 = waiting for S(j), 10 = S(j) complete, 11 = S loaded from TMEM (S buffer released), 12 = max / rescale done,
               13 = P buffer free and my turn on the MUFU pipe, 14 = exp + pack + store done, 15 = P published
Prints the mean time of every phase over the steady-state steps and a few raw steps, in SM cycles."""
import ctypes, os, sys, types
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
os.environ["SLLM_LIB_PATH"] = os.environ.get("PT_TRACE_LIB") or os.path.join(ROOT, "swiftllm_b200", "libswiftllm_b200_pptrace.so")
os.environ["SLLM_PREFILL_ATTN_GEN"] = "0"
import numpy as np
import torch
from swiftllm_b200 import _lib
from swiftllm_b200.worker.kernels.prefill_attn import prefill_attention

Bp, L, nq, nkv, D = 8, 4096, 32, 8, 128
T = Bp * L
g = torch.Generator(device="cuda"); g.manual_seed(0)
q = torch.randn(T, nq, D, device="cuda", dtype=torch.bfloat16, generator=g)
k = torch.randn(T, nkv, D, device="cuda", dtype=torch.bfloat16, generator=g)
v = torch.randn(T, nkv, D, device="cuda", dtype=torch.bfloat16, generator=g)
o = torch.empty_like(q)
st = types.SimpleNamespace(num_prefill_seqs=Bp, prefill_seq_start_locs=torch.arange(Bp, device="cuda", dtype=torch.int32) * L,
                           prefill_seq_lens=torch.full((Bp,), L, device="cuda", dtype=torch.int32), max_prefill_len=L, softmax_scale=D ** -0.5)
for _ in range(3):
    prefill_attention(q, k, v, o, None, None, st)
torch.cuda.synchronize()
buf = np.zeros(8 * 2048, dtype=np.uint64)
fn = _lib.lib().sllm_debug_prefill_trace
fn.restype = ctypes.c_int; fn.argtypes = [ctypes.c_void_p, ctypes.c_longlong]
assert fn(buf.ctypes.data, buf.nbytes) == 0
roles = {0: "issuer", 1: "softmax A", 2: "softmax B"}
ev = {}
for r in roles:
    x = buf[r * 2048:(r + 1) * 2048]
    x = x[x != 0]
    ev[r] = [(int(e >> 8), int(e & 0xff)) for e in x]
t0 = min(e[0][0] for e in ev.values() if e)
print({roles[r]: len(e) for r, e in ev.items()})


def phases(seq, order):
    """seq: [(t, tag)], order: the tags of one step in order -> list of per-step dicts tag -> t."""
    out, cur = [], {}
    for t, tag in seq:
        if tag == order[0] and cur:
            out.append(cur); cur = {}
        cur[tag] = t
    if cur:
        out.append(cur)
    return [s for s in out if all(tg in s for tg in order)]


for r in (1, 2):
    steps = phases(ev[r], [9, 10, 11, 12, 13, 14, 15])
    mid = steps[len(steps) // 4: 3 * len(steps) // 4]
    names = ["wait S(j)", "TMEM load S", "max / rescale", "wait P buffer", "exp + pack + store", "fence + publish"]
    tags = [9, 10, 11, 12, 13, 14, 15]
    print(f"{roles[r]}: {len(steps)} steps; mean cycles per phase over the middle half:")
    for i, nme in enumerate(names):
        print(f"    {nme:22s} {np.mean([s[tags[i + 1]] - s[tags[i]] for s in mid]):8.0f}")
    print(f"    {'step period':22s} {np.mean([b[9] - a[9] for a, b in zip(mid, mid[1:])]):8.0f}")
iss = ev[0]
def spans(b_tag, e_tag):
    return [(a, b) for (a, ta), (b, tb) in zip(iss, iss[1:]) if ta == b_tag and tb == e_tag]
for nme, bt, et in (("S_A", 1, 2), ("S_B", 7, 8), ("PV_A", 3, 4), ("PV_B", 5, 6)):
    sp = spans(bt, et); h = len(sp) // 4
    if sp:
        print(f"issuer {nme}: {len(sp)} issues, {np.mean([b - a for a, b in sp][h:3 * h]):.0f} cycles each, period {np.mean(np.diff([a for a, _ in sp])[h:3 * h]):.0f}")
print("raw events of 3 steady-state steps (cycles since the CTA's first event):")
allev = sorted([(t - t0, roles[r], tag) for r, e in ev.items() for t, tag in e])
mid_t = allev[len(allev) // 2][0]
for t, rn, tag in allev:
    if mid_t <= t < mid_t + 9000:
        print(f"   {t:8d}  {rn:10s} {tag}")
