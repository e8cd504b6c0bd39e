"""One launch of the paged-decode kernel at the benchmarked geometry, for `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum`
(bench.py runs this under ncu as a subprocess to fill `roofline.traffic` live; it is never timed).

The KV cache here holds ONE layer ([blocks, 1, nkv, 16, D]: 2 x 4.3 GB at BASELINE configs[1]) instead of the 32 of the
benchmark, because the benchmark process still owns its 128 GiB cache while this runs.  Bytes per launch do not depend on the
layer count: the kernel reads the same sum(len) * nkv * D * 2 * 2 bytes of pages, block-table entries and q, and writes o.
Blocks are assigned to sequences through a random permutation so the page gather is as scattered as in a served cache."""
import argparse
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--seqlen", type=int, default=4096)
ap.add_argument("--nq", type=int, default=32)
ap.add_argument("--nkv", type=int, default=8)
ap.add_argument("--head-dim", type=int, default=128)
ap.add_argument("--dtype", default="bfloat16")
a = ap.parse_args()

from swiftllm_b200.worker.kernels.paged_attn import paged_attention  # noqa: E402

dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
dt = getattr(torch, a.dtype)
bs = 16
bps = (a.seqlen + bs - 1) // bs
nblk = a.batch * bps
g = torch.Generator(device=dev); g.manual_seed(3)
with torch.inference_mode():
    kc = torch.empty((nblk, 1, a.nkv, bs, a.head_dim), dtype=dt, device=dev).normal_(generator=g)
    vc = torch.empty((nblk, 1, a.nkv, bs, a.head_dim), dtype=dt, device=dev).normal_(generator=g)
    q = torch.empty((a.batch, a.nq, a.head_dim), dtype=dt, device=dev).normal_(generator=g)
    bt = torch.randperm(nblk, generator=g, device=dev).to(torch.int32).view(a.batch, bps).contiguous()
    st = types.SimpleNamespace(num_decoding_seqs=a.batch, num_prefill_seqs=0, seq_ids=torch.arange(a.batch, dtype=torch.int32, device=dev),
                               decoding_seq_lens=torch.full((a.batch,), a.seqlen, dtype=torch.int32, device=dev),
                               softmax_scale=a.head_dim ** -0.5, max_decoding_len=a.seqlen, paged_attn_seq_block_size=0)
    o = torch.zeros((a.batch, a.nq * a.head_dim), dtype=dt, device=dev)
    torch.cuda.synchronize()
    for _ in range(2):                 # ncu is told to skip the first launch (cold instruction cache, lazy module load)
        paged_attention(q, kc, vc, bt, None, types.SimpleNamespace(block_size=bs), st, 0, o)
    torch.cuda.synchronize()
assert torch.isfinite(o.float()).all()
print("ok")
