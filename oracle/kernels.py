"""
CPU restatements (torch-CPU / numpy) of every kernel in the reference's
`swiftllm/worker/kernels/`.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Each function follows the rounding order of the reference kernel it cites
(SURVEY.md Appendix A).  `h(.)` below means "round to the storage dtype"
(fp16 in the reference; bf16 in the bf16 build), `f(.)` means fp32.

All functions are functional (return new tensors) unless they end in
`_inplace`, in which case they mutate like the reference wrapper does.
"""
from __future__ import annotations

import math
import numpy as np
import torch

LOG2E = 1.442695040888963  # the literal used at paged_attn.py:193 / prefill_attn.py:122


# --------------------------------------------------------------------------
# A1/A2  RMSNorm            reference: swiftllm/worker/kernels/rmsnorm.py
# --------------------------------------------------------------------------
def rmsnorm(x: torch.Tensor, weight: torch.Tensor, eps: float) -> torch.Tensor:
    """rmsnorm.py:5-24: x32=f(x); var=sum(x32^2)/H; rstd=1/sqrt(var+eps);
    out = h((x32*rstd)*f(w))."""
    dt = x.dtype
    x32 = x.to(torch.float32)
    # the fp32 reduction order is unspecified by the reference ("tl.sum"); numpy's pairwise sum is what the
    # Triton interpreter used for the golden vectors, so use it to stay bit-identical with them
    var = torch.from_numpy(np.sum((x32 * x32).numpy(), axis=-1, keepdims=True) / np.float32(x.shape[-1]))
    rstd = 1.0 / torch.sqrt(var + eps)
    return ((x32 * rstd) * weight.to(torch.float32)).to(dt)


def fused_add_rmsnorm(x: torch.Tensor, residual: torch.Tensor, weight: torch.Tensor, eps: float):
    """rmsnorm.py:39-65: s = h(x + r) (add in the storage dtype, stored to the
    residual), then rmsnorm(s).  Returns (x_out, residual_out)."""
    s = x + residual           # storage-dtype add, correctly rounded (see DESIGN.md numerics)
    return rmsnorm(s, weight, eps), s


# --------------------------------------------------------------------------
# A3  SiLU-and-mul          reference: swiftllm/worker/kernels/silu_and_mul.py
# --------------------------------------------------------------------------
def silu_and_mul(x: torch.Tensor) -> torch.Tensor:
    """silu_and_mul.py:15-23.  x is [T, 2F] = [up | gate].  g = h(f(gate)/(1+exp(-f(gate))));
    out[:, :F] = h(up * g) (product in storage dtype); out[:, F:] = gate untouched."""
    dt = x.dtype
    F = x.shape[1] // 2
    gate = x[:, F:].to(torch.float32)
    # numpy's expf (what the Triton interpreter evaluates tl.exp with) - torch's vectorised expf differs
    # from it by 1 fp32 ulp on a few inputs, which flips the fp16 rounding of g about once per 10^4 elements
    e = torch.from_numpy(np.exp(-gate.numpy()))
    g = (gate / (1.0 + e)).to(dt)
    out = x.clone()
    out[:, :F] = x[:, :F] * g
    return out


# --------------------------------------------------------------------------
# A4  Rotary embedding      reference: swiftllm/worker/kernels/rotary_emb.py
# --------------------------------------------------------------------------
def rotary_embedding(q: torch.Tensor, k: torch.Tensor, cos: torch.Tensor, sin: torch.Tensor):
    """rotary_emb.py:28-41.  NeoX half-split pairing, all arithmetic in the storage
    dtype with every product and sum individually rounded (no FMA contraction):
    x0' = h(h(x0*c) - h(x1*s)); x1' = h(h(x0*s) + h(x1*c)).
    q [T,nq,D], k [T,nkv,D], cos/sin [T,D/2].  Returns (q_out, k_out)."""
    D = q.shape[-1]
    c = cos[:, None, :]
    s = sin[:, None, :]

    def rot(x):
        x0 = x[..., : D // 2]
        x1 = x[..., D // 2:]
        o0 = x0 * c - x1 * s
        o1 = x0 * s + x1 * c
        return torch.cat([o0, o1], dim=-1)

    return rot(q), rot(k)


def rope_tables(head_dim: int, rope_theta: float, max_position_embeddings: int, rope_scaling, dtype: torch.dtype):
    """model.py:177-225 (_init_to_get_rotary).  Returns (cos, sin) of shape
    [max_seq_len+128, D/2] in `dtype`, computed in fp32 on CPU."""
    base = rope_theta
    if isinstance(rope_scaling, dict):
        scaling_factor = rope_scaling.get("factor", 4.0)
        low = rope_scaling.get("low_freq_factor", 1.0)
        high = rope_scaling.get("high_freq_factor", 1.0)
        orig = rope_scaling.get("original_max_position_embeddings", max_position_embeddings)
        max_seq_len = int(orig * scaling_factor)
        t = torch.arange(max_seq_len + 128, dtype=torch.float32)
        dim_half = head_dim // 2
        split = int(dim_half * low / (low + high))
        inv_low = 1.0 / (base ** (torch.arange(0, split * 2, 2, dtype=torch.float32) / head_dim))
        inv_high = 1.0 / (base ** (torch.arange(split * 2, head_dim, 2, dtype=torch.float32) / head_dim))
        freqs = torch.cat([torch.outer(t / low, inv_low), torch.outer(t / high, inv_high)], dim=-1)
    else:
        factor = rope_scaling
        max_seq_len = max_position_embeddings * factor
        inv_freq = 1.0 / (base ** (torch.arange(0, head_dim, 2, dtype=torch.float32) / head_dim))
        t = torch.arange(int(max_seq_len + 128), dtype=torch.float32) / factor
        freqs = torch.outer(t, inv_freq)
    return torch.cos(freqs).to(dtype), torch.sin(freqs).to(dtype)


# --------------------------------------------------------------------------
# A7  KV-cache store        reference: swiftllm/worker/kernels/kvcache_mgmt.py
# --------------------------------------------------------------------------
def store_kvcache_inplace(k, v, k_cache, v_cache, block_table, seq_ids, prefill_seq_start_locs,
                          prefill_seq_lens, decoding_seq_lens, num_prefill_seqs, num_prefill_tokens,
                          block_size, cur_layer, prefill_prefix_lens=None):
    """kvcache_mgmt.py:10-79 (+ the commented torch loop at :124-132).
    k,v [T,nkv,D]; caches [num_blocks, L, nkv, bs, D]; block_table int32 [max_seqs, max_blocks_per_seq].
    Exact copies.
    `prefill_prefix_lens` (chunked prefill; not in the reference): token t of prefill chunk i goes to
    position prefill_prefix_lens[i] + t of its sequence (the reference always writes a prompt from position 0)."""
    seq_ids = [int(s) for s in seq_ids]
    for i in range(num_prefill_seqs):
        start = int(prefill_seq_start_locs[i])
        n = int(prefill_seq_lens[i])
        pre = 0 if prefill_prefix_lens is None else int(prefill_prefix_lens[i])
        sid = seq_ids[i]
        for t in range(n):
            blk = int(block_table[sid, (pre + t) // block_size])
            off = (pre + t) % block_size
            k_cache[blk, cur_layer, :, off, :] = k[start + t]
            v_cache[blk, cur_layer, :, off, :] = v[start + t]
    for j in range(len(decoding_seq_lens)):
        pos = int(decoding_seq_lens[j]) - 1
        sid = seq_ids[num_prefill_seqs + j]
        blk = int(block_table[sid, pos // block_size])
        off = pos % block_size
        k_cache[blk, cur_layer, :, off, :] = k[num_prefill_tokens + j]
        v_cache[blk, cur_layer, :, off, :] = v[num_prefill_tokens + j]


# --------------------------------------------------------------------------
# A5  Paged (decode) attention     reference: swiftllm/worker/kernels/paged_attn.py
# --------------------------------------------------------------------------
def _npsum(t: torch.Tensor, axis: int) -> torch.Tensor:
    """fp32 reduction in numpy's order (the order the Triton interpreter used for the golden vectors)."""
    return torch.from_numpy(np.asarray(np.sum(t.numpy(), axis=axis, dtype=np.float32)))


def _gather_kv(cache, block_table_row, seq_len, block_size, cur_layer):
    nblk = (seq_len + block_size - 1) // block_size
    blocks = [cache[int(block_table_row[b]), cur_layer] for b in range(nblk)]   # each [nkv, bs, D]
    return torch.cat(blocks, dim=1)[:, :seq_len, :]                             # [nkv, seq_len, D]


def paged_attention_exact(q, k_cache, v_cache, block_table, seq_ids, seq_lens, softmax_scale,
                          block_size, cur_layer, out_dtype=None):
    """The mathematical definition (the reference's commented torch loop,
    paged_attn.py:224-259) evaluated in fp64.  q [Bd,nq,D] -> o [Bd, nq*D] (fp64 unless out_dtype)."""
    Bd, nq, D = q.shape
    nkv = k_cache.shape[2]
    g = nq // nkv
    out = torch.zeros((Bd, nq, D), dtype=torch.float64)
    for i in range(Bd):
        L = int(seq_lens[i])
        row = block_table[int(seq_ids[i])]
        K = _gather_kv(k_cache, row, L, block_size, cur_layer).to(torch.float64).repeat_interleave(g, dim=0)
        V = _gather_kv(v_cache, row, L, block_size, cur_layer).to(torch.float64).repeat_interleave(g, dim=0)
        s = torch.einsum("hd,hld->hl", q[i].to(torch.float64), K) * softmax_scale
        p = torch.softmax(s, dim=-1)
        out[i] = torch.einsum("hl,hld->hd", p, V)
    out = out.reshape(Bd, nq * D)
    return out if out_dtype is None else out.to(out_dtype)


def paged_attention_fast(q, k_cache, v_cache, block_table, seq_ids, seq_lens, softmax_scale, block_size, cur_layer,
                         out_dtype=None):
    """Same definition as paged_attention_exact (paged_attn.py:224-259) as vectorised torch-eager fp32: pages gathered
    with index_select, GQA group handled by a batched matmul (no K/V replication).  This is the CPU baseline arm of
    bench.py (`cpu_baseline`, `--impl reference`)."""
    Bd, nq, D = q.shape
    nkv = k_cache.shape[2]
    g = nq // nkv
    out = torch.empty((Bd, nq, D), dtype=torch.float32)
    kl, vl = k_cache[:, cur_layer], v_cache[:, cur_layer]                  # [num_blocks, nkv, bs, D] views
    for i in range(Bd):
        L = int(seq_lens[i])
        nb = (L + block_size - 1) // block_size
        ids = torch.as_tensor(np.asarray(block_table[int(seq_ids[i])][:nb]), dtype=torch.long)
        K = kl.index_select(0, ids).permute(1, 0, 2, 3).reshape(nkv, nb * block_size, D)[:, :L].float()
        V = vl.index_select(0, ids).permute(1, 0, 2, 3).reshape(nkv, nb * block_size, D)[:, :L].float()
        s = torch.matmul(q[i].view(nkv, g, D).float(), K.transpose(1, 2)) * softmax_scale      # [nkv, g, L]
        p = torch.softmax(s, dim=-1)
        out[i] = torch.matmul(p, V).reshape(nq, D)
    out = out.reshape(Bd, nq * D)
    return out if out_dtype is None else out.to(out_dtype)


def paged_attention_phase1(q, k_cache, v_cache, block_table, seq_ids, seq_lens, softmax_scale,
                           block_size, cur_layer, seq_block_size, num_seq_blocks):
    """paged_attn.py:9-108 in the reference's rounding order (as executed by the
    Triton 3.6 interpreter):  per page  s_j = h( sum_d h(q_d*k_jd) ) where the sum over d is
    accumulated in fp32 and rounded once (numpy's half reduction), then
    s_j = h(s_j * h(scale*log2e)); tail -> -inf; online softmax in fp32 with exp2.
    Returns (mid_o fp32 [Bd,nq,nsb,D] normalised partials, lse2 fp32 [Bd,nq,nsb]); entries of
    splits that do not exist are left as NaN (the reference leaves them uninitialised)."""
    dt = q.dtype
    Bd, nq, D = q.shape
    nkv = k_cache.shape[2]
    g = nq // nkv
    scale_h = torch.tensor(softmax_scale * LOG2E, dtype=torch.float32).to(dt)   # tl.float16 kernel arg
    mid_o = torch.full((Bd, nq, num_seq_blocks, D), float("nan"), dtype=torch.float32)
    lse = torch.full((Bd, nq, num_seq_blocks), float("nan"), dtype=torch.float32)
    pages_per_split = seq_block_size // block_size
    for i in range(Bd):
        L = int(seq_lens[i])
        row = block_table[int(seq_ids[i])]
        for h in range(nq):
            kvh = h // g
            qh = q[i, h]
            for sb in range(num_seq_blocks):
                start_tok = sb * seq_block_size
                if start_tok >= L:
                    continue
                n_pages = min(pages_per_split, (L - start_tok + block_size - 1) // block_size)
                m = torch.tensor(-1e20, dtype=torch.float32)
                l = torch.tensor(0.0, dtype=torch.float32)
                acc = torch.zeros(D, dtype=torch.float32)
                for p in range(n_pages):
                    blk = int(row[sb * pages_per_split + p])
                    kb = k_cache[blk, cur_layer, kvh]          # [bs, D]
                    vb = v_cache[blk, cur_layer, kvh]
                    prod = qh[None, :] * kb                    # h(q*k)
                    s = _npsum(prod.to(torch.float32), 1).to(dt)   # fp32 accumulate, one rounding
                    s = s * scale_h
                    tok = start_tok + p * block_size + torch.arange(block_size)
                    s32 = torch.where(tok < L, s.to(torch.float32), torch.tensor(float("-inf")))
                    m_new = torch.maximum(m, s32.max())
                    pexp = torch.exp2(s32 - m_new)
                    alpha = torch.exp2(m - m_new)
                    acc = acc * alpha + _npsum(pexp[:, None] * vb.to(torch.float32), 0)
                    l = l * alpha + _npsum(pexp, 0)
                    m = m_new
                mid_o[i, h, sb] = acc / l
                lse[i, h, sb] = torch.log2(l) + m
    return mid_o, lse


def paged_attention_phase2(mid_o, lse, seq_lens, seq_block_size, out_dtype):
    """paged_attn.py:111-149: exp2-weighted merge of the valid splits; o = h(acc/l).  -> [Bd, nq*D]"""
    Bd, nq, nsb, D = mid_o.shape
    o = torch.zeros((Bd, nq, D), dtype=out_dtype)
    for i in range(Bd):
        n = (int(seq_lens[i]) + seq_block_size - 1) // seq_block_size
        for h in range(nq):
            m = torch.tensor(-1e20, dtype=torch.float32)
            l = torch.tensor(0.0, dtype=torch.float32)
            acc = torch.zeros(D, dtype=torch.float32)
            for sb in range(n):
                cur = lse[i, h, sb]
                m_new = torch.maximum(m, cur)
                old = torch.exp2(m - m_new)
                e = torch.exp2(cur - m_new)
                acc = acc * old + e * mid_o[i, h, sb]
                l = l * old + e
                m = m_new
            o[i, h] = (acc / l).to(out_dtype)
    return o.reshape(Bd, nq * D)


def paged_attention_ref_order(q, k_cache, v_cache, block_table, seq_ids, seq_lens, softmax_scale,
                              block_size, cur_layer, seq_block_size, num_seq_blocks):
    """paged_attn.py:152-222 = phase 1 + phase 2 in the reference's rounding order."""
    mid_o, lse = paged_attention_phase1(q, k_cache, v_cache, block_table, seq_ids, seq_lens, softmax_scale,
                                        block_size, cur_layer, seq_block_size, num_seq_blocks)
    return paged_attention_phase2(mid_o, lse, seq_lens, seq_block_size, q.dtype)


# --------------------------------------------------------------------------
# A6  Prefill attention     reference: swiftllm/worker/kernels/prefill_attn.py
#                           (and the flash_attn_varlen_func call, transformer_layer.py:86-96)
# --------------------------------------------------------------------------
def prefill_attention_exact(q, k, v, start_locs, seq_lens, softmax_scale, out_dtype=None, compute_dtype=torch.float64):
    """Causal varlen attention over packed [T,nq,D] q and [T,nkv,D] k/v, evaluated in `compute_dtype`
    (fp64 = the definition; fp32 = what the flash_attn stand-in of oracle/ref_shim.py computed when the
    model-level golden was generated).  -> [T, nq, D].  Rows not covered by any sequence are left zero."""
    T, nq, D = q.shape
    nkv = k.shape[1]
    g = nq // nkv
    cd = compute_dtype
    out = torch.zeros((T, nq, D), dtype=cd)
    for b in range(len(seq_lens)):
        s0, L = int(start_locs[b]), int(seq_lens[b])
        if L == 0:
            continue
        Q = q[s0:s0 + L].to(cd).transpose(0, 1)                          # [nq, L, D]
        K = k[s0:s0 + L].to(cd).transpose(0, 1).repeat_interleave(g, 0)
        V = v[s0:s0 + L].to(cd).transpose(0, 1).repeat_interleave(g, 0)
        s = torch.einsum("hqd,hkd->hqk", Q, K) * softmax_scale
        mask = torch.tril(torch.ones(L, L, dtype=torch.bool))
        s = s.masked_fill(~mask, float("-inf"))
        p = torch.softmax(s, dim=-1)
        out[s0:s0 + L] = torch.einsum("hqk,hkd->hqd", p, V).transpose(0, 1)
    return out if out_dtype is None else out.to(out_dtype)


def prefix_prefill_attention_exact(q, k_cache, v_cache, block_table, seq_ids, start_locs, chunk_lens, prefix_lens,
                                   softmax_scale, block_size, cur_layer, out_dtype=None, compute_dtype=torch.float64):
    """Chunked ("prefix-aware") prefill attention - SURVEY.md §8 f-1.  The reference has no such kernel (its prefill
    attention never reads the KV cache); the definition is the one that makes a prompt processed in
    chunks equal to the same prompt processed at once by `prefill_attention_exact`: chunk b holds the tokens at
    positions [prefix_lens[b], prefix_lens[b] + chunk_lens[b]) of sequence seq_ids[b]; the keys/values of ALL positions
    below prefix + chunk are read from the paged cache (the chunk's own K/V are stored first, like decode tokens,
    transformer_layer.py:70-71 order); query i attends to positions <= prefix + i.
    q [Tp, nq, D] packed chunks -> [Tp, nq, D]; rows not covered by any chunk stay zero."""
    T, nq, D = q.shape
    nkv = k_cache.shape[2]
    g = nq // nkv
    cd = compute_dtype
    out = torch.zeros((T, nq, D), dtype=cd)
    for b in range(len(chunk_lens)):
        s0, n, pre = int(start_locs[b]), int(chunk_lens[b]), int(prefix_lens[b])
        if n == 0:
            continue
        L = pre + n
        row = block_table[int(seq_ids[b])]
        K = _gather_kv(k_cache, row, L, block_size, cur_layer).to(cd).repeat_interleave(g, dim=0)       # [nq, L, D]
        V = _gather_kv(v_cache, row, L, block_size, cur_layer).to(cd).repeat_interleave(g, dim=0)
        Q = q[s0:s0 + n].to(cd).transpose(0, 1)                                                          # [nq, n, D]
        s = torch.einsum("hqd,hkd->hqk", Q, K) * softmax_scale
        qpos = pre + torch.arange(n)[:, None]
        s = s.masked_fill(torch.arange(L)[None, :] > qpos, float("-inf"))
        p = torch.softmax(s, dim=-1)
        out[s0:s0 + n] = torch.einsum("hqk,hkd->hqd", p, V).transpose(0, 1)
    return out if out_dtype is None else out.to(out_dtype)


def prefill_attention_ref_order(q, k, v, start_locs, seq_lens, softmax_scale, block_k: int = 128):
    """prefill_attn.py:52-100 rounding order: S=f(QK^T)*(scale*log2e) fp32; mask value -1e20;
    online softmax fp32/exp2 over K blocks; P rounded to the storage dtype before P.V; o=h(acc/l)."""
    dt = q.dtype
    T, nq, D = q.shape
    nkv = k.shape[1]
    g = nq // nkv
    sc = softmax_scale * LOG2E
    out = torch.zeros((T, nq, D), dtype=dt)
    for b in range(len(seq_lens)):
        s0, L = int(start_locs[b]), int(seq_lens[b])
        if L == 0:
            continue
        for h in range(nq):
            Q = q[s0:s0 + L, h].to(torch.float32)
            K = k[s0:s0 + L, h // g].to(torch.float32)
            V = v[s0:s0 + L, h // g]
            m = torch.full((L,), -1e20, dtype=torch.float32)
            l = torch.zeros(L, dtype=torch.float32)
            acc = torch.zeros((L, D), dtype=torch.float32)
            for ks in range(0, L, block_k):
                ke = min(L, ks + block_k)
                S = (Q @ K[ks:ke].T) * sc
                qi = torch.arange(L)[:, None]
                kj = torch.arange(ks, ke)[None, :]
                S = torch.where(qi >= kj, S, torch.tensor(-1e20))
                m_new = torch.maximum(m, S.max(dim=1).values)
                alpha = torch.exp2(m - m_new)
                P = torch.exp2(S - m_new[:, None])
                l = l * alpha + P.sum(dim=1)
                acc = acc * alpha[:, None] + P.to(dt).to(torch.float32) @ V[ks:ke].to(torch.float32)
                m = m_new
            out[s0:s0 + L, h] = (acc / l[:, None]).to(dt)
    return out


# --------------------------------------------------------------------------
# Block-table kernels        reference: swiftllm/worker/kernels/block_mgmt.py
# (numpy, exact integer work)
# --------------------------------------------------------------------------
def set_block_table_and_num_seq_alloc_blocks(num_seq_allocated_blocks, block_table, candidate_blocks,
                                             seq_ids, block_needed):
    """block_mgmt.py:5-46.  In place on numpy arrays."""
    cum = np.cumsum(block_needed)
    for i, sid in enumerate(seq_ids):
        sid = int(sid)
        need = int(block_needed[i])
        start = int(cum[i]) - need
        have = int(num_seq_allocated_blocks[sid])
        for j in range(need):
            block_table[sid, have + j] = candidate_blocks[start + j]
        num_seq_allocated_blocks[sid] = have + need


def unset_block_table_and_num_seq_alloc_blocks(num_seq_allocated_blocks, block_table, seq_ids, is_block_free):
    """block_mgmt.py:49-80.  In place."""
    for sid in seq_ids:
        sid = int(sid)
        n = int(num_seq_allocated_blocks[sid])
        for j in range(n):
            is_block_free[int(block_table[sid, j])] = True
        num_seq_allocated_blocks[sid] = 0


def gather_allocated_blocks_and_unset(num_seq_allocated_blocks, block_table, seq_ids, is_block_free):
    """block_mgmt.py:83-127.  Returns the gathered int32 block ids; in place otherwise."""
    out = []
    for sid in seq_ids:
        sid = int(sid)
        n = int(num_seq_allocated_blocks[sid])
        for j in range(n):
            b = int(block_table[sid, j])
            out.append(b)
            is_block_free[b] = True
        num_seq_allocated_blocks[sid] = 0
    return np.asarray(out, dtype=np.int32)


class BlockManagerOracle:
    """block_manager.py:5-103 restated on numpy.  Lowest-numbered free blocks first
    (torch.nonzero order, block_manager.py:50), handed to sequences in batch order."""

    def __init__(self, num_blocks, max_seqs_in_block_table, max_blocks_per_seq, block_size):
        self.num_blocks = num_blocks
        self.num_free_blocks = num_blocks
        self.block_size = block_size
        self.num_seq_allocated_blocks = np.zeros(max_seqs_in_block_table, dtype=np.int32)
        self.block_table = np.full((max_seqs_in_block_table, max_blocks_per_seq), -1, dtype=np.int32)
        self.is_block_free = np.ones(num_blocks, dtype=bool)

    def allocate_blocks_for_seqs(self, seq_ids, target_lens):
        seq_ids = np.asarray(seq_ids, dtype=np.int64)
        target = (np.asarray(target_lens, dtype=np.int64) + self.block_size - 1) // self.block_size
        have = self.num_seq_allocated_blocks[seq_ids]
        assert (have <= target).all()
        need = (target - have).astype(np.int64)
        n = int(need.sum())
        if n > self.num_free_blocks:
            raise RuntimeError("No enough free blocks available")
        selected = np.nonzero(self.is_block_free)[0][:n]
        self.num_free_blocks -= n
        self.is_block_free[selected] = False
        set_block_table_and_num_seq_alloc_blocks(self.num_seq_allocated_blocks, self.block_table,
                                                 selected, seq_ids, need)
        return selected.astype(np.int64)

    def free_blocks_for_seqs(self, seq_ids):
        seq_ids = np.asarray(seq_ids, dtype=np.int64)
        self.num_free_blocks += int(self.num_seq_allocated_blocks[seq_ids].sum())
        unset_block_table_and_num_seq_alloc_blocks(self.num_seq_allocated_blocks, self.block_table,
                                                   seq_ids, self.is_block_free)

    def gather_allocated_blocks_and_free(self, seq_ids):
        ids = gather_allocated_blocks_and_unset(self.num_seq_allocated_blocks, self.block_table,
                                                np.asarray(seq_ids, dtype=np.int64), self.is_block_free)
        self.num_free_blocks += len(ids)
        return ids


# --------------------------------------------------------------------------
# swap_blocks                reference: csrc/src/block_swapping.cpp:22-85
# --------------------------------------------------------------------------
def swap_blocks_inplace(src_ids, dst_ids, is_swap_in, k_cache, v_cache, k_swap, v_swap):
    """Copy whole blocks (all layers/heads) between the GPU cache and the CPU swap space.
    The run-coalescing in the reference (block_swapping.cpp:36-41) does not change results."""
    for s, d in zip(src_ids, dst_ids):
        s, d = int(s), int(d)
        if is_swap_in:
            k_cache[d] = k_swap[s]
            v_cache[d] = v_swap[s]
        else:
            k_swap[d] = k_cache[s]
            v_swap[d] = v_cache[s]


def coalesce_runs(src_ids, dst_ids):
    """The segmenting rule of block_swapping.cpp:33-44: maximal runs where both id lists
    advance by exactly +1.  Returns [(src_start, dst_start, length)]."""
    runs = []
    i, n = 0, len(src_ids)
    while i < n:
        j = i + 1
        while j < n and src_ids[j] == src_ids[j - 1] + 1 and dst_ids[j] == dst_ids[j - 1] + 1:
            j += 1
        runs.append((int(src_ids[i]), int(dst_ids[i]), j - i))
        i = j
    return runs


# --------------------------------------------------------------------------
# seq_block_size heuristic   reference: swiftllm/worker/model.py:305-324
# --------------------------------------------------------------------------
def select_seq_block_size(num_kv_heads: int, decoding_seq_lens: list[int]):
    seq_block_size = 2048
    total = sum(decoding_seq_lens)
    max_len = max(decoding_seq_lens) if decoding_seq_lens else 0
    while num_kv_heads * (total / seq_block_size) < 1024 and seq_block_size // 2 >= 64 and \
            max_len / (seq_block_size // 2) <= 128:
        seq_block_size //= 2
    return seq_block_size, (max_len + seq_block_size - 1) // seq_block_size
