"""CPU: a numpy MIRROR of the index / masking logic of the two PAGED prefill kernels (csrc/prefill_attn_tc_kernel.cuh and
csrc/prefill_attn_kernel.cuh with PAGED = true), step for step as the CUDA code does it - which kv steps a CTA visits, which
pages a step loads (whole pages for the TMA kernel, row-wise zero fill for the cp.async kernel), when the causal / tail mask is
applied (`need_mask`) and with which limit (`lim`), which warps skip a tile, which rows are written - with the tensor-core
products replaced by fp64 matmuls.  It is checked against the fp64 definition (oracle/kernels.py) on random ragged
(prefix, chunk) layouts, with NaN in every K slot and a large finite value in every V slot that must not contribute.

This cannot validate the CUDA code (that needs a GPU: tests/test_chunked_prefill_gpu.py); it validates the ARITHMETIC OF THE
INDICES that was changed without a GPU at hand (prefix offsets in nkt / need_mask / lim / skip), so an off-by-one there shows
up here first.  What the mirror also documents: slots of an owned page past the sequence end are multiplied by P = 0 in the TMA
kernel, so they must hold finite values (they do: the cache is zero-initialised and only ever receives finite K/V)."""
import numpy as np
import pytest
import torch

from oracle import kernels as K

BS = 16


def _gather_page(cache, blk, layer, kvh):
    return cache[blk, layer, kvh].double().numpy()          # [16, D]


def mirror_tc(q, kc, vc, bt, sids, starts, chunk, prefix, scale, layer):
    """prefill_attn_tc_kernel<T, PAGED=true>: CTA = (seq, head, 256 query rows = tiles A and B), 64-token kv steps."""
    Tn, nq, D = q.shape
    nkv = kc.shape[2]
    out = np.zeros((Tn, nq, D))
    BQ, BK = 128, 64
    for seq in range(len(chunk)):
        ln, pre, tok0 = chunk[seq], prefix[seq], starts[seq]
        kvlen = pre + ln
        npages = (kvlen + BS - 1) // BS
        for head in range(nq):
            kvh = head // (nq // nkv)
            for qblk in range((max(chunk) + 2 * BQ - 1) // (2 * BQ)):
                q0 = qblk * 2 * BQ
                if q0 >= ln:
                    continue
                active_b = q0 + BQ < ln
                nkt = [(min(kvlen, pre + q0 + BQ) + BK - 1) // BK, (min(kvlen, pre + q0 + 2 * BQ) + BK - 1) // BK if active_b else 0]
                for x in range(2):
                    rows = q0 + x * BQ + np.arange(BQ)                 # chunk-relative query rows of this tile
                    S_all, V_all = [], []
                    for j in range(nkt[x]):
                        # producer: pages 4j .. 4j+3 that exist are loaded WHOLE; the rest of the stage is zero (or stale, finite)
                        Kt, Vt = np.zeros((BK, D)), np.zeros((BK, D))
                        for pg in range(min(4, npages - 4 * j)):
                            blk = int(bt[sids[seq], 4 * j + pg])
                            Kt[pg * BS:(pg + 1) * BS] = _gather_page(kc, blk, layer, kvh)
                            Vt[pg * BS:(pg + 1) * BS] = _gather_page(vc, blk, layer, kvh)
                        qrow = np.zeros((BQ, D))
                        ok = tok0 + rows < Tn
                        qrow[ok] = q[tok0 + rows[ok], head].double().numpy()      # TMA: rows past the tensor are zero-filled
                        with np.errstate(invalid="ignore"):
                            S = qrow @ Kt.T
                        c0 = j * BK
                        need_mask = c0 + BK - 1 > pre + q0 + x * BQ or c0 + BK > kvlen
                        if need_mask:
                            lim = np.minimum(pre + rows, kvlen - 1) - c0
                            S = np.where(np.arange(BK)[None, :] <= lim[:, None], S, -np.inf)
                        S_all.append(S); V_all.append(Vt)
                    if not S_all:
                        continue
                    S = np.concatenate(S_all, 1) * scale
                    Vt = np.concatenate(V_all, 0)
                    P = np.exp(S - S.max(1, keepdims=True))
                    with np.errstate(invalid="ignore"):
                        O = (P @ Vt) / P.sum(1, keepdims=True)         # 0 * NaN = NaN here exactly as on the tensor core
                    wr = rows < ln                                      # epilogue: only rows inside the chunk are written
                    out[tok0 + rows[wr], head] = O[wr]
    return out


def mirror_mma(q, kc, vc, bt, sids, starts, chunk, prefix, scale, layer, bs):
    """prefill_attn_kernel<T, D, PAGED=true>: CTA = (seq, head, 128 query rows), 8 warps x 16 rows, 64-token kv tiles whose
    rows are gathered one by one (rows at positions >= kvlen are zero-filled), per-warp tile skipping above the diagonal."""
    Tn, nq, D = q.shape
    nkv = kc.shape[2]
    out = np.zeros((Tn, nq, D))
    BQ, BK = 128, 64
    for seq in range(len(chunk)):
        ln, pre, tok0 = chunk[seq], prefix[seq], starts[seq]
        kvlen = pre + ln
        for head in range(nq):
            kvh = head // (nq // nkv)
            for qb in range((max(chunk) + BQ - 1) // BQ):
                q0 = qb * BQ
                if q0 >= ln:
                    continue
                ntiles = (min(kvlen, pre + q0 + BQ) + BK - 1) // BK
                for warp in range(8):
                    rows = q0 + warp * 16 + np.arange(16)
                    qrow = np.zeros((16, D))
                    ok = rows < ln                                       # Q rows past the chunk are zero-filled
                    qrow[ok] = q[tok0 + rows[ok], head].double().numpy()
                    S_all, V_all = [], []
                    for tile in range(ntiles):
                        kt0 = tile * BK
                        if kt0 > pre + q0 + warp * 16 + 15:              # warp-uniform skip
                            continue
                        Kt, Vt = np.zeros((BK, D)), np.zeros((BK, D))
                        for r in range(BK):
                            t = kt0 + r
                            if t < kvlen:
                                blk = int(bt[sids[seq], t // bs])
                                Kt[r] = kc[blk, layer, kvh, t % bs].double().numpy()
                                Vt[r] = vc[blk, layer, kvh, t % bs].double().numpy()
                        with np.errstate(invalid="ignore"):
                            S = qrow @ Kt.T
                        need_mask = kt0 + BK - 1 > pre + q0 + warp * 16 or kt0 + BK > kvlen
                        if need_mask:
                            cc = kt0 + np.arange(BK)[None, :]
                            rr = (pre + rows)[:, None]
                            S = np.where((cc > rr) | (cc >= kvlen), -np.inf, S)
                        S_all.append(S); V_all.append(Vt)
                    S = np.concatenate(S_all, 1) * scale
                    Vt = np.concatenate(V_all, 0)
                    m = np.maximum(S.max(1, keepdims=True), -1e30)
                    P = np.exp(S - m)
                    with np.errstate(invalid="ignore", divide="ignore"):
                        O = (P @ Vt) / P.sum(1, keepdims=True)
                    wr = pre + rows < kvlen
                    out[tok0 + rows[wr], head] = O[wr]
    return out


def _case(rng, bs, D, nq, nkv, n_seqs):
    prefix = [int(rng.integers(0, 300)) if rng.random() < 0.8 else 0 for _ in range(n_seqs)]
    chunk = [int(rng.integers(1, 330)) for _ in range(n_seqs)]
    need = [(p + c + bs - 1) // bs for p, c in zip(prefix, chunk)]
    nblk = sum(need) + 2
    perm = rng.permutation(nblk)
    sids = rng.permutation(n_seqs + 2)[:n_seqs].tolist()
    bt = np.full((n_seqs + 2, max(need) + 1), -1, dtype=np.int32)
    p = 0
    for s, n in zip(sids, need):
        bt[s, :n] = perm[p:p + n]; p += n
    starts = list(np.cumsum([0] + chunk[:-1]))
    Tn = sum(chunk) + 3
    q = torch.from_numpy(rng.standard_normal((Tn, nq, D)).astype(np.float32)).half()
    kc = torch.from_numpy(rng.standard_normal((nblk, 2, nkv, bs, D)).astype(np.float32)).half()
    vc = torch.from_numpy(rng.standard_normal((nblk, 2, nkv, bs, D)).astype(np.float32)).half()
    # poison every slot that must not contribute: K with NaN, V with a huge finite value (see the module docstring)
    for s, n, pc in zip(sids, need, [a + b for a, b in zip(prefix, chunk)]):
        tail = pc - (n - 1) * bs
        if tail < bs:
            kc[int(bt[s, n - 1]), :, :, tail:] = float("nan")
            vc[int(bt[s, n - 1]), :, :, tail:] = 3.0e4
    owned = {int(b) for s, n in zip(sids, need) for b in bt[s, :n]}
    for b in range(nblk):
        if b not in owned:
            kc[b] = float("nan"); vc[b] = float("nan")
    return q, kc, vc, bt, sids, starts, chunk, prefix


@pytest.mark.parametrize("seed", range(4))
def test_mirror_of_the_tcgen05_paged_prefill_kernel_matches_the_definition(seed):
    rng = np.random.default_rng(seed)
    q, kc, vc, bt, sids, starts, chunk, prefix = _case(rng, BS, 32, 4, 2, 3)
    ref = K.prefix_prefill_attention_exact(q, kc, vc, bt, sids, starts, chunk, prefix, 32 ** -0.5, BS, 1).numpy()
    got = mirror_tc(q, kc, vc, bt, sids, starts, chunk, prefix, 32 ** -0.5, 1)
    assert np.isfinite(got).all()
    assert np.abs(got - ref).max() < 1e-9


@pytest.mark.parametrize("seed,bs", [(0, 16), (1, 16), (2, 32), (3, 4)])
def test_mirror_of_the_mma_sync_paged_prefill_kernel_matches_the_definition(seed, bs):
    rng = np.random.default_rng(100 + seed)
    q, kc, vc, bt, sids, starts, chunk, prefix = _case(rng, bs, 32, 4, 2, 3)
    ref = K.prefix_prefill_attention_exact(q, kc, vc, bt, sids, starts, chunk, prefix, 32 ** -0.5, bs, 1).numpy()
    got = mirror_mma(q, kc, vc, bt, sids, starts, chunk, prefix, 32 ** -0.5, 1, bs)
    assert np.isfinite(got).all()
    assert np.abs(got - ref).max() < 1e-9


# ----------------------------------------------------------------------------- store_kv_prefill_kernel<T, PREFIX = true>
def mirror_store_prefix(k, kc, bt, sids, starts, chunk, prefix, layer, bs):
    """csrc/kvcache.cu: grid (cdiv(max_chunk, bs) + 1, num_seqs); CTA (x, b) fills page prefix_b / bs + x of sequence b with the
    chunk tokens that fall into it.  Returns how often every cache element was written."""
    nkv, D = k.shape[1], k.shape[2]
    hits = np.zeros(kc.shape, dtype=np.int32)
    for b in range(len(chunk)):
        for x in range((max(chunk) + bs - 1) // bs + 1):
            pre, ln = prefix[b], chunk[b]
            pb = x + pre // bs
            lo, hi = max(pb * bs, pre), min(pb * bs + bs, pre + ln)
            if lo >= hi:
                continue
            off0, ntok, row0 = lo - pb * bs, hi - lo, starts[b] + (lo - pre)
            blk = int(bt[sids[b], pb])
            for i in range(ntok * nkv * (D // 8)):                  # the thread loop (order irrelevant)
                c, h, t = i % (D // 8), (i // (D // 8)) % nkv, i // ((D // 8) * nkv)
                kc[blk, layer, h, off0 + t, 8 * c:8 * c + 8] = k[row0 + t, h, 8 * c:8 * c + 8]
                hits[blk, layer, h, off0 + t, 8 * c:8 * c + 8] += 1
    return hits


@pytest.mark.parametrize("bs", [4, 16])
def test_mirror_of_the_prefix_store_kernel_writes_every_slot_exactly_once(bs):
    rng = np.random.default_rng(bs)
    for _ in range(5):
        n = 4
        prefix = [int(rng.integers(0, 60)) for _ in range(n)]
        chunk = [int(rng.integers(1, 50)) for _ in range(n)]
        need = [(p + c + bs - 1) // bs for p, c in zip(prefix, chunk)]
        perm = rng.permutation(sum(need) + 1)
        bt = np.full((n, max(need) + 1), -1, dtype=np.int32); p0 = 0
        for s, m in enumerate(need):
            bt[s, :m] = perm[p0:p0 + m]; p0 += m
        starts = list(np.cumsum([0] + chunk[:-1]))
        k = torch.from_numpy(rng.standard_normal((sum(chunk), 2, 16)).astype(np.float32)).half()
        kc = torch.zeros(sum(need) + 1, 2, 2, bs, 16, dtype=torch.float16)
        ko, vo = torch.zeros_like(kc), torch.zeros_like(kc)
        hits = mirror_store_prefix(k, kc, bt, list(range(n)), starts, chunk, prefix, 1, bs)
        K.store_kvcache_inplace(k, k, ko, vo, bt, list(range(n)), starts, chunk, [], n, sum(chunk), bs, 1, prefill_prefix_lens=prefix)
        assert torch.equal(kc, ko)
        assert hits.max() == 1 and hits.sum() == sum(chunk) * 2 * 16          # every (token, head, d) written exactly once


# ----------------------------------------------------------------------------- swap_blocks_gather_kernel
def test_mirror_of_the_swap_gather_kernel_covers_every_vector_exactly_once():
    """csrc/swap_gather.cu: grid (n, SW_SPLIT = 8), 256 threads, two 16-byte vectors per thread and iteration."""
    SPLIT, THREADS = 8, 256
    for block_vecs in (1, 7, 255, 256, 257, 4096, 65536, 65536 + 13):
        hits = np.zeros(block_vecs, dtype=np.int32)
        per = (block_vecs + SPLIT - 1) // SPLIT
        for y in range(SPLIT):
            lo, hi = y * per, min(block_vecs, y * per + per)
            for tid in range(THREADS):
                i = lo + tid
                while i < hi:
                    hits[i] += 1
                    j = i + THREADS
                    if j < hi:
                        hits[j] += 1
                    i += 2 * THREADS
        assert (hits == 1).all(), block_vecs
