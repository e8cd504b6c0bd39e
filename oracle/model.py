"""
CPU torch-eager restatement of the reference data plane `swiftllm.worker.model.LlamaModel`
(model.py:18-408) and its layers (layers/pre_layer.py, transformer_layer.py, post_layer.py),
built on the kernel restatements in `oracle/kernels.py`.  TEST INFRASTRUCTURE ONLY: it is the
parity checker for the CUDA path and the `cpu_baseline` / `--impl reference` arm of bench.py.

GEMMs use torch's CPU `F.linear` (the reference uses cuBLAS via the same call, linear.py:12:
third-party arithmetic, "parity unpinned"); greedy argmax is `torch.argmax` (post_layer.py:39).
"""
from __future__ import annotations

import itertools
import numpy as np
import torch
import torch.nn.functional as F

from . import kernels as K


class OracleWeights:
    """Same attribute names as the reference's LlamaWeight / LlamaTransformerLayerWeight (weight.py:56-177):
    wte, lm_head, final_norm, layers[i].{attn_norm,q_proj,k_proj,v_proj,o_proj,ffn_norm,up_gate_proj,down_proj}
    with up_gate_proj = cat(up, gate) (weight.py:133)."""

    class Layer:
        pass

    def __init__(self, num_layers):
        self.layers = [OracleWeights.Layer() for _ in range(num_layers)]

    @staticmethod
    def random(cfg: dict, dtype=torch.float16, seed=0, std=0.02):
        """Seeded synthetic weights (N(0,std) projections, 1+N(0,std) norms)."""
        g = torch.Generator().manual_seed(seed)
        H, nq = cfg["hidden_size"], cfg["num_attention_heads"]
        nkv = cfg.get("num_key_value_heads", nq)
        D = H // nq
        Fd, V, L = cfg["intermediate_size"], cfg["vocab_size"], cfg["num_hidden_layers"]
        w = OracleWeights(L)
        rn = lambda *s: (torch.randn(*s, generator=g) * std).to(dtype)
        nr = lambda n: (1 + torch.randn(n, generator=g) * std).to(dtype)
        w.wte, w.lm_head, w.final_norm = rn(V, H), rn(V, H), nr(H)
        for lw in w.layers:
            lw.attn_norm, lw.ffn_norm = nr(H), nr(H)
            lw.q_proj, lw.k_proj, lw.v_proj = rn(H, H), rn(nkv * D, H), rn(nkv * D, H)
            lw.o_proj, lw.up_gate_proj, lw.down_proj = rn(H, H), rn(2 * Fd, H), rn(H, Fd)
        return w

    @staticmethod
    def from_golden(z, num_layers):
        w = OracleWeights(num_layers)
        w.wte, w.lm_head, w.final_norm = (torch.from_numpy(z[n]) for n in ("wte", "lm_head", "final_norm"))
        for i, lw in enumerate(w.layers):
            for n in ("attn_norm", "ffn_norm", "q_proj", "k_proj", "v_proj", "o_proj", "up_gate_proj", "down_proj"):
                setattr(lw, n, torch.from_numpy(z[f"l{i}_{n}"]))
        return w


class OracleLlama:
    """The worker API of model.py on CPU.  `attn` selects how attention is evaluated:
       "ref"   - reference rounding order (paged_attention_ref_order; fp32-softmax prefill stand-in),
       "exact" - fp64 definitions,
       "fast"  - the same definitions as vectorised fp32 torch-eager (CPU baseline arm of bench.py)."""

    def __init__(self, cfg: dict, weights: OracleWeights, *, block_size=16, num_blocks=64, num_cpu_blocks=8,
                 max_seqs_in_block_table=64, max_blocks_per_seq=64, attn="ref", dtype=torch.float16):
        self.cfg = cfg
        self.w = weights
        self.dtype = dtype
        self.attn = attn
        self.H = cfg["hidden_size"]; self.nq = cfg["num_attention_heads"]
        self.nkv = cfg.get("num_key_value_heads", self.nq)
        self.D = self.H // self.nq
        self.F = cfg["intermediate_size"]; self.L = cfg["num_hidden_layers"]
        self.eps = cfg["rms_norm_eps"]
        self.block_size = block_size
        rs = cfg.get("rope_scaling", 1.0)
        rs = 1.0 if rs is None else rs
        self.cos, self.sin = K.rope_tables(self.D, cfg.get("rope_theta", 10000), cfg["max_position_embeddings"], rs, dtype)
        shape = (num_blocks, self.L, self.nkv, block_size, self.D)
        self.k_cache = torch.zeros(shape, dtype=dtype); self.v_cache = torch.zeros(shape, dtype=dtype)
        sshape = (num_cpu_blocks,) + shape[1:]
        self.k_swap = torch.zeros(sshape, dtype=dtype); self.v_swap = torch.zeros(sshape, dtype=dtype)
        self.gpu_block_manager = K.BlockManagerOracle(num_blocks, max_seqs_in_block_table, max_blocks_per_seq, block_size)
        self.cpu_block_manager = K.BlockManagerOracle(num_cpu_blocks, max_seqs_in_block_table, max_blocks_per_seq, block_size)
        self.last_logits = None

    # model.py:252-359
    @torch.inference_mode()
    def forward(self, input_ids_list, seq_ids_list, decoding_seq_lens_list, ignore_kvcache=False,
                prefill_prefix_lens_list=None):
        """`prefill_prefix_lens_list` (chunked prefill; not in the reference): prefill entry i is the
        chunk of its prompt that starts at position prefill_prefix_lens_list[i]; the earlier positions are already in
        the KV cache.  None = the reference's contract (every prefill entry is a whole prompt)."""
        num_prefill_seqs = len(input_ids_list) - len(decoding_seq_lens_list)
        flat = list(itertools.chain(*input_ids_list))
        prefill_lens = [len(s) for s in input_ids_list[:num_prefill_seqs]]
        chunked = prefill_prefix_lens_list is not None
        prefix = [int(x) for x in prefill_prefix_lens_list] if chunked else [0] * num_prefill_seqs
        assert len(prefix) == num_prefill_seqs and not (chunked and ignore_kvcache)
        seq_lengths = [p + n for p, n in zip(prefix, prefill_lens)] + list(decoding_seq_lens_list)
        B, T = len(input_ids_list), len(flat)
        Tp = T - (B - num_prefill_seqs)
        starts = list(np.cumsum([0] + prefill_lens[:-1])) if prefill_lens else []
        positions = [p0 + p for p0, n in zip(prefix, prefill_lens) for p in range(n)] + [l - 1 for l in decoding_seq_lens_list]
        if not ignore_kvcache:
            self.gpu_block_manager.allocate_blocks_for_seqs(seq_ids_list, seq_lengths)
        S, nsb = K.select_seq_block_size(self.nkv, list(decoding_seq_lens_list))
        pos = torch.tensor(positions, dtype=torch.long)
        cos, sin = self.cos[pos], self.sin[pos]
        bt = self.gpu_block_manager.block_table
        scale = self.D ** -0.5

        import time as _time
        t_start = _time.perf_counter()
        x = self.w.wte[torch.tensor(flat, dtype=torch.long)]            # pre_layer.py:19
        res = torch.zeros_like(x)                                       # model.py:237
        t_layers0 = _time.perf_counter()
        for li, lw in enumerate(self.w.layers):                         # transformer_layer.py:31-130
            x, res = K.fused_add_rmsnorm(x, res, lw.attn_norm, self.eps)
            q = F.linear(x, lw.q_proj).view(T, self.nq, self.D)
            k = F.linear(x, lw.k_proj).view(T, self.nkv, self.D)
            v = F.linear(x, lw.v_proj).view(T, self.nkv, self.D)
            q, k = K.rotary_embedding(q, k, cos, sin)
            if not ignore_kvcache:
                K.store_kvcache_inplace(k, v, self.k_cache, self.v_cache, bt, seq_ids_list, starts, prefill_lens,
                                        decoding_seq_lens_list, num_prefill_seqs, Tp, self.block_size, li,
                                        prefill_prefix_lens=prefix if chunked else None)
            o = torch.empty((T, self.H), dtype=self.dtype)
            if num_prefill_seqs > 0 and chunked:
                op = K.prefix_prefill_attention_exact(q[:Tp], self.k_cache, self.v_cache, bt, seq_ids_list[:num_prefill_seqs],
                                                      starts, prefill_lens, prefix, scale, self.block_size, li, self.dtype,
                                                      compute_dtype=torch.float64 if self.attn == "exact" else torch.float32)
                o[:Tp] = op.reshape(Tp, self.H)
            elif num_prefill_seqs > 0:
                if self.attn == "exact":
                    op = K.prefill_attention_exact(q[:Tp], k[:Tp], v[:Tp], starts, prefill_lens, scale, self.dtype)
                else:
                    op = K.prefill_attention_exact(q[:Tp], k[:Tp], v[:Tp], starts, prefill_lens, scale, self.dtype,
                                                   compute_dtype=torch.float32)
                o[:Tp] = op.reshape(Tp, self.H)
            if len(decoding_seq_lens_list) > 0:
                sids = seq_ids_list[num_prefill_seqs:]
                if self.attn == "exact":
                    od = K.paged_attention_exact(q[Tp:], self.k_cache, self.v_cache, bt, sids, decoding_seq_lens_list,
                                                 scale, self.block_size, li, self.dtype)
                elif self.attn == "fast":
                    od = K.paged_attention_fast(q[Tp:], self.k_cache, self.v_cache, bt, sids, decoding_seq_lens_list,
                                                scale, self.block_size, li, self.dtype)
                else:
                    od = K.paged_attention_ref_order(q[Tp:], self.k_cache, self.v_cache, bt, sids, decoding_seq_lens_list,
                                                     scale, self.block_size, li, S, nsb)
                o[Tp:] = od
            o = F.linear(o, lw.o_proj)
            o, res = K.fused_add_rmsnorm(o, res, lw.ffn_norm, self.eps)
            ug = F.linear(o, lw.up_gate_proj)
            ug = K.silu_and_mul(ug)
            x = F.linear(ug[:, : self.F], lw.down_proj)
        t_layers1 = _time.perf_counter()
        x = x + res                                                     # model.py:247
        last_idx = [s + n - 1 for s, n in zip(starts, prefill_lens)] + list(range(Tp, T))   # post_layer.py:24-31
        last = K.rmsnorm(x[torch.tensor(last_idx, dtype=torch.long)], self.w.final_norm, self.eps)
        logits = F.linear(last, self.w.lm_head)
        self.last_logits = logits
        toks = torch.argmax(logits, dim=1).tolist()
        t_end = _time.perf_counter()
        # wall-clock split of the last call (used by bench.py's CPU baseline extrapolation)
        self.last_times = {"layers": t_layers1 - t_layers0, "pre_post": (t_end - t_start) - (t_layers1 - t_layers0)}
        return toks

    # model.py:361-399
    def _swap(self, seq_ids_list, is_swap_in):
        src = self.cpu_block_manager if is_swap_in else self.gpu_block_manager
        dst = self.gpu_block_manager if is_swap_in else self.cpu_block_manager
        lens = src.num_seq_allocated_blocks[np.asarray(seq_ids_list)] * self.block_size
        s_ids = src.gather_allocated_blocks_and_free(seq_ids_list)
        d_ids = dst.allocate_blocks_for_seqs(seq_ids_list, lens)
        K.swap_blocks_inplace(s_ids, d_ids, is_swap_in, self.k_cache, self.v_cache, self.k_swap, self.v_swap)

    def swap_in_seqs(self, seq_ids_list): self._swap(seq_ids_list, True)
    def swap_out_seqs(self, seq_ids_list): self._swap(seq_ids_list, False)

    # model.py:401-408
    def free_seqs_resources(self, seq_ids_list):
        self.gpu_block_manager.free_blocks_for_seqs(seq_ids_list)
        self.cpu_block_manager.free_blocks_for_seqs(seq_ids_list)
