"""TEST INFRASTRUCTURE: run the PRODUCT's host path (swiftllm_b200.LlamaModel: metadata staging, infer state, block manager
with its host mirror, layer sequencing, last-token gather, swap / free) on the CPU of the build container, with every CUDA
kernel wrapper replaced by the oracle's restatement of that kernel.  Nothing here ships; the product itself has no CPU path
(its wrappers fail loudly without the CUDA library).  What this buys: the Python side of the data plane is exercised by
`-m "not gpu"` tests against `oracle.model.OracleLlama`, so host-logic regressions are caught before a GPU is involved."""
from __future__ import annotations

import contextlib
import importlib

import torch
from torch.overrides import TorchFunctionMode

from oracle import kernels as K


def _is_cuda_dev(x):
    return (isinstance(x, torch.device) and x.type == "cuda") or (isinstance(x, str) and x.startswith("cuda"))


class _CudaToCpu(TorchFunctionMode):
    """device=cuda -> cpu (positional or keyword), pinned host memory -> ordinary memory."""

    def __torch_function__(self, func, types_, args=(), kwargs=None):
        kwargs = dict(kwargs or {})
        if getattr(func, "__name__", "") == "pin_memory":
            return args[0]
        if kwargs.get("pin_memory", False):
            kwargs["pin_memory"] = False
        to_cuda = False
        if "device" in kwargs and _is_cuda_dev(kwargs["device"]):
            kwargs["device"] = "cpu"; to_cuda = True
        if any(_is_cuda_dev(a) for a in args):
            args = tuple("cpu" if _is_cuda_dev(a) else a for a in args); to_cuda = True
        out = func(*args, **kwargs)
        # a host -> device transfer always yields a NEW tensor; `.to("cpu")` of a CPU tensor would alias its source
        if to_cuda and getattr(func, "__name__", "") in ("to", "cuda") and args and out is args[0]:
            out = out.clone()
        return out


# ---------------------------------------------------------------- oracle-backed stand-ins of the kernel wrappers
def _fused_add_rmsnorm_inplace(x, residual, weight, eps):
    xo, ro = K.fused_add_rmsnorm(x, residual, weight, eps)
    x.copy_(xo); residual.copy_(ro)


def _rmsnorm_inplace(x, weight, eps):
    x.copy_(K.rmsnorm(x, weight, eps))


def _rotary_embedding_inplace(q, k, infer_state):
    qo, ko = K.rotary_embedding(q, k, infer_state.position_cos, infer_state.position_sin)
    q.copy_(qo); k.copy_(ko)


def _silu_and_mul_inplace(x):
    F = x.shape[1] // 2
    x[:, :F] = K.silu_and_mul(x)[:, :F]


def _store_kvcache(k, v, k_cache, v_cache, block_table, model_config, engine_config, st, cur_layer):
    kw = {}
    if getattr(st, "prefill_prefix_lens", None) is not None:
        kw["prefill_prefix_lens"] = st.prefill_prefix_lens.tolist()
    K.store_kvcache_inplace(k, v, k_cache, v_cache, block_table.numpy(), st.seq_ids.tolist(), st.prefill_seq_start_locs.tolist(),
                            st.prefill_seq_lens.tolist(), st.decoding_seq_lens.tolist(), st.num_prefill_seqs, st.num_prefill_tokens,
                            k_cache.shape[3], cur_layer, **kw)


def _rotary_store_kvcache_decode(q, k, v, k_cache, v_cache, block_table, st, cur_layer):
    assert st.num_prefill_seqs == 0
    _rotary_embedding_inplace(q, k, st)
    K.store_kvcache_inplace(k, v, k_cache, v_cache, block_table.numpy(), st.seq_ids.tolist(), [], [], st.decoding_seq_lens.tolist(),
                            0, 0, k_cache.shape[3], cur_layer)


def _paged_attention(q, k_cache, v_cache, block_table, model_config, engine_config, st, cur_layer, o):
    out = K.paged_attention_exact(q, k_cache, v_cache, block_table.numpy(), st.seq_ids[st.num_prefill_seqs:].tolist(),
                                  st.decoding_seq_lens.tolist(), st.softmax_scale, k_cache.shape[3], cur_layer, q.dtype)
    o.copy_(out.reshape(o.shape))


def _prefill_attention(q, k, v, o, model_config, engine_config, st):
    out = K.prefill_attention_exact(q, k, v, st.prefill_seq_start_locs.tolist(), st.prefill_seq_lens.tolist(), st.softmax_scale, q.dtype)
    o.copy_(out.reshape(o.shape))


def _prefill_attention_paged(q, k_cache, v_cache, block_table, o, model_config, engine_config, st, cur_layer):
    out = K.prefix_prefill_attention_exact(q, k_cache, v_cache, block_table.numpy(), st.seq_ids[:st.num_prefill_seqs].tolist(),
                                           st.prefill_seq_start_locs.tolist(), st.prefill_seq_lens.tolist(),
                                           st.prefill_prefix_lens.tolist(), st.softmax_scale, k_cache.shape[3], cur_layer, q.dtype)
    o.copy_(out.reshape(o.shape))


def _allocate_blocks(nsab, block_table, is_free, seq_ids, target_lens, block_size, new_blocks, status):
    """block_manager.py:43-79 of the reference: the lowest free ids, ascending, handed out in batch order."""
    free_ids = torch.nonzero(is_free).flatten().tolist()
    p, out = 0, []
    for sid, tl in zip(seq_ids.tolist(), target_lens.tolist()):
        have = int(nsab[sid])
        need = (tl + block_size - 1) // block_size - have
        for i in range(need):
            b = free_ids[p]; p += 1
            block_table[sid, have + i] = b
            is_free[b] = False
            out.append(b)
        nsab[sid] = have + need
    if new_blocks is not None and out:
        new_blocks.copy_(torch.tensor(out, dtype=torch.int64))


def _unset(nsab, block_table, seq_ids, is_free):
    for sid in seq_ids.tolist():
        n = int(nsab[sid])
        is_free[block_table[sid, :n].long()] = True
        nsab[sid] = 0


def _gather_and_unset(nsab, block_table, seq_ids, is_free, total_blocks=None):
    ids = []
    for sid in seq_ids.tolist():
        ids += block_table[sid, : int(nsab[sid])].tolist()
    _unset(nsab, block_table, seq_ids, is_free)
    return torch.tensor(ids, dtype=torch.int32)


def _swap_blocks(src_ids, dst_ids, is_swap_in, k_cache, v_cache, k_swap, v_swap):
    K.swap_blocks_inplace(src_ids, dst_ids, is_swap_in, k_cache, v_cache, k_swap, v_swap)


class GlooFusedAllReduce:
    """CPU stand-in of swiftllm_b200.worker.tp_comm.FusedAllReduce with the SAME observable semantics as the CUDA kernels
    (csrc/allreduce_norm.cu), exchanged over gloo: partials are summed in rank order in fp32, residual <- h(h(sum) + residual),
    output = rmsnorm(residual) * weight.  two_shot: only the rows a rank owns (t % world == rank) are reduced / added /
    normalised by it; its other residual rows are POISONED with NaN (the kernel simply does not maintain them), and the
    normalised rows are gathered from their owners."""

    def __init__(self, max_tokens, hidden, dtype, device, group=None, two_shot=False, nvls=False, ll=False):
        import torch.distributed as dist
        self.group = group if group is not None else dist.group.WORLD
        self.rank, self.world = dist.get_rank(self.group), dist.get_world_size(self.group)
        self.two_shot, self.nvls, self.hidden, self.max_tokens = bool(two_shot) or bool(nvls) or bool(ll), bool(nvls), hidden, max_tokens
        self.max_fused_tokens = max_tokens
        self.ll = bool(ll)        # the LL kernel has the two-shot kernel's observable semantics (row owners, sharded residual)
        self.data = torch.zeros((2, max_tokens, hidden), dtype=dtype)

    def partial_out(self, slot, num_tokens):
        return self.data[slot, :num_tokens]

    def reduce_add_norm(self, slot, num_tokens, residual, weight, eps):
        import torch.distributed as dist
        mine = self.data[slot, :num_tokens].contiguous()
        parts = [torch.empty_like(mine) for _ in range(self.world)]
        dist.all_gather(parts, mine, group=self.group)
        acc = torch.zeros(mine.shape, dtype=torch.float32)
        for prt in parts:                                  # fixed rank order, fp32 accumulation
            acc += prt.float()
        s = acc.to(mine.dtype) + residual                  # h(h(sum) + r)
        if not self.two_shot:
            residual.copy_(s)
            return K.rmsnorm(s, weight, eps)
        own = (torch.arange(num_tokens) % self.world) == self.rank
        residual[own] = s[own]
        residual[~own] = float("nan")
        out = torch.zeros_like(s)
        if bool(own.any()):
            out[own] = K.rmsnorm(residual[own], weight, eps)
        dist.all_reduce(out, group=self.group)             # every row has exactly one non-zero contributor: exact
        return out


def _swap_blocks_device(src_ids, dst_ids, is_swap_in, k_cache, v_cache, k_swap, v_swap):
    assert src_ids.dtype == torch.int64 and dst_ids.dtype == torch.int64
    K.swap_blocks_inplace(src_ids.tolist(), dst_ids.tolist(), is_swap_in, k_cache, v_cache, k_swap, v_swap)


_PATCHES = [
    ("swiftllm_b200.worker.layers.transformer_layer", "fused_add_rmsnorm_inplace", _fused_add_rmsnorm_inplace),
    ("swiftllm_b200.worker.layers.transformer_layer", "rotary_embedding_inplace", _rotary_embedding_inplace),
    ("swiftllm_b200.worker.layers.transformer_layer", "silu_and_mul_inplace", _silu_and_mul_inplace),
    ("swiftllm_b200.worker.layers.transformer_layer", "store_kvcache", _store_kvcache),
    ("swiftllm_b200.worker.layers.transformer_layer", "paged_attention", _paged_attention),
    ("swiftllm_b200.worker.layers.transformer_layer", "rotary_store_kvcache_decode", _rotary_store_kvcache_decode),
    ("swiftllm_b200.worker.layers.transformer_layer", "prefill_attention", _prefill_attention),
    ("swiftllm_b200.worker.layers.transformer_layer", "prefill_attention_paged", _prefill_attention_paged),
    ("swiftllm_b200.worker.layers.post_layer", "rmsnorm_inplace", _rmsnorm_inplace),
    ("swiftllm_b200.worker.block_manager", "_allocate_kernel", _allocate_blocks),
    ("swiftllm_b200.worker.block_manager", "unset_block_table_and_num_seq_alloc_blocks", _unset),
    ("swiftllm_b200.worker.block_manager", "gather_allocated_blocks_and_unset", _gather_and_unset),
    ("swiftllm_b200.swiftllm_c", "swap_blocks", _swap_blocks),
    ("swiftllm_b200.swiftllm_c", "swap_blocks_device", _swap_blocks_device),
]


@contextlib.contextmanager
def product_on_cpu(extra_patches=()):
    """Inside this context `swiftllm_b200.LlamaModel` runs on CPU tensors with oracle-backed kernels."""
    from oracle.ref_shim import cuda_as_cpu        # fake CUDA streams / events (shared with the reference's CPU shim)
    saved = []
    cur_dev = torch.cuda.current_device
    torch.cuda.current_device = lambda: 0
    try:
        for mod_name, attr, fn in list(_PATCHES) + list(extra_patches):
            mod = importlib.import_module(mod_name)
            saved.append((mod, attr, getattr(mod, attr)))
            setattr(mod, attr, fn)
        with cuda_as_cpu(), _CudaToCpu():
            yield
    finally:
        for mod, attr, old in saved:
            setattr(mod, attr, old)
        torch.cuda.current_device = cur_dev
