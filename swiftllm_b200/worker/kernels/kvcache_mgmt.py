"""Reference: swiftllm/worker/kernels/kvcache_mgmt.py (store_kvcache :81-122)."""
import torch

from swiftllm_b200 import _lib
from swiftllm_b200.worker.infer_state import LlamaInferState


def store_kvcache(
    k: torch.Tensor,
    v: torch.Tensor,
    k_cache: torch.Tensor,
    v_cache: torch.Tensor,
    block_table: torch.Tensor,
    model_config,
    engine_config,
    infer_state: LlamaInferState,
    cur_layer: int
):
    ks, vs = _lib.row_stride(k), _lib.row_stride(v)      # contiguous (reference) or slices of a fused QKV output
    assert k_cache.is_contiguous()
    assert v_cache.is_contiguous()
    assert block_table.is_contiguous()
    assert infer_state.seq_ids.is_contiguous()
    assert infer_state.decoding_seq_lens.is_contiguous()
    assert block_table.dtype == torch.int32 and infer_state.seq_ids.dtype == torch.int32
    _lib.require_device(k_cache)
    num_layers, nkv, bs, D = k_cache.shape[1:]
    prefix = getattr(infer_state, "prefill_prefix_lens", None)
    if prefix is not None:
        # chunked prefill (not in the reference): chunk token t goes to position prefix + t
        assert prefix.dtype == torch.int32 and prefix.is_contiguous() and prefix.shape[0] == infer_state.num_prefill_seqs
        _lib.check(_lib.lib().sllm_store_kvcache_chunked(
            k.data_ptr(), v.data_ptr(), k_cache.data_ptr(), v_cache.data_ptr(), block_table.data_ptr(),
            infer_state.seq_ids.data_ptr(),
            _lib.ptr(infer_state.prefill_seq_start_locs), _lib.ptr(infer_state.prefill_seq_lens), prefix.data_ptr(),
            _lib.ptr(infer_state.decoding_seq_lens),
            infer_state.num_prefill_seqs, infer_state.num_decoding_seqs, infer_state.num_prefill_tokens,
            infer_state.max_prefill_len, cur_layer, num_layers, nkv, bs, D, block_table.shape[1], ks, vs,
            _lib.dtype_tag(k.dtype), _lib.stream()), "store_kvcache_chunked")
        return
    _lib.check(_lib.lib().sllm_store_kvcache(
        k.data_ptr(), v.data_ptr(), k_cache.data_ptr(), v_cache.data_ptr(), block_table.data_ptr(),
        infer_state.seq_ids.data_ptr(),
        _lib.ptr(infer_state.prefill_seq_start_locs), _lib.ptr(infer_state.prefill_seq_lens),
        _lib.ptr(infer_state.decoding_seq_lens),
        infer_state.num_prefill_seqs, infer_state.num_decoding_seqs, infer_state.num_prefill_tokens,
        infer_state.max_prefill_len, cur_layer, num_layers, nkv, bs, D, block_table.shape[1], ks, vs,
        _lib.dtype_tag(k.dtype), _lib.stream()), "store_kvcache")


def rotary_store_kvcache_decode(
    q: torch.Tensor,            # [num_decoding_seqs, num_q_heads, head_dim], rotated in place
    k: torch.Tensor,            # [num_decoding_seqs, num_kv_heads, head_dim], rotated in place
    v: torch.Tensor,            # [num_decoding_seqs, num_kv_heads, head_dim]
    k_cache: torch.Tensor,
    v_cache: torch.Tensor,
    block_table: torch.Tensor,
    infer_state: LlamaInferState,
    cur_layer: int,
):
    """rotary_embedding_inplace (rotary_emb.py:44-58) + the decode part of store_kvcache (kvcache_mgmt.py:50-79) in one launch,
    for batches that hold decoding rows only (two launches per layer in the reference).  Bit-identical results."""
    assert infer_state.num_prefill_seqs == 0, "the fused rotary + store launch serves pure-decode batches"
    qs, ks, vs = _lib.row_stride(q), _lib.row_stride(k), _lib.row_stride(v)
    cos, sin = infer_state.position_cos, infer_state.position_sin
    assert cos.is_contiguous() and sin.is_contiguous() and cos.dtype == q.dtype and sin.dtype == q.dtype
    assert k_cache.is_contiguous() and v_cache.is_contiguous() and block_table.is_contiguous()
    assert block_table.dtype == torch.int32 and infer_state.seq_ids.dtype == torch.int32
    _lib.require_device(q)
    Bd, nq, D = q.shape
    assert cos.shape == (Bd, D // 2) and sin.shape == (Bd, D // 2)
    num_layers, nkv, bs, _ = k_cache.shape[1:]
    _lib.check(_lib.lib().sllm_rotary_store_kvcache_decode(
        q.data_ptr(), k.data_ptr(), v.data_ptr(), cos.data_ptr(), sin.data_ptr(), k_cache.data_ptr(), v_cache.data_ptr(),
        block_table.data_ptr(), infer_state.seq_ids.data_ptr(), infer_state.decoding_seq_lens.data_ptr(), Bd, cur_layer,
        num_layers, nq, nkv, bs, D, block_table.shape[1], qs, ks, vs, _lib.dtype_tag(q.dtype), _lib.stream()),
        "rotary_store_kvcache_decode")
