"""Reference: swiftllm/worker/kernels/prefill_attn.py (prefill_attention :102-139); also replaces the
vllm_flash_attn.flash_attn_varlen_func call at swiftllm/worker/layers/transformer_layer.py:86-96."""
import torch

from swiftllm_b200 import _lib
from swiftllm_b200.worker.infer_state import LlamaInferState


def prefill_attention(
    q: torch.Tensor,    # [num_prefill_tokens, num_q_heads, head_dim]
    k: torch.Tensor,    # [num_prefill_tokens, num_kv_heads, head_dim]
    v: torch.Tensor,    # [num_prefill_tokens, num_kv_heads, head_dim]
    o: torch.Tensor,    # [num_prefill_tokens, num_q_heads, head_dim]
    model_config,
    engine_config,
    infer_state: LlamaInferState,
):
    assert o.is_contiguous()
    qs, ks, vs = _lib.row_stride(q), _lib.row_stride(k), _lib.row_stride(v)
    _lib.require_device(q)
    Tp, nq, D = q.shape
    nkv = k.shape[1]
    if infer_state.num_prefill_seqs == 0:
        return
    _lib.check(_lib.lib().sllm_prefill_attention(
        q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr(),
        infer_state.prefill_seq_start_locs.data_ptr(), infer_state.prefill_seq_lens.data_ptr(),
        infer_state.softmax_scale, infer_state.num_prefill_seqs, infer_state.max_prefill_len, Tp, nq, nkv, D, qs, ks, vs,
        _lib.dtype_tag(q.dtype), _lib.stream()), "prefill_attention")
