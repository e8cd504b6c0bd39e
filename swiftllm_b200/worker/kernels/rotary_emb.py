"""Reference: swiftllm/worker/kernels/rotary_emb.py (rotary_embedding_inplace :44-58)."""
import torch

from swiftllm_b200 import _lib
from swiftllm_b200.worker.infer_state import LlamaInferState


def rotary_embedding_inplace(
    q: torch.Tensor,  # [num_tokens, num_q_heads, head_dim]
    k: torch.Tensor,  # [num_tokens, num_k_heads, head_dim]
    infer_state: LlamaInferState
):
    qs, ks = _lib.row_stride(q), _lib.row_stride(k)      # contiguous (reference) or slices of a fused QKV output
    cos, sin = infer_state.position_cos, infer_state.position_sin
    assert cos.is_contiguous() and sin.is_contiguous()
    assert cos.dtype == q.dtype and sin.dtype == q.dtype
    _lib.require_device(q)
    T, nq, D = q.shape
    nkv = k.shape[1]
    assert cos.shape == (T, D // 2) and sin.shape == (T, D // 2)
    _lib.check(_lib.lib().sllm_rotary_embedding_inplace(
        q.data_ptr(), k.data_ptr(), cos.data_ptr(), sin.data_ptr(), T, nq, nkv, D, qs, ks,
        _lib.dtype_tag(q.dtype), _lib.stream()), "rotary_embedding_inplace")
