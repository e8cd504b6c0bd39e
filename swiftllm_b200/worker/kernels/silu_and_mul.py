"""Reference: swiftllm/worker/kernels/silu_and_mul.py (silu_and_mul_inplace :25-34)."""
import torch

from swiftllm_b200 import _lib


def silu_and_mul_inplace(
    x: torch.Tensor  # [num_tokens, 2*ffn_inter_dim] = [up | gate]; result in x[:, :ffn_inter_dim]
):
    assert x.is_contiguous()
    _lib.require_device(x)
    num_tokens = x.shape[0]
    ffn_inter_dim = x.shape[1] // 2
    _lib.check(_lib.lib().sllm_silu_and_mul_inplace(
        x.data_ptr(), num_tokens, ffn_inter_dim, _lib.dtype_tag(x.dtype), _lib.stream()), "silu_and_mul_inplace")
