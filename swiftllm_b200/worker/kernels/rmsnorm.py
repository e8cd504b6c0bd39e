"""Reference: swiftllm/worker/kernels/rmsnorm.py (rmsnorm_inplace :26-37, fused_add_rmsnorm_inplace :67-89)."""
import torch

from swiftllm_b200 import _lib


def rmsnorm_inplace(
    input_and_output: torch.Tensor,  # [num_tokens, hidden_size]
    weight: torch.Tensor,
    eps: float
):
    assert input_and_output.is_contiguous()
    assert weight.is_contiguous()
    _lib.require_device(input_and_output)
    T, H = input_and_output.shape
    _lib.check(_lib.lib().sllm_rmsnorm_inplace(
        input_and_output.data_ptr(), weight.data_ptr(), eps, T, H,
        _lib.dtype_tag(input_and_output.dtype), _lib.stream()), "rmsnorm_inplace")


def fused_add_rmsnorm_inplace(
    input_and_output: torch.Tensor,  # [num_tokens, hidden_size]
    residual_io: torch.Tensor,
    weight: torch.Tensor,
    eps: float
):
    """r = x + r (rounded to the storage dtype), x = rms_norm(r, w)."""
    assert input_and_output.is_contiguous()
    assert residual_io.is_contiguous()
    assert weight.is_contiguous()
    _lib.require_device(input_and_output)
    T, H = input_and_output.shape
    _lib.check(_lib.lib().sllm_fused_add_rmsnorm_inplace(
        input_and_output.data_ptr(), residual_io.data_ptr(), weight.data_ptr(), eps, T, H,
        _lib.dtype_tag(input_and_output.dtype), _lib.stream()), "fused_add_rmsnorm_inplace")
