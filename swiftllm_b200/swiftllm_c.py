"""Replacement of the reference's native pybind11 module `swiftllm_c` (csrc/src/entrypoints.cpp:5-7):
`swap_blocks(source_block_ids, target_block_ids, is_swap_in, k_cache, v_cache, k_swap, v_swap)`
(csrc/src/block_swapping.cpp:22-85).  The native side is `sllm_swap_blocks` in csrc/host.cu."""
import ctypes

import torch

from swiftllm_b200 import _lib


def swap_blocks(source_block_ids, target_block_ids, is_swap_in: bool,
                k_cache: torch.Tensor, v_cache: torch.Tensor, k_swap: torch.Tensor, v_swap: torch.Tensor):
    n = len(source_block_ids)
    assert n == len(target_block_ids)
    if n == 0:
        return
    _lib.require_device(k_cache)
    assert not k_swap.is_cuda and not v_swap.is_cuda
    assert k_cache.is_contiguous() and v_cache.is_contiguous() and k_swap.is_contiguous() and v_swap.is_contiguous()
    Arr = ctypes.c_int64 * n
    src = Arr(*[int(x) for x in source_block_ids])
    dst = Arr(*[int(x) for x in target_block_ids])
    block_bytes = k_cache.numel() * k_cache.element_size() // k_cache.shape[0]
    _lib.check(_lib.lib().sllm_swap_blocks(
        ctypes.cast(src, ctypes.c_void_p), ctypes.cast(dst, ctypes.c_void_p), n, 1 if is_swap_in else 0,
        k_cache.data_ptr(), v_cache.data_ptr(), k_swap.data_ptr(), v_swap.data_ptr(), block_bytes, _lib.stream()),
        "swap_blocks")


def swap_blocks_device(source_block_ids: torch.Tensor, target_block_ids: torch.Tensor, is_swap_in: bool,
                       k_cache: torch.Tensor, v_cache: torch.Tensor, k_swap: torch.Tensor, v_swap: torch.Tensor):
    """Same data movement as swap_blocks, but the id lists are DEVICE tensors (int64) and one kernel moves all blocks through the
    pinned, device-mapped swap space: no `.tolist()` syncs (the reference does two per swap, model.py:374-377).  Addition of this
    implementation; needs `pin_swap_space`."""
    n = source_block_ids.numel()
    assert n == target_block_ids.numel()
    if n == 0:
        return
    _lib.require_device(k_cache)
    assert source_block_ids.is_cuda and target_block_ids.is_cuda
    assert source_block_ids.dtype == torch.int64 and target_block_ids.dtype == torch.int64
    assert source_block_ids.is_contiguous() and target_block_ids.is_contiguous()
    assert not k_swap.is_cuda and not v_swap.is_cuda and k_swap.is_pinned() and v_swap.is_pinned(), \
        "swap_blocks_device needs a pinned swap space (EngineConfig.pin_swap_space)"
    assert k_cache.is_contiguous() and v_cache.is_contiguous() and k_swap.is_contiguous() and v_swap.is_contiguous()
    block_bytes = k_cache.numel() * k_cache.element_size() // k_cache.shape[0]
    _lib.check(_lib.lib().sllm_swap_blocks_gathered(
        source_block_ids.data_ptr(), target_block_ids.data_ptr(), n, 1 if is_swap_in else 0,
        k_cache.data_ptr(), v_cache.data_ptr(), k_swap.data_ptr(), v_swap.data_ptr(), block_bytes, _lib.stream()),
        "swap_blocks_gathered")
