"""CPU dry run of GPU tests whose kernels were written without GPU access (tests/test_chunked_prefill_gpu.py, test_decode_fusion_gpu.py,
test_swap_device_gpu.py): the test FUNCTIONS are executed here with every kernel wrapper replaced by the oracle's restatement
(tests/cpu_shim.py) and "cuda" mapped to the CPU.  This cannot say anything about the kernels; it makes sure the tests
themselves - shapes, metadata, oracle calls, assertions, the model-level schedules - are sound, so that the first GPU minutes
spent on them test the kernels and not the tests."""
import importlib

import pytest
import torch

import cpu_shim
from cpu_shim import product_on_cpu

_EXTRA = [
    ("swiftllm_b200.worker.kernels.kvcache_mgmt", "store_kvcache", cpu_shim._store_kvcache),
    ("swiftllm_b200.worker.kernels.kvcache_mgmt", "rotary_store_kvcache_decode", cpu_shim._rotary_store_kvcache_decode),
    ("swiftllm_b200.worker.kernels.rotary_emb", "rotary_embedding_inplace", cpu_shim._rotary_embedding_inplace),
    ("swiftllm_b200.worker.kernels.prefill_attn", "prefill_attention", cpu_shim._prefill_attention),
    ("swiftllm_b200.worker.kernels.prefill_attn", "prefill_attention_paged", cpu_shim._prefill_attention_paged),
    ("swiftllm_b200.worker.kernels.paged_attn", "paged_attention", cpu_shim._paged_attention),
]


def _module(name):
    return importlib.import_module(name)


def test_dry_run_chunked_prefill_gpu_tests(golden):
    m = _module("test_chunked_prefill_gpu")
    with product_on_cpu(_EXTRA):
        m.test_store_kvcache_chunked_bit_exact(torch.float16, 16)
        m.test_store_kvcache_chunked_bit_exact(torch.bfloat16, 32)
        m.test_prefill_attention_paged_vs_exact_oracle(torch.float16, dict(nq=4, nkv=2, D=64, bs=16), "gen1-mma.sync")
        m.test_prefill_attention_paged_vs_exact_oracle(torch.bfloat16, dict(nq=8, nkv=2, D=128, bs=32), "gen1-mma.sync")
        m.test_prefill_attention_paged_nan_in_unowned_pages_is_harmless(torch.float16, "gen2-tcgen05")
        m.test_prefill_attention_paged_zero_prefix_equals_packed_kernel(torch.bfloat16, "gen2-tcgen05")
        m.test_model_chunked_prefill_matches_whole_prompt_and_oracle("float16", "gen2-tcgen05")
        m.test_model_chunked_replay_of_the_reference_golden_trace(golden)


@pytest.mark.slow_cpu
def test_dry_run_sarathi_shape_test():
    m = _module("test_chunked_prefill_gpu")
    with product_on_cpu(_EXTRA):
        m.test_prefill_attention_paged_sarathi_shape_property("gen2-tcgen05")


def test_dry_run_decode_fusion_and_swap_gpu_tests():
    f, s = _module("test_decode_fusion_gpu"), _module("test_swap_device_gpu")
    with product_on_cpu(_EXTRA):
        f.test_rotary_store_decode_equals_separate_kernels_and_oracle(torch.float16, dict(nq=4, nkv=2, D=64, bs=16))
        f.test_rotary_store_decode_equals_separate_kernels_and_oracle(torch.bfloat16, dict(nq=8, nkv=1, D=128, bs=32))
        for variant in (dict(device_swap=True), dict(swap_on_copy_stream=True), dict(device_swap=True, swap_on_copy_stream=True)):
            s.test_model_swap_with_device_ids_equals_memcpy_path(variant)
    # (test_swap_blocks_device_ids_exact and the CUDA-graph model test need real pinned memory / graphs: GPU only)
