"""CPU: host-side logic of the data plane against the oracle."""
import numpy as np
import torch

from oracle import kernels as K
from swiftllm_b200.model_config import LlamaModelConfig, LLAMA3_8B
from swiftllm_b200.worker.model import build_rope_tables, select_seq_block_size
from swiftllm_b200.worker.weight import tp_slice


def test_seq_block_size_matches_reference_heuristic():
    rng = np.random.default_rng(0)
    cases = [[4096] * 256, [11], [131072], [1], [65, 64, 63], [2048] * 7]
    cases += [rng.integers(1, 9000, size=int(rng.integers(1, 300))).tolist() for _ in range(50)]
    for lens in cases:
        for nkv in (1, 2, 8):
            assert select_seq_block_size(nkv, lens, max(lens)) == K.select_seq_block_size(nkv, lens)[0]


def test_rope_tables_match_oracle():
    for cfg in (dict(LLAMA3_8B), dict(LLAMA3_8B, rope_scaling={"factor": 8.0, "low_freq_factor": 1.0, "high_freq_factor": 4.0,
                                                                "original_max_position_embeddings": 1024, "rope_type": "llama3"}),
                dict(LLAMA3_8B, rope_scaling=2.0, max_position_embeddings=512)):
        mc = LlamaModelConfig(cfg)
        for dt in (torch.float16, torch.bfloat16):
            cos, sin = build_rope_tables(mc, dt)
            ocos, osin = K.rope_tables(mc.head_dim, mc.rope_theta, mc.max_position_embeddings, mc.rope_scaling, dt)
            assert torch.equal(cos, ocos) and torch.equal(sin, osin)


def test_tp_slices_partition_the_weights():
    w = torch.arange(8 * 6).reshape(8, 6)
    rows = [tp_slice(w, 0, r, 4) for r in range(4)]
    cols = [tp_slice(w, 1, r, 2) for r in range(2)]
    assert torch.equal(torch.cat(rows, 0), w) and torch.equal(torch.cat(cols, 1), w)
    assert tp_slice(w, None, 1, 2) is w


def test_llama3_8b_config_shapes():
    mc = LlamaModelConfig(LLAMA3_8B)
    assert (mc.num_layers, mc.num_q_heads, mc.num_kv_heads, mc.head_dim, mc.ffn_inter_dim) == (32, 32, 8, 128, 14336)
    assert mc.get_kvslot_size(torch.bfloat16) == 128 * 1024        # 128 KiB of KV per token
