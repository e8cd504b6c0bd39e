"""HF `config.json` -> LlamaModelConfig.

The attribute names are the ones the reference's worker reads (`swiftllm/model_config.py:5-46`): num_layers, num_q_heads,
num_kv_heads, hidden_size, head_dim, vocab_size, max_position_embeddings, ffn_inter_dim, rotary_base / rope_theta,
rms_norm_eps, rope_scaling (1.0 when absent or null; a dict for Llama 3.2 style scaling) and get_kvslot_size()."""
import json
import os

import torch

# attribute <- key of the HF config (required)
_REQUIRED = {
    "num_layers": "num_hidden_layers",
    "num_q_heads": "num_attention_heads",
    "hidden_size": "hidden_size",
    "vocab_size": "vocab_size",
    "max_position_embeddings": "max_position_embeddings",
    "ffn_inter_dim": "intermediate_size",
    "rms_norm_eps": "rms_norm_eps",
}


class LlamaModelConfig:
    def __init__(self, model_config: dict):
        arch, act = model_config.get("model_type"), model_config.get("hidden_act")
        assert arch == "llama", f"only the llama architecture is served, got model_type={arch!r}"
        assert act == "silu", f"the FFN kernel is SiLU-gated (silu_and_mul), got hidden_act={act!r}"
        missing = [k for k in _REQUIRED.values() if k not in model_config]
        assert not missing, f"config.json lacks {missing}"
        for attr, key in _REQUIRED.items():
            setattr(self, attr, model_config[key])
        self.num_kv_heads = model_config.get("num_key_value_heads") or self.num_q_heads      # MHA checkpoints omit it
        assert self.hidden_size % self.num_q_heads == 0 and self.num_q_heads % self.num_kv_heads == 0
        self.head_dim = self.hidden_size // self.num_q_heads
        theta = model_config.get("rope_theta")
        self.rope_theta = 10000 if theta is None else theta
        self.rotary_base = model_config.get("rotary_base", 10000) if theta is None else theta  # older configs' name for it
        scaling = model_config.get("rope_scaling")
        self.rope_scaling = 1.0 if scaling is None else scaling
        self.raw = dict(model_config)

    def get_kvslot_size(self, dtype: torch.dtype = torch.float16) -> int:
        """Bytes of KV cache one token occupies: K and V, every layer, every kv head."""
        per_layer = 2 * self.num_kv_heads * self.head_dim
        return per_layer * self.num_layers * dtype.itemsize

    @staticmethod
    def load_from_model_path(model_path: str) -> "LlamaModelConfig":
        with open(os.path.join(model_path, "config.json"), "r", encoding="utf-8") as f:
            return LlamaModelConfig(json.load(f))


# Llama-3-8B / 70B shapes for synthetic-weight benchmarks
LLAMA3_8B = dict(model_type="llama", num_hidden_layers=32, num_attention_heads=32, num_key_value_heads=8,
                 hidden_size=4096, intermediate_size=14336, vocab_size=128256, max_position_embeddings=8192,
                 rope_theta=500000.0, rms_norm_eps=1e-5, hidden_act="silu", rope_scaling=None)
LLAMA3_70B = dict(model_type="llama", num_hidden_layers=80, num_attention_heads=64, num_key_value_heads=8,
                  hidden_size=8192, intermediate_size=28672, vocab_size=128256, max_position_embeddings=8192,
                  rope_theta=500000.0, rms_norm_eps=1e-5, hidden_act="silu", rope_scaling=None)
