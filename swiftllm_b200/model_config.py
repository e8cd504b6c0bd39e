"""HF config.json -> LlamaModelConfig.  Same attribute names as the reference (swiftllm/model_config.py:5-46)."""
import json
import os

import torch


class LlamaModelConfig:
    def __init__(self, model_config: dict):
        assert model_config["model_type"] == "llama"
        self.num_layers = model_config["num_hidden_layers"]
        self.num_q_heads = model_config["num_attention_heads"]
        self.num_kv_heads = model_config.get("num_key_value_heads", self.num_q_heads)
        self.hidden_size = model_config["hidden_size"]
        self.head_dim = self.hidden_size // self.num_q_heads
        self.vocab_size = model_config["vocab_size"]
        self.max_position_embeddings = model_config["max_position_embeddings"]
        self.ffn_inter_dim = model_config["intermediate_size"]
        self.rotary_base = model_config.get("rope_theta", model_config.get("rotary_base", 10000))
        self.rms_norm_eps = model_config["rms_norm_eps"]
        self.rope_scaling = model_config.get("rope_scaling", 1.0)
        self.rope_theta = model_config.get("rope_theta", 10000)
        if self.rope_scaling is None:
            self.rope_scaling = 1.0
        assert model_config["hidden_act"] == "silu"
        self.raw = dict(model_config)

    def get_kvslot_size(self, dtype: torch.dtype = torch.float16) -> int:
        """Bytes of the K+V cache of one token over all layers and kv heads."""
        return (2 * self.num_layers * self.num_kv_heads * self.head_dim) * dtype.itemsize

    @staticmethod
    def load_from_model_path(model_path: str) -> "LlamaModelConfig":
        with open(os.path.join(model_path, "config.json"), "r", encoding="utf-8") as f:
            return LlamaModelConfig(json.loads(f.read()))


# Llama-3-8B / 70B shapes for synthetic-weight benchmarks
LLAMA3_8B = dict(model_type="llama", num_hidden_layers=32, num_attention_heads=32, num_key_value_heads=8,
                 hidden_size=4096, intermediate_size=14336, vocab_size=128256, max_position_embeddings=8192,
                 rope_theta=500000.0, rms_norm_eps=1e-5, hidden_act="silu", rope_scaling=None)
LLAMA3_70B = dict(model_type="llama", num_hidden_layers=80, num_attention_heads=64, num_key_value_heads=8,
                  hidden_size=8192, intermediate_size=28672, vocab_size=128256, max_position_embeddings=8192,
                  rope_theta=500000.0, rms_norm_eps=1e-5, hidden_act="silu", rope_scaling=None)
