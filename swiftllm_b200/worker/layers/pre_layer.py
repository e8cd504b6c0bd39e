"""Reference: swiftllm/worker/layers/pre_layer.py:6-20 (token embedding lookup)."""
import torch


class LlamaPreLayer:
    def __init__(self, model_config, weights):
        self.model_config = model_config
        self.weights = weights

    def forward(self, input_ids: torch.Tensor) -> torch.Tensor:
        return torch.embedding(self.weights.wte, input_ids, padding_idx=-1)
