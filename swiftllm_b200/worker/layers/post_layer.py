"""Reference: swiftllm/worker/layers/post_layer.py:9-40 (last-token gather, final RMSNorm, lm_head, greedy argmax)."""
import torch

from swiftllm_b200.worker.infer_state import LlamaInferState
from swiftllm_b200.worker.kernels.linear import linear
from swiftllm_b200.worker.kernels.rmsnorm import rmsnorm_inplace


class LlamaPostLayer:
    def __init__(self, model_config, weights):
        self.model_config = model_config
        self.weights = weights
        self.last_logits = None     # kept for parity tests (the reference discards them)
        self.keep_logits = False

    def forward(self, input_embds: torch.Tensor, infer_state: LlamaInferState, already_normed: bool = False) -> torch.Tensor:
        """`already_normed`: input_embds rows are rmsnorm(final residual) already (TP fused-exchange path; RMSNorm is
        row-wise, so normalising before or after the last-token gather is the same computation)."""
        idx = infer_state.last_token_indices
        if idx is None:
            idx = torch.cat((
                infer_state.prefill_seq_start_locs + infer_state.prefill_seq_lens - 1,
                torch.arange(infer_state.num_prefill_tokens, infer_state.num_tokens, device=input_embds.device, dtype=torch.int32)
            ), dim=0)
        last_input = input_embds.index_select(0, idx)          # [batch_size, hidden_size], fresh contiguous buffer
        if not already_normed:
            rmsnorm_inplace(last_input, self.weights.final_norm, self.model_config.rms_norm_eps)
        logits = linear(last_input, self.weights.lm_head)      # [batch_size, vocab_size]
        if self.keep_logits:
            self.last_logits = logits
        return torch.argmax(logits, dim=1)
