"""Reference: swiftllm/worker/layers/post_layer.py:9-40 (last-token gather, final RMSNorm, lm_head, greedy argmax).

Addition (tensor parallelism): with a vocabulary-sharded lm_head every rank computes the logits of its
V / tp_size rows only and the greedy token is found with ONE all-gather of per-rank (max logit, global index) pairs
(8 bytes per sequence and rank) instead of every rank streaming the whole lm_head."""
import torch
import torch.distributed as dist

from swiftllm_b200.worker.infer_state import LlamaInferState
from swiftllm_b200.worker.kernels.linear import linear
from swiftllm_b200.worker.kernels.rmsnorm import rmsnorm_inplace


def merge_sharded_argmax(pairs: torch.Tensor) -> torch.Tensor:
    """pairs fp32 [ranks, batch, 2] = (max logit of the rank's vocabulary shard, its GLOBAL token index) -> int64 [batch]:
    the index of the largest logit; among equal maxima the smallest index (first occurrence, the convention of
    torch.argmax over the unsharded logits, post_layer.py:39).  Indices < 2^24 are exact in fp32."""
    vals, idx = pairs[..., 0], pairs[..., 1]
    best = vals.max(dim=0).values
    cand = torch.where(vals == best, idx, torch.full_like(idx, float("inf")))
    return cand.min(dim=0).values.to(torch.int64)


class LlamaPostLayer:
    def __init__(self, model_config, weights, tp_group=None, tp_rank: int = 0, tp_size: int = 1):
        self.model_config = model_config
        self.weights = weights
        self.tp_group, self.tp_rank, self.tp_size = tp_group, tp_rank, tp_size
        self.sharded = bool(getattr(weights, "lm_head_sharded", False))
        assert model_config.vocab_size < (1 << 24) or not self.sharded
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
        logits = linear(last_input, self.weights.lm_head)      # [batch_size, vocab_size] (or this rank's V / tp_size columns)
        if not self.sharded:
            if self.keep_logits:
                self.last_logits = logits
            return torch.argmax(logits, dim=1)

        B, Vs = logits.shape
        local = torch.argmax(logits, dim=1)
        mine = torch.stack((logits.gather(1, local[:, None]).squeeze(1).float(),
                            (local + self.tp_rank * Vs).float()), dim=1)             # [B, 2]
        # outputs in the concatenated form ([ranks * B, ...]): accepted by both NCCL and gloo
        pairs = torch.empty((self.tp_size * B, 2), dtype=torch.float32, device=logits.device)
        dist.all_gather_into_tensor(pairs, mine.contiguous(), group=self.tp_group)
        if self.keep_logits:                                   # parity tests only: reassemble the full logits
            full = torch.empty((self.tp_size * B, Vs), dtype=logits.dtype, device=logits.device)
            dist.all_gather_into_tensor(full, logits.contiguous(), group=self.tp_group)
            self.last_logits = full.view(self.tp_size, B, Vs).permute(1, 0, 2).reshape(B, self.tp_size * Vs)
        return merge_sharded_argmax(pairs.view(self.tp_size, B, 2))
