"""Per-forward metadata handed to every kernel wrapper.

The field NAMES are the interface the reference's wrappers read (`swiftllm/worker/infer_state.py:5-29`: anything that is
called with an `infer_state` looks these attributes up), so they are kept; grouping, types and the trailing optional fields
are this implementation's.  All tensors are int32 (or the model dtype for cos/sin) on the model's device; instances are always
built with keyword arguments (worker/model.py)."""
import dataclasses
from typing import Optional

import torch


@dataclasses.dataclass
class LlamaInferState:
    # ---- the packed batch: prefill sequences first, then one row per decoding sequence
    batch_size: int                                 # sequences in the batch
    num_tokens: int                                 # rows of the activation matrix
    seq_ids: torch.Tensor                           # [batch_size] block-table row of every sequence
    num_prefill_seqs: int
    num_decoding_seqs: int
    num_prefill_tokens: int                         # rows [0, num_prefill_tokens) are prompt tokens

    # ---- prefill part
    prefill_seq_lens: torch.Tensor                  # [num_prefill_seqs] tokens of each prompt (chunk) in this batch
    prefill_seq_start_locs: torch.Tensor            # [num_prefill_seqs] first row of each prompt
    prefill_seq_start_locs_with_end: torch.Tensor   # [num_prefill_seqs + 1] the same, closed with num_prefill_tokens
    max_prefill_len: int

    # ---- decode part
    decoding_seq_lens: torch.Tensor                 # [num_decoding_seqs] sequence lengths INCLUDING the new token
    max_decoding_len: int
    seq_block_size: int                             # the reference's flash-decoding split (model.py:305-324) ...
    num_seq_blocks: int                             # ... and the number of splits of the longest sequence

    # ---- attention / rotary constants of the step
    softmax_scale: float                            # head_dim ** -0.5
    position_cos: torch.Tensor                      # [num_tokens, head_dim / 2] rows of the RoPE tables at the tokens' positions
    position_sin: torch.Tensor

    ignore_kvcache: bool                            # memory-profiling forward: no KV store, no decode attention

    # ---- additions of this implementation (defaults keep the reference's behaviour)
    # Split size the library should use for paged attention; 0 = choose v1 / v2 itself.  `seq_block_size` above can be forced
    # onto the kernel by copying it here.
    paged_attn_seq_block_size: int = 0
    last_token_indices: Optional[torch.Tensor] = None   # [batch_size] rows to sample from (post_layer.py:24-31), host-computed
    # Chunked ("prefix-aware") prefill, SURVEY.md §8 f-1: prefill entry i holds the tokens at positions
    # [prefill_prefix_lens[i], prefill_prefix_lens[i] + prefill_seq_lens[i]) of its sequence; everything before is already in
    # the KV cache.  None = whole prompts, attention over the packed k/v.
    prefill_prefix_lens: Optional[torch.Tensor] = None  # [num_prefill_seqs]
    max_prefill_kv_len: int = 0                          # max_i(prefix_i + chunk_i); 0 when not chunked
