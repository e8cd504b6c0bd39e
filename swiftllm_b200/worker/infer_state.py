"""Per-forward metadata handed to every kernel wrapper.  Same fields as the reference dataclass
(swiftllm/worker/infer_state.py:5-29); the trailing optional fields are additions of this implementation."""
import dataclasses
from typing import Optional

import torch


@dataclasses.dataclass
class LlamaInferState:
    batch_size: int
    num_tokens: int

    seq_ids: torch.Tensor   # [batch_size]
    softmax_scale: float    # Equal to 1/sqrt(head_dim)

    num_prefill_seqs: int
    num_prefill_tokens: int
    prefill_seq_start_locs: torch.Tensor  # [num_prefill_seqs]
    prefill_seq_start_locs_with_end: torch.Tensor  # [num_prefill_seqs+1]
    prefill_seq_lens: torch.Tensor  # [num_prefill_seqs]
    max_prefill_len: int

    num_decoding_seqs: int
    decoding_seq_lens: torch.Tensor  # [num_decoding_seqs]
    max_decoding_len: int

    seq_block_size: int
    num_seq_blocks: int

    position_cos: torch.Tensor  # [num_tokens, head_dim//2]
    position_sin: torch.Tensor  # [num_tokens, head_dim//2]

    ignore_kvcache: bool    # Skip storing the key/value cache, useful when profiling the number of kv blocks

    # ---- additions (not in the reference) ----
    # seq_block_size the library should use for paged attention: 0 = choose v1/v2 automatically.  The
    # reference's `seq_block_size` above is still computed with its heuristic (model.py:320-324) and can be
    # forced onto the kernel by setting this field to it.
    paged_attn_seq_block_size: int = 0
    last_token_indices: Optional[torch.Tensor] = None   # [batch_size] (post_layer.py:24-31), precomputed on the host
    # Chunked ("prefix-aware") prefill, SURVEY.md §8 f-1: prefill entry i holds the tokens at positions
    # [prefill_prefix_lens[i], prefill_prefix_lens[i] + prefill_seq_lens[i]) of its sequence; everything before is
    # already in the KV cache.  None = the reference's contract (whole prompts, attention over the packed k/v).
    prefill_prefix_lens: Optional[torch.Tensor] = None  # [num_prefill_seqs] int32
    max_prefill_kv_len: int = 0                          # max_i(prefix_i + chunk_i); 0 when not chunked
