"""SARATHI-style piggybacking on top of `LlamaModel.forward(..., prefill_prefix_lens_list=...)`.

The reference's scheduler launches whole-prompt prefill batches and marks where piggybacking would go
(swiftllm/server/scheduler.py:93-94: "If you want decoding requests to be piggybacked, you can do it here"), but its worker
cannot run a partial prompt.  This module is the worker-side half of that hook: a pure function that turns "prompts in
progress + running decodes + a token budget" into the arguments of ONE forward call, and the bookkeeping to apply to its
result.  It holds no state and knows nothing about queues, preemption or request objects - that stays in the control plane.
"""
from __future__ import annotations

import dataclasses
from typing import Optional, Sequence


@dataclasses.dataclass
class PrefillProgress:
    seq_id: int
    prompt: Sequence[int]
    done: int = 0                      # prompt tokens already in the KV cache

    @property
    def remaining(self) -> int:
        return len(self.prompt) - self.done


@dataclasses.dataclass
class PiggybackStep:
    input_ids_list: list               # prefill chunks first, then one token per decoding sequence (forward()'s layout)
    seq_ids_list: list
    decoding_seq_lens_list: list
    prefill_prefix_lens_list: Optional[list]      # None when the step carries no prefill work
    chunks: list                       # (index into `prefilling`, chunk length, is the last chunk of its prompt)


def plan_piggyback_step(prefilling: Sequence[PrefillProgress], decoding_seq_ids: Sequence[int],
                        decoding_last_tokens: Sequence[int], decoding_seq_lens: Sequence[int],
                        max_tokens_in_step: int, max_chunk: Optional[int] = None) -> PiggybackStep:
    """Every running decode takes one token of the budget; what is left goes to the prompts in progress, first come first
    served, each getting at most `max_chunk` tokens (default: whatever fits).  decoding_seq_lens[i] includes the new token,
    as in forward()."""
    assert len(decoding_seq_ids) == len(decoding_last_tokens) == len(decoding_seq_lens)
    budget = max_tokens_in_step - len(decoding_seq_ids)
    ids, sids, prefix, chunks = [], [], [], []
    for i, p in enumerate(prefilling):
        if budget <= 0:
            break
        n = min(p.remaining, budget, max_chunk if max_chunk else budget)
        if n <= 0:
            continue
        ids.append(list(p.prompt[p.done:p.done + n])); sids.append(p.seq_id); prefix.append(p.done)
        chunks.append((i, n, p.done + n == len(p.prompt)))
        budget -= n
    return PiggybackStep(input_ids_list=ids + [[t] for t in decoding_last_tokens], seq_ids_list=sids + list(decoding_seq_ids),
                         decoding_seq_lens_list=list(decoding_seq_lens), prefill_prefix_lens_list=prefix if ids else None,
                         chunks=chunks)


def apply_step_result(step: PiggybackStep, prefilling: Sequence[PrefillProgress], tokens: Sequence[int]):
    """Advance the prompts by what the step processed.  Returns (first_tokens, decode_tokens): first_tokens maps the seq_id of
    every prompt whose LAST chunk ran in this step to its first generated token (the tokens sampled after non-final chunks are
    meaningless and dropped); decode_tokens are the new tokens of the decoding sequences, in order."""
    first = {}
    for (i, n, last), tok in zip(step.chunks, tokens):
        prefilling[i].done += n
        if last:
            first[prefilling[i].seq_id] = tok
    return first, list(tokens[len(step.chunks):])
