"""Reference: swiftllm/worker/kernels/paged_attn.py (paged_attention :152-222)."""
import torch

from swiftllm_b200 import _lib
from swiftllm_b200.worker.infer_state import LlamaInferState


# bench.py sets this to a list to collect (start_event, end_event) pairs around every launch (roofline timing
# of the dominant kernel on the stream it runs on).  None = no instrumentation.
TIMING_EVENTS = None


def paged_attention(
    q: torch.Tensor,                    # [num_decoding_seqs, num_q_heads, head_dim]
    k_cache: torch.Tensor,
    v_cache: torch.Tensor,
    block_table: torch.Tensor,
    model_config,
    engine_config,
    infer_state: LlamaInferState,
    cur_layer: int,
    o: torch.Tensor     # [num_decoding_seqs, num_q_heads*head_dim]
):
    qs = _lib.row_stride(q)                              # contiguous (reference) or a slice of a fused QKV output
    assert k_cache.is_contiguous()
    assert v_cache.is_contiguous()
    assert block_table.is_contiguous()
    assert o.is_contiguous()
    sbs = getattr(infer_state, "paged_attn_seq_block_size", 0)
    assert sbs % engine_config.block_size == 0
    _lib.require_device(q)
    Bd, nq, D = q.shape
    if Bd == 0:
        return
    num_blocks, num_layers, nkv, bs, _ = k_cache.shape
    l = _lib.lib()
    max_len = infer_state.max_decoding_len
    ws_bytes = l.sllm_paged_attention_workspace_bytes(Bd, nq, D, max_len, sbs, nkv)
    # the reference allocates mid_o / mid_o_logexpsum per call too (paged_attn.py:170-180)
    ws = torch.empty((ws_bytes // 4,), dtype=torch.float32, device=q.device) if ws_bytes > 0 else None
    seq_ids = infer_state.seq_ids[infer_state.num_prefill_seqs:]
    if TIMING_EVENTS is not None:
        ev0 = torch.cuda.Event(enable_timing=True); ev0.record()
    _lib.check(l.sllm_paged_attention(
        q.data_ptr(), k_cache.data_ptr(), v_cache.data_ptr(), block_table.data_ptr(), seq_ids.data_ptr(),
        infer_state.decoding_seq_lens.data_ptr(), o.data_ptr(), _lib.ptr(ws), ws_bytes,
        infer_state.softmax_scale, Bd, max_len, sbs, cur_layer, num_layers, nq, nkv, bs, D,
        block_table.shape[1], num_blocks, qs, _lib.dtype_tag(q.dtype), _lib.stream()), "paged_attention")
    if TIMING_EVENTS is not None:
        ev1 = torch.cuda.Event(enable_timing=True); ev1.record()
        TIMING_EVENTS.append((ev0, ev1))
