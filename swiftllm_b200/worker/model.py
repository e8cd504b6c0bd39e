"""LlamaModel - the data-plane worker.  Same public API as the reference's
`swiftllm.worker.model.LlamaModel` (swiftllm/worker/model.py:18-408):

    LlamaModel(engine_config) -> load_weights() -> profile_num_blocks() -> init_kvcache_and_swap(num_blocks)
    forward(input_ids_list, seq_ids_list, decoding_seq_lens_list, ignore_kvcache=False) -> list[int]
    swap_in_seqs / swap_out_seqs / free_seqs_resources

and the same observable state (k_cache, v_cache, k_swap, v_swap, gpu_block_manager, cpu_block_manager).

B200-first host path: all per-step metadata travels in ONE pinned staging buffer and one
H2D copy (the reference issues ~6 small copies, model.py:272-297); block allocation is one sync-free kernel
(three syncs + torch.nonzero in the reference, block_manager.py:50,70,75); pure-decode steps can be replayed
from a CUDA graph; the only sync of a step is the final `.tolist()` of the sampled tokens.
"""
from __future__ import annotations

import itertools
import math
from typing import Callable, Optional

import numpy as np
import torch
import torch.distributed as dist

from swiftllm_b200 import swiftllm_c
from swiftllm_b200.engine_config import EngineConfig
from swiftllm_b200.model_config import LlamaModelConfig
from swiftllm_b200.utils import GB
from swiftllm_b200.worker.block_manager import BlockManager
from swiftllm_b200.worker.weight import load_weights

from .infer_state import LlamaInferState
from .layers.post_layer import LlamaPostLayer
from .layers.pre_layer import LlamaPreLayer
from .layers.transformer_layer import LlamaTransformerLayer

_DTYPES = {"float16": torch.float16, "bfloat16": torch.bfloat16}


def select_seq_block_size(num_kv_heads: int, decoding_seq_lens_list: list, max_decoding_len: int) -> int:
    """The reference's flash-decoding split heuristic, verbatim in behaviour (model.py:305-324)."""
    seq_block_size = 2048
    decoding_seq_lens_sum = sum(decoding_seq_lens_list)
    while num_kv_heads * (decoding_seq_lens_sum / seq_block_size) < 1024 and seq_block_size // 2 >= 64 and \
            max_decoding_len / (seq_block_size // 2) <= 128:
        seq_block_size //= 2
    return seq_block_size


def build_rope_tables(model_config: LlamaModelConfig, dtype: torch.dtype):
    """cos/sin tables [max_seq_len + 128, head_dim/2] (model.py:177-225), computed in fp32 on the HOST so the
    tables are bit-identical on every device and to the CPU oracle, then rounded to the model dtype."""
    rope_scaling = model_config.rope_scaling
    base = model_config.rope_theta
    max_pos = model_config.max_position_embeddings
    dim = model_config.head_dim
    if isinstance(rope_scaling, dict):      # Llama 3.2 style dictionary (model.py:183-211)
        factor = rope_scaling.get("factor", 4.0)
        low = rope_scaling.get("low_freq_factor", 1.0)
        high = rope_scaling.get("high_freq_factor", 1.0)
        orig = rope_scaling.get("original_max_position_embeddings", max_pos)
        max_seq_len = int(orig * factor)
        t = torch.arange(max_seq_len + 128, dtype=torch.float32)
        split = int((dim // 2) * low / (low + high))
        inv_low = 1.0 / (base ** (torch.arange(0, split * 2, 2, dtype=torch.float32) / dim))
        inv_high = 1.0 / (base ** (torch.arange(split * 2, dim, 2, dtype=torch.float32) / dim))
        freqs = torch.cat([torch.outer(t / low, inv_low), torch.outer(t / high, inv_high)], dim=-1)
    else:
        max_seq_len = max_pos * rope_scaling
        inv_freq = 1.0 / (base ** (torch.arange(0, dim, 2, dtype=torch.float32) / dim))
        t = torch.arange(int(max_seq_len + 128), dtype=torch.float32) / rope_scaling
        freqs = torch.outer(t, inv_freq)
    return torch.cos(freqs).to(dtype), torch.sin(freqs).to(dtype)


class LlamaModel:
    @torch.inference_mode()
    def __init__(self, engine_config: EngineConfig, model_config: Optional[LlamaModelConfig] = None):
        """`model_config` may be given directly (synthetic-weight runs); otherwise it is read from
        engine_config.model_path/config.json like the reference (model.py:42)."""
        self.engine_config = engine_config
        self.model_config = model_config if model_config is not None else \
            LlamaModelConfig.load_from_model_path(engine_config.model_path)
        self.dtype = _DTYPES[getattr(engine_config, "dtype", "float16")]
        self.tp_size = getattr(engine_config, "tp_size", 1)
        self.tp_rank = getattr(engine_config, "tp_rank", 0)
        mc = self.model_config
        assert mc.num_q_heads % self.tp_size == 0 and mc.num_kv_heads % self.tp_size == 0 and \
            mc.ffn_inter_dim % self.tp_size == 0, "tp_size must divide q heads, kv heads and ffn_inter_dim"
        self.device = torch.device("cuda", torch.cuda.current_device())

        self.weight = None
        self._cos_cached = self._sin_cached = None
        self.pre_layer = None
        self.transformer_layers = None
        self.post_layer = None
        self.num_blocks = None
        self.k_cache = self.v_cache = None
        self.k_swap = self.v_swap = None
        self.cpu_block_manager = self.gpu_block_manager = None
        self.tp_group = None
        self._graphs = {}
        # pinned staging ring for per-step metadata (eager path): a slot is rewritten only after the H2D copy that last read it
        # has completed (event per slot), so back-to-back forward_async calls never race the copy engine
        self._stage_ring = []
        self._stage_next = 0
        self._swap_stream = None        # EngineConfig.swap_on_copy_stream: dedicated copy stream + "copies done" event
        self._swap_done = None

    # ------------------------------------------------------------------ init
    @torch.inference_mode()
    def load_weights(self, weight_getter: Optional[Callable] = None):
        """Load weights (checkpoint, dummy, or a custom getter(key, shape, dtype) -> full tensor) and build layers."""
        if self.tp_size > 1:
            assert dist.is_initialized(), "tensor parallelism needs torch.distributed (NCCL) to be initialised"
            self.tp_group = dist.group.WORLD
        self.weight = load_weights(self.model_config, self.dtype, self.engine_config.model_path,
                                   self.engine_config.use_dummy, getter=weight_getter,
                                   tp_rank=self.tp_rank, tp_size=self.tp_size, device=self.device,
                                   shard_lm_head=self._shard_lm_head())
        cos, sin = build_rope_tables(self.model_config, self.dtype)
        self._cos_cached, self._sin_cached = cos.to(self.device), sin.to(self.device)

        self.comm = None
        want_fused = getattr(self.engine_config, "fused_allreduce", None)
        explicit = want_fused is not None
        if want_fused is None:
            # measured on B200 / NVSwitch (profiles/r2_tp_sweep_*.jsonl), decode batch 256 x 4096:
            #   tp 2, 4: one-shot peer-memory kernel (14 / 21 us per exchange vs NCCL 22 / 30 us + the add/norm launch)
            #   tp 8:    row-owner kernel with in-switch reduction and broadcast (24.6 us vs NCCL 33.7 + 3.2 us)
            candidates = [True] if 2 <= self.tp_size <= 4 else (["two_shot_nvls", "two_shot"] if self.tp_size == 8 else [])
        else:
            candidates = [want_fused] if want_fused else []
        for cand in candidates if self.tp_size > 1 else []:
            two_shot = cand in ("two_shot", "two_shot_nvls", "ll", "ll_nvls")      # row-owner exchanges
            nvls = cand in ("two_shot_nvls", "ll_nvls")                            # in-switch reduction / broadcast (multimem)
            ll = cand in ("ll", "ll_nvls")                                         # barrier-free push kernel
            from swiftllm_b200.worker.tp_comm import FusedAllReduce
            comm = None
            try:
                comm = FusedAllReduce(self.engine_config.max_tokens_in_batch, self.model_config.hidden_size,
                                      self.dtype, self.device, self.tp_group, two_shot=two_shot, nvls=nvls, ll=ll)
            except Exception as e:      # noqa: BLE001  (no peer access / symmetric memory / multicast unavailable)
                if explicit:
                    raise
                import warnings
                warnings.warn(f"fused exchange `{cand}` unavailable ({e}); trying the next option")
            # every rank must take the same path
            ok = torch.tensor([1 if comm is not None else 0], device=self.device)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=self.tp_group)
            if int(ok.item()) == 1:
                self.comm = comm
                break
            del comm
        decoding_piggyback_stream = torch.cuda.Stream()
        self.pre_layer = LlamaPreLayer(self.model_config, self.weight)
        self.transformer_layers = [
            LlamaTransformerLayer(self.model_config, self.engine_config, self.weight.layers[i],
                                  decoding_piggyback_stream, i, tp_group=self.tp_group, comm=self.comm)
            for i in range(self.model_config.num_layers)
        ]
        self.post_layer = LlamaPostLayer(self.model_config, self.weight, tp_group=self.tp_group, tp_rank=self.tp_rank,
                                         tp_size=self.tp_size)

    def _shard_lm_head(self) -> bool:
        want = getattr(self.engine_config, "shard_lm_head", None)
        if want is None:
            want = self.tp_size >= 2 and self.model_config.vocab_size % self.tp_size == 0
        return bool(want) and self.tp_size > 1

    def _kvslot_bytes(self) -> int:
        return self.model_config.get_kvslot_size(self.dtype) // self.tp_size

    @torch.inference_mode()
    def profile_num_blocks(self) -> int:
        """model.py:94-131: forged maximum-size prefill, peak memory -> number of KV blocks that fit."""
        torch.cuda.empty_cache()
        torch.cuda.reset_peak_memory_stats()
        num_tokens = self.engine_config.max_tokens_in_batch
        batch_size = self.engine_config.max_batch_size
        input_lens = [num_tokens // batch_size] * batch_size
        input_lens[-1] += num_tokens % batch_size
        input_ids = [[0] * n for n in input_lens]
        self.k_cache = self.v_cache = None
        _ = self.forward(input_ids, list(range(batch_size)), [], ignore_kvcache=True)
        torch.cuda.synchronize()
        free_memory, total_memory = torch.cuda.mem_get_info()
        peak_memory = total_memory - free_memory
        useable_memory = total_memory * self.engine_config.gpu_mem_utilization
        print(f"[Model.profile] GPU total memory: {total_memory/GB:.2f} GB, runtime peak memory: {peak_memory/GB:.2f} GB")
        if useable_memory < peak_memory:
            raise RuntimeError(f"Peak memory {peak_memory/GB:.2f} GB exceeds usable memory {useable_memory/GB:.2f} GB "
                               f"({total_memory/GB:.2f} GB * {self.engine_config.gpu_mem_utilization})")
        block_size_bytes = self.engine_config.block_size * self._kvslot_bytes()
        num_gpu_blocks = math.floor((useable_memory - peak_memory) / block_size_bytes)
        torch.cuda.empty_cache()
        if self.tp_size > 1:
            # every rank must run the SAME allocator state (block ids and the out-of-blocks RuntimeError are decided on each
            # rank's host mirror): take the minimum over ranks
            t = torch.tensor([num_gpu_blocks], dtype=torch.int64, device=self.device)
            dist.all_reduce(t, op=dist.ReduceOp.MIN, group=self.tp_group)
            num_gpu_blocks = int(t.item())
        return num_gpu_blocks

    @torch.inference_mode()
    def init_kvcache_and_swap(self, num_blocks: int):
        """model.py:134-175.  KV cache [num_blocks, L, nkv/tp, block_size, D] (zeros, like the reference), CPU swap
        space (pinned here), and the two block managers."""
        if self.tp_size > 1 and dist.is_initialized():
            # identical num_blocks on every rank <=> identical block ids and a collective out-of-blocks decision (all ranks
            # raise the same RuntimeError at the same step instead of one rank leaving its peers inside an exchange)
            t = torch.tensor([num_blocks, -num_blocks], dtype=torch.int64, device=self.device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.tp_group)
            if int(t[0].item()) != num_blocks or int(-t[1].item()) != num_blocks:
                raise RuntimeError(f"init_kvcache_and_swap: num_blocks differs between tensor-parallel ranks "
                                   f"(this rank {num_blocks}, min {int(-t[1].item())}, max {int(t[0].item())}); use the value "
                                   f"profile_num_blocks() returns (it is min-reduced over ranks)")
        self.num_blocks = num_blocks
        mc, ec = self.model_config, self.engine_config
        shape = (num_blocks, mc.num_layers, mc.num_kv_heads // self.tp_size, ec.block_size, mc.head_dim)
        self.k_cache = torch.zeros(shape, dtype=self.dtype, device=self.device)
        self.v_cache = torch.zeros(shape, dtype=self.dtype, device=self.device)
        sshape = (ec.num_cpu_blocks,) + shape[1:]
        pin = bool(getattr(ec, "pin_swap_space", True)) and ec.num_cpu_blocks > 0
        self.k_swap = torch.zeros(sshape, dtype=self.dtype, device="cpu", pin_memory=pin)
        self.v_swap = torch.zeros(sshape, dtype=self.dtype, device="cpu", pin_memory=pin)
        self.gpu_block_manager = BlockManager("GPU", num_blocks, ec.max_seqs_in_block_table, ec.max_blocks_per_seq,
                                              ec.block_size, device=self.device)
        self.cpu_block_manager = BlockManager("CPU", ec.num_cpu_blocks, ec.max_seqs_in_block_table,
                                              ec.max_blocks_per_seq, ec.block_size, device=self.device)
        self._graphs = {}

    # ------------------------------------------------------------------ forward
    @torch.inference_mode()
    def _forward(self, input_ids: torch.Tensor, infer_state: LlamaInferState, k_cache=None, v_cache=None,
                 block_table=None) -> torch.Tensor:
        """model.py:228-249."""
        input_embds = self.pre_layer.forward(input_ids)
        self._wait_for_swaps()          # first KV-cache access of the step is layer 0's store: pending swap copies must be done
        residual_buf = torch.zeros_like(input_embds)
        k_cache = self.k_cache if k_cache is None else k_cache
        v_cache = self.v_cache if v_cache is None else v_cache
        if block_table is None and not infer_state.ignore_kvcache:
            block_table = self.gpu_block_manager.block_table
        for layer in self.transformer_layers:
            input_embds = layer.forward(input_embds, residual_buf, k_cache, v_cache, block_table, infer_state)
        if self.comm is not None and residual_buf.shape[0] <= self.comm.max_fused_tokens:
            # the last layer's down_proj partials are still un-reduced: exchange + add + final RMSNorm in one kernel
            normed = self.comm.reduce_add_norm(1, residual_buf.shape[0], residual_buf, self.weight.final_norm,
                                               self.model_config.rms_norm_eps)
            return self.post_layer.forward(normed, infer_state, already_normed=True)
        input_embds += residual_buf
        return self.post_layer.forward(input_embds, infer_state)

    _STAGE_SLOTS = 4

    def _stage_slot(self, n: int):
        """Next slot of the persistent pinned ring, grown geometrically; waits (host side) for the copy that last used it."""
        i = self._stage_next
        self._stage_next = (i + 1) % self._STAGE_SLOTS
        if len(self._stage_ring) <= i:
            self._stage_ring.append(None)
        slot = self._stage_ring[i]
        if slot is None or slot[0].numel() < n:
            cap = max(4096, 1 << (max(n, 1) - 1).bit_length())
            if slot is not None:
                slot[1].synchronize()
            slot = (torch.empty((cap,), dtype=torch.int32).pin_memory(), torch.cuda.Event())
            self._stage_ring[i] = slot
        else:
            slot[1].synchronize()                # no-op unless the copy engine is more than _STAGE_SLOTS steps behind
        return slot

    def _wait_for_swaps(self):
        """Device-side wait (no host sync) for swap copies enqueued on the copy stream since the last forward."""
        if self._swap_done is not None and not (torch.cuda.is_available() and torch.cuda.is_current_stream_capturing()):
            torch.cuda.current_stream().wait_event(self._swap_done)
            self._swap_done = None

    def _stage_metadata(self, *parts):
        """One persistent pinned host buffer (ring of slots), one async H2D copy; returns int32 device views (one per list)."""
        flat = np.fromiter(itertools.chain(*parts), dtype=np.int32)
        host, ev = self._stage_slot(flat.size)
        host[:flat.size].numpy()[:] = flat
        dev = host[:flat.size].to(self.device, non_blocking=True)
        ev.record()
        views, off = [], 0
        for p in parts:
            views.append(dev[off:off + len(p)])
            off += len(p)
        return views

    @torch.inference_mode()
    def forward(self, input_ids_list: list, seq_ids_list: list, decoding_seq_lens_list: list,
                ignore_kvcache: bool = False, prefill_prefix_lens_list: Optional[list] = None) -> list:
        """model.py:252-359.  Batch layout contract: prefill sequences first (full prompt each), then decoding
        sequences (exactly one token each); decoding_seq_lens_list[i] includes the new token.

        `prefill_prefix_lens_list` (addition, SURVEY.md §8 f-1: chunked / SARATHI-style prefill, which the reference's
        forward cannot express): when given, prefill entry i is the CHUNK of its prompt that starts at position
        prefill_prefix_lens_list[i]; positions below it were written to the KV cache by earlier calls.  The chunk's
        K/V are stored at their positions and its queries attend to the whole prefix + chunk through the block
        table.  The token returned for a prefill entry is sampled after the chunk's last token (meaningful for the
        final chunk of a prompt only)."""
        return self.forward_async(input_ids_list, seq_ids_list, decoding_seq_lens_list, ignore_kvcache,
                                  prefill_prefix_lens_list).tolist()

    @torch.inference_mode()
    def forward_async(self, input_ids_list, seq_ids_list, decoding_seq_lens_list, ignore_kvcache=False,
                      prefill_prefix_lens_list=None) -> torch.Tensor:
        """Same as forward() but returns the token tensor on the device (no sync)."""
        mc = self.model_config
        num_prefill_seqs = len(input_ids_list) - len(decoding_seq_lens_list)
        flat_ids = list(itertools.chain(*input_ids_list))
        prefill_lens = [len(s) for s in input_ids_list[:num_prefill_seqs]]
        chunked = prefill_prefix_lens_list is not None and num_prefill_seqs > 0
        if chunked:
            prefix_lens = [int(x) for x in prefill_prefix_lens_list]
            assert len(prefix_lens) == num_prefill_seqs and min(prefix_lens) >= 0, \
                "prefill_prefix_lens_list needs one non-negative entry per prefill sequence"
            assert not ignore_kvcache, "chunked prefill reads the KV cache: ignore_kvcache is not applicable"
            # first chunks only (every prefix 0): nothing to read back from the cache, so this IS the reference's whole-prompt
            # contract - packed-k/v attention and the plain store
            chunked = max(prefix_lens) > 0
        else:
            prefix_lens = [0] * num_prefill_seqs
        seq_lengths_list = [p + n for p, n in zip(prefix_lens, prefill_lens)] + list(decoding_seq_lens_list)
        batch_size, num_tokens = len(input_ids_list), len(flat_ids)
        num_prefill_tokens = num_tokens - (batch_size - num_prefill_seqs)
        prefill_starts = list(itertools.accumulate([0] + prefill_lens[:-1])) if prefill_lens else []
        max_prefill_len = max(prefill_lens) if prefill_lens else 0
        max_decoding_len = max(decoding_seq_lens_list) if decoding_seq_lens_list else 0
        positions = [p0 + p for p0, n in zip(prefix_lens, prefill_lens) for p in range(n)] + \
            [l - 1 for l in decoding_seq_lens_list]
        last_idx = [s + n - 1 for s, n in zip(prefill_starts, prefill_lens)] + list(range(num_prefill_tokens, num_tokens))

        if (getattr(self.engine_config, "use_cuda_graph", False) and num_prefill_seqs == 0 and not ignore_kvcache
                and batch_size > 0):
            return self._forward_decode_graph(flat_ids, seq_ids_list, list(decoding_seq_lens_list), max_decoding_len)

        ids, seq_ids, seq_lengths, p_lens, p_starts, pos, last, p_prefix = self._stage_metadata(
            flat_ids, seq_ids_list, seq_lengths_list, prefill_lens, prefill_starts, positions, last_idx,
            prefix_lens if chunked else [])

        if not ignore_kvcache:
            self.gpu_block_manager.allocate_blocks_for_seqs(seq_ids, seq_lengths, seq_ids_list=seq_ids_list,
                                                            target_lens_list=seq_lengths_list, want_new_blocks=False)

        seq_block_size = select_seq_block_size(mc.num_kv_heads, list(decoding_seq_lens_list), max_decoding_len)
        infer_state = LlamaInferState(
            batch_size=batch_size, num_tokens=num_tokens,
            seq_ids=seq_ids, softmax_scale=mc.head_dim ** -0.5,
            num_prefill_seqs=num_prefill_seqs, num_prefill_tokens=num_prefill_tokens,
            prefill_seq_start_locs=p_starts,
            # ends at num_prefill_tokens (the reference puts num_tokens there, model.py:336-339, which over-declares
            # the last prompt by the number of decoding tokens on mixed batches)
            prefill_seq_start_locs_with_end=torch.cat([p_starts, p_starts.new_tensor([num_prefill_tokens])])
            if num_prefill_seqs > 0 else p_starts,
            prefill_seq_lens=p_lens, max_prefill_len=max_prefill_len,
            num_decoding_seqs=batch_size - num_prefill_seqs, decoding_seq_lens=seq_lengths[num_prefill_seqs:],
            max_decoding_len=max_decoding_len,
            seq_block_size=seq_block_size, num_seq_blocks=(max_decoding_len + seq_block_size - 1) // seq_block_size,
            position_cos=self._cos_cached.index_select(0, pos), position_sin=self._sin_cached.index_select(0, pos),
            ignore_kvcache=ignore_kvcache,
            paged_attn_seq_block_size=0, last_token_indices=last,
            prefill_prefix_lens=p_prefix if chunked else None,
            max_prefill_kv_len=max(seq_lengths_list[:num_prefill_seqs]) if chunked else 0,
        )
        return self._forward(ids, infer_state)

    # ------------------------------------------------------------------ CUDA-graph decode path
    def _forward_decode_graph(self, flat_ids, seq_ids_list, lens_list, max_len):
        """Pure-decode step replayed from a CUDA graph.  Static device buffers hold the step's metadata; the graph
        (keyed by batch size and the attention launch geometry) contains allocation-free work only: block
        allocation, gathers, every layer, lm_head and argmax."""
        from swiftllm_b200 import _lib
        B = len(flat_ids)
        mc = self.model_config
        nkv = mc.num_kv_heads // self.tp_size
        nq = mc.num_q_heads // self.tp_size
        # launch geometry of paged attention depends on max_len only through the split count: bucket it
        bucket = ((max_len + 1023) // 1024) * 1024
        ws_bytes = _lib.lib().sllm_paged_attention_workspace_bytes(B, nq, mc.head_dim, bucket, 0, nkv)
        key = (B, bucket if ws_bytes > 0 else 0)
        g = self._graphs.pop(key, None)
        if g is None:
            # bounded cache (least recently used graph dropped first); all graphs share ONE memory pool - they are replayed
            # one at a time on one stream and their only output (`tokens`) is copied out before the next replay
            cap = max(1, int(getattr(self.engine_config, "max_cuda_graphs", 16)))
            while len(self._graphs) >= cap:
                self._graphs.pop(next(iter(self._graphs)))
            g = self._capture_decode_graph(B, bucket)
        self._graphs[key] = g                    # (re)insert as most recently used
        # host-side bookkeeping of the allocator (exhaustion check + mirror), no device work here
        bm = self.gpu_block_manager
        idx = np.asarray(seq_ids_list, dtype=np.int64)
        target = (np.asarray(lens_list, dtype=np.int64) + bm.block_size - 1) // bm.block_size
        have = bm._host_nsab[idx]
        assert (have <= target).all()
        total = int((target - have).sum())
        if total > bm.num_free_blocks:
            raise RuntimeError(f"No enough free blocks available on GPU ({bm.num_blocks} in total, "
                               f"{bm.num_free_blocks} free, {total} requested)")
        bm._host_nsab[idx] = target
        bm.num_free_blocks -= total
        # pinned staging: a ring of (buffer, event) pairs per graph; a buffer is rewritten only after the H2D copy that last
        # read it has finished, so back-to-back forward_async calls are safe
        i = g["next"]; g["next"] = (i + 1) % len(g["host"])
        host, ev = g["host"][i]
        ev.synchronize()
        hv = host.numpy()
        hv[:B] = flat_ids; hv[B:2 * B] = seq_ids_list; hv[2 * B:3 * B] = lens_list
        g["meta"].copy_(host, non_blocking=True)
        ev.record()
        self._wait_for_swaps()
        g["graph"].replay()
        # the graph's static output is overwritten by the next replay: hand out a copy (stream-ordered, no sync)
        return g["tokens"].clone()

    def _capture_decode_graph(self, B: int, bucket_len: int):
        from swiftllm_b200.worker.kernels.block_mgmt import allocate_blocks_for_seqs as alloc_kernel
        mc, ec = self.model_config, self.engine_config
        dev = self.device
        meta = torch.zeros((3 * B,), dtype=torch.int32, device=dev)
        host = [(torch.zeros((3 * B,), dtype=torch.int32).pin_memory(), torch.cuda.Event()) for _ in range(self._STAGE_SLOTS)]
        ids, seq_ids, lens = meta[:B], meta[B:2 * B], meta[2 * B:]
        last = torch.arange(B, dtype=torch.int32, device=dev)
        empty = torch.empty((0,), dtype=torch.int32, device=dev)

        def step(bm, k_cache, v_cache):
            alloc_kernel(bm.num_seq_allocated_blocks, bm.block_table, bm.is_block_free, seq_ids, lens, bm.block_size,
                         None, bm._status)
            pos = (lens - 1).long()
            st = LlamaInferState(
                batch_size=B, num_tokens=B, seq_ids=seq_ids, softmax_scale=mc.head_dim ** -0.5,
                num_prefill_seqs=0, num_prefill_tokens=0, prefill_seq_start_locs=empty,
                prefill_seq_start_locs_with_end=empty, prefill_seq_lens=empty, max_prefill_len=0,
                num_decoding_seqs=B, decoding_seq_lens=lens, max_decoding_len=bucket_len,
                seq_block_size=2048, num_seq_blocks=(bucket_len + 2047) // 2048,
                position_cos=self._cos_cached.index_select(0, pos), position_sin=self._sin_cached.index_select(0, pos),
                ignore_kvcache=False, paged_attn_seq_block_size=0, last_token_indices=last)
            return self._forward(ids, st, k_cache=k_cache, v_cache=v_cache, block_table=bm.block_table)

        # Warm-up (library handles, lazy kernel loading, GEMM heuristics for these shapes) must not touch the real
        # cache or block tables: it runs against a one-block scratch cache with B one-token sequences.
        scratch_bm = BlockManager("scratch", B, B, 1, ec.block_size, device=dev)
        sshape = (B,) + tuple(self.k_cache.shape[1:])
        k_s = torch.zeros(sshape, dtype=self.dtype, device=dev)
        v_s = torch.zeros(sshape, dtype=self.dtype, device=dev)
        meta.zero_()
        seq_ids.copy_(torch.arange(B, dtype=torch.int32, device=dev))
        lens.fill_(1)
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(2):
                step(scratch_bm, k_s, v_s)
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        del k_s, v_s, scratch_bm
        graph = torch.cuda.CUDAGraph()
        if getattr(self, "_graph_pool", None) is None:
            self._graph_pool = torch.cuda.graph_pool_handle()
        with torch.cuda.graph(graph, pool=self._graph_pool):
            tokens = step(self.gpu_block_manager, self.k_cache, self.v_cache)
        torch.cuda.synchronize()
        # `last` / `empty` are read by the captured kernels on every replay: keep them alive with the graph
        return dict(graph=graph, meta=meta, host=host, next=0, tokens=tokens, keep=(last, empty))

    # ------------------------------------------------------------------ swap / free
    def _swap(self, seq_ids_list: list, is_swap_in: bool):
        """model.py:361-379, without the two `.tolist()` syncs on the allocation side: the host mirror knows the
        sizes; only the gathered / new block ids themselves have to come back for the memcpy plan."""
        src = self.cpu_block_manager if is_swap_in else self.gpu_block_manager
        dst = self.gpu_block_manager if is_swap_in else self.cpu_block_manager
        seq_ids = torch.tensor(seq_ids_list, dtype=torch.int32, device=self.device)
        nblocks = src.get_num_allocated_blocks_host(seq_ids_list)
        seq_lengths_list = [n * self.engine_config.block_size for n in nblocks]
        seq_lengths = torch.tensor(seq_lengths_list, dtype=torch.int32, device=self.device)
        src_block_ids = src.gather_allocated_blocks_and_free(seq_ids, seq_ids_list=seq_ids_list)
        dst_block_ids = dst.allocate_blocks_for_seqs(seq_ids, seq_lengths, seq_ids_list=seq_ids_list,
                                                     target_lens_list=seq_lengths_list)
        device_ids = bool(getattr(self.engine_config, "device_swap", False))
        if device_ids:
            src_ids, dst_ids = src_block_ids.to(torch.int64), dst_block_ids
        else:
            src_ids, dst_ids = src_block_ids.tolist(), dst_block_ids.tolist()
        copy = swiftllm_c.swap_blocks_device if device_ids else swiftllm_c.swap_blocks
        if not getattr(self.engine_config, "swap_on_copy_stream", False):
            # ids stay on the device (device_swap): one gather/scatter kernel over the mapped swap space, sync-free
            copy(src_ids, dst_ids, is_swap_in, self.k_cache, self.v_cache, self.k_swap, self.v_swap)
            return
        # copy stream: the copy starts once everything enqueued so far on this stream (the block-table kernels above, the KV
        # stores of earlier steps) has finished; the next forward waits for `_swap_done` before it touches the cache
        if self._swap_stream is None:
            self._swap_stream = torch.cuda.Stream(device=self.device)
        ready = torch.cuda.Event()
        ready.record()
        with torch.cuda.stream(self._swap_stream):
            self._swap_stream.wait_event(ready)
            copy(src_ids, dst_ids, is_swap_in, self.k_cache, self.v_cache, self.k_swap, self.v_swap)
            if device_ids and src_ids.is_cuda:       # the id tensors were allocated on the main stream
                src_ids.record_stream(self._swap_stream); dst_ids.record_stream(self._swap_stream)
            self._swap_done = torch.cuda.Event()
            self._swap_done.record()

    @torch.inference_mode()
    def swap_in_seqs(self, seq_ids_list: list):
        """Move the blocks of the given sequences from the CPU swap space to the GPU cache."""
        self._swap(seq_ids_list, True)

    @torch.inference_mode()
    def swap_out_seqs(self, seq_ids_list: list):
        """Move the blocks of the given sequences from the GPU cache to the CPU swap space."""
        self._swap(seq_ids_list, False)

    @torch.inference_mode()
    def free_seqs_resources(self, seq_ids_list: list):
        """model.py:401-408."""
        seq_ids = torch.tensor(seq_ids_list, dtype=torch.int32, device=self.device)
        self.gpu_block_manager.free_blocks_for_seqs(seq_ids, seq_ids_list=seq_ids_list)
        self.cpu_block_manager.free_blocks_for_seqs(seq_ids, seq_ids_list=seq_ids_list)
