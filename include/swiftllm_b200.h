/*
 * swiftllm_b200.h - C ABI of the B200-native (sm_100a) data plane that replaces every kernel in
 * swiftLLM's `swiftllm/worker/kernels/` plus the native `swiftllm_c` module.
 *
 * Drop-in boundary: the reference binds its kernels as Python functions on torch tensors (one wrapper
 * per op, in place / into caller-provided outputs, current CUDA stream).  This library exports exactly
 * one entry point per wrapper, taking raw device pointers, sizes, a dtype tag and the stream; the host
 * package `swiftllm_b200/worker/kernels/*.py` binds them with ctypes under the reference's own names and
 * signatures (see INTEGRATION.md for the stub a swiftLLM maintainer would add).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the parameter name says `host_`;
 *   - tensors are contiguous in the reference's layouts (cited per function);
 *   - all functions enqueue on `stream` (a cudaStream_t) and return immediately;
 *   - return value: 0 on success, non-zero on error (`sllm_last_error()` gives the message).  Invalid
 *     shapes are rejected before anything is launched (the reference raises AssertionError there);
 *   - no CPU fallback exists: on a non-sm_100 device every launch fails with an error.
 *
 * Reference paths below are relative to the swiftLLM repository root.
 */
#ifndef SWIFTLLM_B200_H
#define SWIFTLLM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SLLM_ABI_VERSION 1

typedef enum { SLLM_F16 = 0, SLLM_BF16 = 1 } sllm_dtype_t;
typedef void* sllm_stream_t; /* cudaStream_t */

/* ---- library ---- */
int sllm_abi_version(void);
const char* sllm_last_error(void);          /* thread-local message of the last failing call */
int sllm_device_check(int device);          /* 0 iff `device` is compute capability 10.x */

/* ---- RMSNorm: swiftllm/worker/kernels/rmsnorm.py:26-37 (rmsnorm_inplace), :67-89 (fused_add_rmsnorm_inplace)
 * x [num_tokens, hidden] in place; weight [hidden].  fused: residual <- x + residual (rounded to dtype),
 * x <- rmsnorm(residual) * weight.  hidden % 8 == 0. */
int sllm_rmsnorm_inplace(void* x, const void* weight, float eps, int64_t num_tokens, int hidden,
                         sllm_dtype_t dtype, sllm_stream_t stream);
int sllm_fused_add_rmsnorm_inplace(void* x, void* residual, const void* weight, float eps, int64_t num_tokens,
                                   int hidden, sllm_dtype_t dtype, sllm_stream_t stream);

/* ---- Rotary embedding: swiftllm/worker/kernels/rotary_emb.py:44-58 (rotary_embedding_inplace)
 * q [T, nq, D], k [T, nkv, D] in place; cos/sin [T, D/2] (infer_state.position_cos/sin).  NeoX half-split.
 * D % 16 == 0.  *_row_stride: elements between consecutive token rows (nq*D / nkv*D when contiguous, as in the
 * reference; larger when q/k/v are column slices of one fused QKV GEMM output); heads stay contiguous within a token. */
int sllm_rotary_embedding_inplace(void* q, void* k, const void* cos, const void* sin, int64_t num_tokens,
                                  int num_q_heads, int num_kv_heads, int head_dim, int64_t q_row_stride,
                                  int64_t k_row_stride, sllm_dtype_t dtype, sllm_stream_t stream);

/* ---- SwiGLU gate: swiftllm/worker/kernels/silu_and_mul.py:25-34 (silu_and_mul_inplace)
 * x [T, 2*ffn_inter_dim] = [up | gate]; x[:, :F] <- up * silu(gate).  F % 8 == 0. */
int sllm_silu_and_mul_inplace(void* x, int64_t num_tokens, int64_t ffn_inter_dim, sllm_dtype_t dtype,
                              sllm_stream_t stream);

/* ---- KV-cache store: swiftllm/worker/kernels/kvcache_mgmt.py:81-122 (store_kvcache)
 * k,v [num_tokens, nkv, D] (prefill tokens first, then one row per decoding seq);
 * caches [num_blocks, num_layers, nkv, block_size, D]; block_table int32 [*, max_blocks_per_seq];
 * seq_ids int32 [num_prefill_seqs + num_decoding_seqs]. */
int sllm_store_kvcache(const void* k, const void* v, void* k_cache, void* v_cache, const int32_t* block_table,
                       const int32_t* seq_ids, const int32_t* prefill_seq_start_locs,
                       const int32_t* prefill_seq_lens, const int32_t* decoding_seq_lens, int num_prefill_seqs,
                       int num_decoding_seqs, int64_t num_prefill_tokens, int max_prefill_len, int cur_layer,
                       int num_layers, int num_kv_heads, int block_size, int head_dim, int max_blocks_per_seq,
                       int64_t k_row_stride, int64_t v_row_stride, sllm_dtype_t dtype, sllm_stream_t stream);

/* ---- Decode-step fusion (new): rotary_embedding_inplace (rotary_emb.py:44-58) + the decode part of store_kvcache
 * (kvcache_mgmt.py:50-79) in ONE launch for a batch of decoding rows only: q [Bd, nq, D] and k [Bd, nkv, D] rotated in place
 * with cos/sin [Bd, D/2]; the rotated k row and the v row of sequence i are written to position decoding_seq_lens[i] - 1 of
 * block-table row seq_ids[i].  Same results, bit for bit, as the two separate calls. */
int sllm_rotary_store_kvcache_decode(void* q, void* k, const void* v, const void* cos, const void* sin, void* k_cache,
                                     void* v_cache, const int32_t* block_table, const int32_t* seq_ids,
                                     const int32_t* decoding_seq_lens, int num_decoding_seqs, int cur_layer,
                                     int num_layers, int num_q_heads, int num_kv_heads, int block_size, int head_dim,
                                     int max_blocks_per_seq, int64_t q_row_stride, int64_t k_row_stride,
                                     int64_t v_row_stride, sllm_dtype_t dtype, sllm_stream_t stream);

/* ---- Paged (decode) attention: swiftllm/worker/kernels/paged_attn.py:152-222 (paged_attention)
 * q [Bd, nq, D]; o [Bd, nq*D]; seq_ids = infer_state.seq_ids[num_prefill_seqs:]; seq_lens = decoding_seq_lens.
 * seq_block_size: tokens per flash-decoding split (multiple of block_size); 0 = let the library choose
 * (v1 single pass when the batch alone fills the GPU, v2 split + merge otherwise).
 * workspace: fp32 scratch for split partials, >= sllm_paged_attention_workspace_bytes(...) bytes
 * (the reference allocates mid_o / mid_o_logexpsum per call, paged_attn.py:170-180). */
int64_t sllm_paged_attention_workspace_bytes(int num_decoding_seqs, int num_q_heads, int head_dim, int max_seq_len,
                                             int seq_block_size, int num_kv_heads);
int sllm_paged_attention(const void* q, const void* k_cache, const void* v_cache, const int32_t* block_table,
                         const int32_t* seq_ids, const int32_t* seq_lens, void* o, void* workspace,
                         int64_t workspace_bytes, float softmax_scale, int num_decoding_seqs, int max_seq_len,
                         int seq_block_size, int cur_layer, int num_layers, int num_q_heads, int num_kv_heads,
                         int block_size, int head_dim, int max_blocks_per_seq, int64_t num_blocks,
                         int64_t q_row_stride, sllm_dtype_t dtype, sllm_stream_t stream);

/* ---- Prefill attention: swiftllm/worker/kernels/prefill_attn.py:102-139 (prefill_attention) and the
 * flash_attn_varlen_func call it stands for (swiftllm/worker/layers/transformer_layer.py:86-96).
 * Causal, varlen, packed: q,o [Tp, nq, D]; k,v [Tp, nkv, D]; start_locs/seq_lens int32 [num_prefill_seqs]. */
int sllm_prefill_attention(const void* q, const void* k, const void* v, void* o, const int32_t* prefill_seq_start_locs,
                           const int32_t* prefill_seq_lens, float softmax_scale, int num_prefill_seqs,
                           int max_prefill_len, int64_t num_prefill_tokens, int num_q_heads, int num_kv_heads,
                           int head_dim, int64_t q_row_stride, int64_t k_row_stride, int64_t v_row_stride,
                           sllm_dtype_t dtype, sllm_stream_t stream);

/* ---- Chunked ("prefix-aware") prefill - SURVEY.md §8 f-1.  NEW: the reference's forward cannot express a partial prompt
 * (its prefill attention only sees the packed k/v of the batch: swiftllm/worker/kernels/prefill_attn.py:102-139,
 * swiftllm/worker/layers/transformer_layer.py:86-96; its store always starts at position 0: kvcache_mgmt.py:10-48).
 * Prefill entry i of the batch is the CHUNK of its prompt that covers positions
 * [prefill_prefix_lens[i], prefill_prefix_lens[i] + prefill_seq_lens[i]); earlier positions are already in the cache.
 *   sllm_store_kvcache_chunked   = sllm_store_kvcache with chunk token t written to position prefix + t;
 *   sllm_prefill_attention_paged = causal attention of the chunk's queries (q, o [Tp, nq, D] packed) over positions
 *       0 .. prefix + chunk - 1 read from the paged caches through rows `seq_ids` of the block table (the chunk's own
 *       K/V must have been stored first); query t sees positions <= prefix + t.  With all prefixes 0 it computes
 *       exactly what sllm_prefill_attention computes.  head_dim 64 / 128. */
int sllm_store_kvcache_chunked(const void* k, const void* v, void* k_cache, void* v_cache, const int32_t* block_table,
                               const int32_t* seq_ids, const int32_t* prefill_seq_start_locs,
                               const int32_t* prefill_seq_lens, const int32_t* prefill_prefix_lens,
                               const int32_t* decoding_seq_lens, int num_prefill_seqs, int num_decoding_seqs,
                               int64_t num_prefill_tokens, int max_prefill_len, int cur_layer, int num_layers,
                               int num_kv_heads, int block_size, int head_dim, int max_blocks_per_seq,
                               int64_t k_row_stride, int64_t v_row_stride, sllm_dtype_t dtype, sllm_stream_t stream);
int sllm_prefill_attention_paged(const void* q, const void* k_cache, const void* v_cache, void* o,
                                 const int32_t* block_table, const int32_t* seq_ids,
                                 const int32_t* prefill_seq_start_locs, const int32_t* prefill_seq_lens,
                                 const int32_t* prefill_prefix_lens, float softmax_scale, int num_prefill_seqs,
                                 int max_prefill_len, int64_t num_prefill_tokens, int cur_layer, int num_layers,
                                 int num_q_heads, int num_kv_heads, int block_size, int head_dim,
                                 int max_blocks_per_seq, int64_t num_blocks, int64_t q_row_stride, sllm_dtype_t dtype,
                                 sllm_stream_t stream);

/* ---- Block-table maintenance: swiftllm/worker/kernels/block_mgmt.py:26-46, :66-80, :106-127
 * block_table int32 [max_seqs, max_blocks_per_seq]; num_seq_allocated_blocks int32 [max_seqs];
 * is_block_free uint8/bool [num_blocks]; candidate_blocks int64 [sum(block_needed)];
 * block_needed int32 [batch]; block_needed_cumsum int32 [batch] (inclusive). */
int sllm_set_block_table_and_num_seq_alloc_blocks(int32_t* num_seq_allocated_blocks, int32_t* block_table,
                                                  const int64_t* candidate_blocks, const int32_t* seq_ids,
                                                  const int32_t* block_needed, const int32_t* block_needed_cumsum,
                                                  int batch_size, int max_blocks_per_seq, sllm_stream_t stream);
int sllm_unset_block_table_and_num_seq_alloc_blocks(int32_t* num_seq_allocated_blocks, const int32_t* block_table,
                                                    const int32_t* seq_ids, uint8_t* is_block_free, int batch_size,
                                                    int max_blocks_per_seq, sllm_stream_t stream);
/* allocated_cumsum int32 [batch] (inclusive cumsum of num_seq_allocated_blocks[seq_ids]);
 * gathered_block_ids int32 [allocated_cumsum[batch-1]] */
int sllm_gather_allocated_blocks_and_unset(int32_t* num_seq_allocated_blocks, const int32_t* block_table,
                                           const int32_t* seq_ids, uint8_t* is_block_free,
                                           const int32_t* allocated_cumsum, int32_t* gathered_block_ids,
                                           int batch_size, int max_blocks_per_seq, sllm_stream_t stream);

/* ---- Sync-free block allocation (replaces swiftllm/worker/block_manager.py:43-79: assert + .item() +
 * torch.nonzero).  One launch: for seq i (batch order) needs max(0, cdiv(target_lens[i], block_size) -
 * num_seq_allocated_blocks[seq_ids[i]]) blocks, taken lowest-free-id first - the reference's order - and
 * appended to the block table.  Writes the new ids (int64, batch order) to new_blocks_out (capacity
 * new_blocks_capacity; may be NULL) and {num_new_blocks, error_flag} to status_out (int32[2], error_flag:
 * 1 = not enough free blocks, 2 = a sequence already holds more blocks than its target; on error nothing
 * is modified). */
int sllm_allocate_blocks_for_seqs(int32_t* num_seq_allocated_blocks, int32_t* block_table, uint8_t* is_block_free,
                                  const int32_t* seq_ids, const int32_t* target_lens, int batch_size,
                                  int block_size, int64_t num_blocks, int max_blocks_per_seq,
                                  int64_t* new_blocks_out, int64_t new_blocks_capacity, int32_t* status_out,
                                  sllm_stream_t stream);

/* ---- Tensor-parallel exchange fused with the op that follows it (new; the reference is single-GPU):
 *   residual <- h(sum over ranks of partial) + residual ;  x_out <- rmsnorm(residual) * weight
 * i.e. the all-reduce after o_proj / down_proj merged with fused_add_rmsnorm_inplace (rmsnorm.py:67-89) in one kernel that
 * reads every rank's partial [num_tokens, hidden] directly over NVLink peer memory, in rank order with fp32 accumulation
 * (bit-identical on all ranks).  host_peer_bufs / host_peer_flags: HOST arrays of `nranks` DEVICE pointers: the partial buffer
 * and the flag pad (uint32 [16 slots][8]) of every rank as mapped into this process (e.g. torch symmetric memory);
 * epoch_state: local device memory, 32 uint32, zeroed once; slot: which of the alternating buffers / flag rows is used.
 * weight == NULL skips the norm (x_out unused). */
int sllm_allreduce_add_rmsnorm(const void* const* host_peer_bufs, void* const* host_peer_flags, int rank, int nranks,
                               int slot, void* epoch_state, void* x_out, void* residual, const void* weight, float eps,
                               int64_t num_tokens, int hidden, sllm_dtype_t dtype, sllm_stream_t stream);

/* Two-shot variant for larger TP degrees: token row t is reduced, added to the residual and normalised by rank t % nranks
 * only, which then stores the normalised row into EVERY rank's x_out buffer (host_peer_xout: HOST array of `nranks` DEVICE
 * pointers, symmetric memory like the partial buffers; the result appears in this rank's own buffer when the call's kernel
 * has finished).  Remote traffic per rank 2*(N-1)/N*T*H instead of (N-1)*T*H bytes; two flag barriers (signal-pad rows
 * `slot` and 8 + `slot`, slot < 8).  `residual` is only maintained for the rows this rank owns; weight is required.
 * mc_buf / mc_xout: NVLS multicast addresses (DEVICE pointers) of the partial buffer of this slot and of x_out, or both NULL.
 * When given, the reduction is ONE multimem.ld_reduce per 16 bytes (the NVSwitch sums all ranks' copies, fp32 accumulation)
 * and the broadcast ONE multimem.st, instead of N peer loads / stores. */
int sllm_allreduce_add_rmsnorm_2shot(const void* const* host_peer_bufs, void* const* host_peer_xout,
                                     void* const* host_peer_flags, const void* mc_buf, void* mc_xout, int rank, int nranks,
                                     int slot, void* epoch_state, void* residual, const void* weight, float eps,
                                     int64_t num_tokens, int hidden, sllm_dtype_t dtype, sllm_stream_t stream);

/* Low-latency variant for decode-sized exchanges (csrc/allreduce_ll.cu): same result as the two-shot call (row t reduced, added
 * to the residual and normalised by rank t % nranks, residual maintained for owned rows only), but WITHOUT barriers: partial rows
 * are pushed to their owner and normalised rows to every rank as 16-byte lines {payload, epoch tag, payload, epoch tag} that the
 * receiver polls ("LL" protocol); the complete plain rows appear in x_out (local).  partial: this rank's [num_tokens, hidden]
 * GEMM output (local).  host_peer_rs_recv / host_peer_ag_recv: HOST arrays of `nranks` DEVICE pointers to every rank's receive
 * buffers OF THIS SLOT (symmetric memory, zeroed once): rs_recv [nranks][rows_per_rank][hidden], ag_recv [>= num_tokens][hidden],
 * 4 bytes per element.  mc_ag_recv: NVLS multicast address of ag_recv (one multimem.st per line instead of nranks-1 peer stores)
 * or NULL.  num_tokens <= nranks * rows_per_rank.  epoch_state as above (shared with the other exchange calls of the slot). */
int sllm_allreduce_add_rmsnorm_ll(const void* partial, void* const* host_peer_rs_recv, void* const* host_peer_ag_recv,
                                  void* mc_ag_recv, int rank, int nranks, int slot, void* epoch_state, void* x_out,
                                  void* residual, const void* weight, float eps, int64_t num_tokens, int hidden,
                                  int64_t rows_per_rank, sllm_dtype_t dtype, sllm_stream_t stream);

/* ---- Block swapping: csrc/src/block_swapping.cpp:22-85 (swiftllm_c.swap_blocks, csrc/src/entrypoints.cpp:5-7)
 * host_src_ids/host_dst_ids: HOST arrays of n block ids.  k_swap/v_swap: HOST memory (pinned or pageable),
 * k_cache/v_cache: device.  block_bytes = bytes of one block (all layers/heads) of k_cache.
 * Consecutive (src,dst) runs are coalesced into one copy each, as the reference does. */
int sllm_swap_blocks(const int64_t* host_src_ids, const int64_t* host_dst_ids, int64_t n, int is_swap_in,
                     void* k_cache, void* v_cache, void* host_k_swap, void* host_v_swap, int64_t block_bytes,
                     sllm_stream_t stream);

/* Sync-free variant (new): src_ids / dst_ids are DEVICE arrays (int64) - the gathered source ids and the
 * newly allocated target ids never travel to the host (the reference does `.tolist()` twice per swap, model.py:374-377) - and
 * one kernel moves every block between the cache and the swap space, which must be pinned, device-mapped host memory
 * (cudaHostAlloc / torch pin_memory).  Same data movement as sllm_swap_blocks. */
int sllm_swap_blocks_gathered(const int64_t* src_ids, const int64_t* dst_ids, int64_t n, int is_swap_in, void* k_cache,
                              void* v_cache, void* host_k_swap, void* host_v_swap, int64_t block_bytes,
                              sllm_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* SWIFTLLM_B200_H */
