// KV-cache page store and block-table maintenance (exact integer / byte work, HBM- or latency-bound).
// Reference: swiftllm/worker/kernels/kvcache_mgmt.py, swiftllm/worker/kernels/block_mgmt.py,
//            swiftllm/worker/block_manager.py:43-79 (allocation order).
// Cache layout (part of the contract, model.py:138-148): [num_blocks, num_layers, nkv, block_size, D].
#include "common.cuh"

namespace sllm {

// ------------------------------------------------------------------ store_kvcache, prefill part
// kvcache_mgmt.py:10-48.  grid (cdiv(max_prefill_len, bs), num_prefill_seqs); one CTA moves one page worth of
// tokens: source rows [tok][head][D] -> page [head][tok][D].  16-byte vectors, both sides coalesced per
// D-row.  Algorithmic bytes: 4 * T * nkv * D * sizeof(T) (K and V, read + write).
// PREFIX (chunked prefill, not in the reference): token t of chunk b lands at position prefix_lens[b] + t of its
// sequence, so the first page of a chunk may be entered in the middle; grid.x = cdiv(max_chunk_len, bs) + 1.
template <typename T, bool PREFIX>
__global__ void __launch_bounds__(256) store_kv_prefill_kernel(
    const T* __restrict__ k, const T* __restrict__ v, T* __restrict__ k_cache, T* __restrict__ v_cache,
    const int32_t* __restrict__ block_table, const int32_t* __restrict__ seq_ids,
    const int32_t* __restrict__ start_locs, const int32_t* __restrict__ seq_lens, int cur_layer, int num_layers,
    int nkv, int bs, int D, int max_blocks_per_seq, int64_t k_stride, int64_t v_stride,
    const int32_t* __restrict__ prefix_lens) {
    const int b = blockIdx.y;
    const int len = seq_lens[b];
    int pb = blockIdx.x;          // page of the sequence this CTA fills
    int off0 = 0;                 // first slot of the page that is written
    int ntok;
    int64_t row0;                 // source row of the first token written
    if constexpr (PREFIX) {
        const int pre = prefix_lens[b];
        pb += pre / bs;
        const int lo = max(pb * bs, pre), hi = min(pb * bs + bs, pre + len);
        if (lo >= hi) return;
        off0 = lo - pb * bs;
        ntok = hi - lo;
        row0 = start_locs[b] + (lo - pre);
    } else {
        const int tok0 = pb * bs;
        if (tok0 >= len) return;
        ntok = min(bs, len - tok0);
        row0 = start_locs[b] + tok0;
    }
    const int64_t blk = block_table[(int64_t)seq_ids[b] * max_blocks_per_seq + pb];
    const int64_t dst0 = (blk * num_layers + cur_layer) * (int64_t)nkv * bs * D + (int64_t)off0 * D;
    const int cpr = D >> 3;                     // 16-byte chunks per D-row
    const int items = ntok * nkv * cpr;
    for (int i = threadIdx.x; i < items; i += blockDim.x) {
        const int c = i % cpr, h = (i / cpr) % nkv, t = i / (cpr * nkv);
        const int64_t so = (int64_t)h * D + 8 * c;
        const int64_t d = dst0 + ((int64_t)h * bs + t) * D + 8 * c;
        st_vec8(k_cache + d, ld_vec8_stream(k + (row0 + t) * k_stride + so));
        st_vec8(v_cache + d, ld_vec8_stream(v + (row0 + t) * v_stride + so));
    }
}

// ------------------------------------------------------------------ store_kvcache, decode part
// kvcache_mgmt.py:50-79.  grid (num_decoding_seqs); new row goes to position len-1.
template <typename T>
__global__ void __launch_bounds__(128) store_kv_decode_kernel(
    const T* __restrict__ k, const T* __restrict__ v, T* __restrict__ k_cache, T* __restrict__ v_cache,
    const int32_t* __restrict__ block_table, const int32_t* __restrict__ seq_ids,
    const int32_t* __restrict__ seq_lens, int cur_layer, int num_layers, int nkv, int bs, int D,
    int max_blocks_per_seq, int64_t k_stride, int64_t v_stride) {
    const int b = blockIdx.x;
    const int pos = seq_lens[b] - 1;
    const int64_t blk = block_table[(int64_t)seq_ids[b] * max_blocks_per_seq + pos / bs];
    const int off = pos % bs;
    const int64_t dst0 = (blk * num_layers + cur_layer) * (int64_t)nkv * bs * D + (int64_t)off * D;
    const int cpr = D >> 3;
    for (int i = threadIdx.x; i < nkv * cpr; i += blockDim.x) {
        const int c = i % cpr, h = i / cpr;
        const int64_t so = (int64_t)h * D + 8 * c;
        const int64_t d = dst0 + (int64_t)h * bs * D + 8 * c;
        st_vec8(k_cache + d, ld_vec8(k + (int64_t)b * k_stride + so));
        st_vec8(v_cache + d, ld_vec8(v + (int64_t)b * v_stride + so));
    }
}

// ------------------------------------------------------------------ block_mgmt.py kernels (one warp per sequence)
__global__ void set_block_table_kernel(int32_t* nsab, int32_t* block_table, const int64_t* cand, const int32_t* seq_ids,
                                       const int32_t* need, const int32_t* need_cumsum, int batch, int mbps) {
    const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (b >= batch) return;
    const int lane = threadIdx.x & 31;
    const int64_t sid = seq_ids[b];
    const int n = need[b];
    const int start = need_cumsum[b] - n;
    const int have = nsab[sid];
    for (int i = lane; i < n; i += 32) block_table[sid * mbps + have + i] = (int32_t)cand[start + i];
    __syncwarp();
    if (lane == 0) nsab[sid] = have + n;
}

template <bool GATHER>
__global__ void unset_block_table_kernel(int32_t* nsab, const int32_t* block_table, const int32_t* seq_ids,
                                         uint8_t* is_free, const int32_t* cumsum, int32_t* gathered, int batch,
                                         int mbps) {
    const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (b >= batch) return;
    const int lane = threadIdx.x & 31;
    const int64_t sid = seq_ids[b];
    const int n = nsab[sid];
    const int out0 = GATHER ? (b > 0 ? cumsum[b - 1] : 0) : 0;
    for (int i = lane; i < n; i += 32) {
        const int32_t id = block_table[sid * mbps + i];
        if (GATHER) gathered[out0 + i] = id;
        is_free[id] = 1;
    }
    __syncwarp();
    if (lane == 0) nsab[sid] = 0;
}

// ------------------------------------------------------------------ sync-free allocation
// Replaces block_manager.py:43-79 (assert(...).all() + sum().item() + torch.nonzero()[:n]) by one launch of a
// single CTA.  Ordering contract (bit-exact block ids): the n lowest-numbered free blocks, ascending, are
// handed to the sequences in batch order - exactly what `torch.nonzero(is_block_free)[:n]` followed by
// block_mgmt.py:19-23 produces.
__device__ __forceinline__ int block_exclusive_scan(int v, int* warp_sums, int& block_total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int n = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += n;
    }
    __syncthreads();                      // protect warp_sums from the previous use
    if (lane == 31) warp_sums[warp] = inc;
    __syncthreads();
    int base = 0, tot = 0;
    for (int w = 0; w < nwarps; w++) {
        int s = warp_sums[w];
        if (w < warp) base += s;
        tot += s;
    }
    block_total = tot;
    return base + inc - v;
}

__global__ void __launch_bounds__(1024) allocate_blocks_kernel(
    int32_t* __restrict__ nsab, int32_t* __restrict__ block_table, uint8_t* __restrict__ is_free,
    const int32_t* __restrict__ seq_ids, const int32_t* __restrict__ target_lens, int batch, int bs,
    int64_t num_blocks, int mbps, int64_t* __restrict__ new_out, int64_t cap, int32_t* __restrict__ status) {
    extern __shared__ int32_t sm[];       // start[batch+1] exclusive scan of need, have[batch]
    int32_t* start = sm;
    int32_t* have = sm + batch + 1;
    __shared__ int warp_sums[32];
    __shared__ int s_err;
    if (threadIdx.x == 0) s_err = 0;
    __syncthreads();

    // 1. per-sequence need and its exclusive scan (batch order)
    int carry = 0;
    for (int b0 = 0; b0 < batch; b0 += blockDim.x) {
        const int b = b0 + threadIdx.x;
        int need = 0;
        if (b < batch) {
            const int h = nsab[seq_ids[b]];
            const int tgt = (target_lens[b] + bs - 1) / bs;
            if (h > tgt) atomicOr(&s_err, 2);
            if (tgt > mbps) atomicOr(&s_err, 4);
            need = max(0, tgt - h);
            have[b] = h;
        }
        int tot;
        const int ex = block_exclusive_scan(need, warp_sums, tot);
        if (b < batch) start[b] = carry + ex;
        carry += tot;
    }
    const int total = carry;
    if (threadIdx.x == 0) start[batch] = total;
    __syncthreads();

    // 2. count the free blocks (nothing may be modified if there are too few)
    const int64_t per_iter = (int64_t)blockDim.x * 16;
    int free_cnt = 0;
    if (s_err == 0 && total > 0) {
        int mine = 0;
        for (int64_t base = 0; base < num_blocks; base += per_iter) {
            const int64_t p = base + (int64_t)threadIdx.x * 16;
            if (p + 16 <= num_blocks && ((reinterpret_cast<uintptr_t>(is_free + p) & 15) == 0)) {
                const uint4 u = *reinterpret_cast<const uint4*>(is_free + p);
                const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
                for (int j = 0; j < 4; j++) mine += __popc(__vcmpne4(w[j], 0u) & 0x01010101u);
            } else {
                for (int j = 0; j < 16 && p + j < num_blocks; j++) mine += is_free[p + j] != 0;
            }
        }
        int tot;
        block_exclusive_scan(mine, warp_sums, tot);
        free_cnt = tot;
        if (free_cnt < total && threadIdx.x == 0) s_err |= 1;
    }
    __syncthreads();
    const int err = s_err | ((new_out != nullptr && total > cap) ? 8 : 0);
    if (threadIdx.x == 0) { status[0] = err ? 0 : total; status[1] = err; }
    if (err != 0 || total == 0) return;

    // 3. hand out the `total` lowest free ids in ascending order
    int running = 0;
    for (int64_t base = 0; base < num_blocks && running < total; base += per_iter) {
        const int64_t p = base + (int64_t)threadIdx.x * 16;
        uint8_t f[16];
        int mine = 0;
#pragma unroll
        for (int j = 0; j < 16; j++) {
            f[j] = (p + j < num_blocks) ? is_free[p + j] : 0;
            mine += f[j] != 0;
        }
        int tot;
        int rank = running + block_exclusive_scan(mine, warp_sums, tot);
#pragma unroll
        for (int j = 0; j < 16; j++) {
            if (f[j] && rank < total) {
                // sequence that owns candidate #rank: last b with start[b] <= rank
                int lo = 0, hi = batch;             // invariant start[lo] <= rank < start[hi]
                while (hi - lo > 1) {
                    const int mid = (lo + hi) >> 1;
                    if (start[mid] <= rank) lo = mid; else hi = mid;
                }
                const int64_t sid = seq_ids[lo];
                block_table[sid * mbps + have[lo] + (rank - start[lo])] = (int32_t)(p + j);
                if (new_out) new_out[rank] = p + j;
                is_free[p + j] = 0;
            }
            rank += f[j] != 0;
        }
        running += tot;
    }
    __syncthreads();
    for (int b = threadIdx.x; b < batch; b += blockDim.x) nsab[seq_ids[b]] = have[b] + (start[b + 1] - start[b]);
}

}  // namespace sllm

using namespace sllm;

// prefix_lens == nullptr: the reference's contract (every prefill entry starts at position 0)
static int store_kvcache_impl(const void* k, const void* v, void* k_cache, void* v_cache, const int32_t* block_table,
                              const int32_t* seq_ids, const int32_t* prefill_seq_start_locs, const int32_t* prefill_seq_lens,
                              const int32_t* prefix_lens, const int32_t* decoding_seq_lens, int num_prefill_seqs,
                              int num_decoding_seqs, int64_t num_prefill_tokens, int max_prefill_len, int cur_layer,
                              int num_layers, int nkv, int block_size, int head_dim, int max_blocks_per_seq,
                              int64_t k_row_stride, int64_t v_row_stride, sllm_dtype_t dtype, sllm_stream_t stream_) {
    SLLM_REQUIRE(head_dim > 0 && head_dim % 8 == 0, "store_kvcache: head_dim (%d) must be a multiple of 8", head_dim);
    SLLM_REQUIRE(k_row_stride >= (int64_t)nkv * head_dim && v_row_stride >= (int64_t)nkv * head_dim && k_row_stride % 8 == 0 &&
                 v_row_stride % 8 == 0, "store_kvcache: bad row strides (%lld, %lld)", (long long)k_row_stride, (long long)v_row_stride);
    SLLM_REQUIRE(block_size > 0 && nkv > 0 && num_layers > 0 && cur_layer >= 0 && cur_layer < num_layers,
                 "store_kvcache: bad cache geometry (layer %d of %d)", cur_layer, num_layers);
    SLLM_REQUIRE(num_prefill_seqs >= 0 && num_decoding_seqs >= 0, "store_kvcache: negative batch");
    if (num_prefill_seqs + num_decoding_seqs == 0) return 0;
    SLLM_REQUIRE(k && v && k_cache && v_cache && block_table && seq_ids, "store_kvcache: null pointer");
    cudaStream_t stream = (cudaStream_t)stream_;
    if (num_prefill_seqs > 0 && max_prefill_len > 0) {
        SLLM_REQUIRE(prefill_seq_start_locs && prefill_seq_lens, "store_kvcache: null prefill metadata");
        if (prefix_lens) {
            dim3 grid(cdiv(max_prefill_len, block_size) + 1, num_prefill_seqs);
            SLLM_DISPATCH_DTYPE(dtype, (store_kv_prefill_kernel<T, true><<<grid, 256, 0, stream>>>(
                                           (const T*)k, (const T*)v, (T*)k_cache, (T*)v_cache, block_table, seq_ids,
                                           prefill_seq_start_locs, prefill_seq_lens, cur_layer, num_layers, nkv, block_size,
                                           head_dim, max_blocks_per_seq, k_row_stride, v_row_stride, prefix_lens)));
        } else {
            dim3 grid(cdiv(max_prefill_len, block_size), num_prefill_seqs);
            SLLM_DISPATCH_DTYPE(dtype, (store_kv_prefill_kernel<T, false><<<grid, 256, 0, stream>>>(
                                           (const T*)k, (const T*)v, (T*)k_cache, (T*)v_cache, block_table, seq_ids,
                                           prefill_seq_start_locs, prefill_seq_lens, cur_layer, num_layers, nkv, block_size,
                                           head_dim, max_blocks_per_seq, k_row_stride, v_row_stride, nullptr)));
        }
        int e = check_launch("store_kvcache(prefill)");
        if (e) return e;
    }
    if (num_decoding_seqs > 0) {
        SLLM_REQUIRE(decoding_seq_lens, "store_kvcache: null decoding_seq_lens");
        const size_t elem = 2;
        const char* kd = (const char*)k + (size_t)num_prefill_tokens * (size_t)k_row_stride * elem;
        const char* vd = (const char*)v + (size_t)num_prefill_tokens * (size_t)v_row_stride * elem;
        SLLM_DISPATCH_DTYPE(dtype, (store_kv_decode_kernel<T><<<num_decoding_seqs, 128, 0, stream>>>(
                                       (const T*)kd, (const T*)vd, (T*)k_cache, (T*)v_cache, block_table,
                                       seq_ids + num_prefill_seqs, decoding_seq_lens, cur_layer, num_layers, nkv,
                                       block_size, head_dim, max_blocks_per_seq, k_row_stride, v_row_stride)));
        return check_launch("store_kvcache(decode)");
    }
    return 0;
}

extern "C" {

int sllm_store_kvcache(const void* k, const void* v, void* k_cache, void* v_cache, const int32_t* block_table,
                       const int32_t* seq_ids, const int32_t* prefill_seq_start_locs, const int32_t* prefill_seq_lens,
                       const int32_t* decoding_seq_lens, int num_prefill_seqs, int num_decoding_seqs,
                       int64_t num_prefill_tokens, int max_prefill_len, int cur_layer, int num_layers, int nkv,
                       int block_size, int head_dim, int max_blocks_per_seq, int64_t k_row_stride, int64_t v_row_stride,
                       sllm_dtype_t dtype, sllm_stream_t stream) {
    return store_kvcache_impl(k, v, k_cache, v_cache, block_table, seq_ids, prefill_seq_start_locs, prefill_seq_lens, nullptr,
                              decoding_seq_lens, num_prefill_seqs, num_decoding_seqs, num_prefill_tokens, max_prefill_len,
                              cur_layer, num_layers, nkv, block_size, head_dim, max_blocks_per_seq, k_row_stride, v_row_stride,
                              dtype, stream);
}

int sllm_store_kvcache_chunked(const void* k, const void* v, void* k_cache, void* v_cache, const int32_t* block_table,
                               const int32_t* seq_ids, const int32_t* prefill_seq_start_locs, const int32_t* prefill_seq_lens,
                               const int32_t* prefill_prefix_lens, const int32_t* decoding_seq_lens, int num_prefill_seqs,
                               int num_decoding_seqs, int64_t num_prefill_tokens, int max_prefill_len, int cur_layer,
                               int num_layers, int nkv, int block_size, int head_dim, int max_blocks_per_seq,
                               int64_t k_row_stride, int64_t v_row_stride, sllm_dtype_t dtype, sllm_stream_t stream) {
    SLLM_REQUIRE(num_prefill_seqs == 0 || prefill_prefix_lens, "store_kvcache_chunked: null prefill_prefix_lens");
    return store_kvcache_impl(k, v, k_cache, v_cache, block_table, seq_ids, prefill_seq_start_locs, prefill_seq_lens,
                              prefill_prefix_lens, decoding_seq_lens, num_prefill_seqs, num_decoding_seqs, num_prefill_tokens,
                              max_prefill_len, cur_layer, num_layers, nkv, block_size, head_dim, max_blocks_per_seq,
                              k_row_stride, v_row_stride, dtype, stream);
}

int sllm_set_block_table_and_num_seq_alloc_blocks(int32_t* nsab, int32_t* block_table, const int64_t* cand,
                                                  const int32_t* seq_ids, const int32_t* need, const int32_t* need_cumsum,
                                                  int batch, int mbps, sllm_stream_t stream) {
    if (batch <= 0) return 0;
    SLLM_REQUIRE(nsab && block_table && seq_ids && need && need_cumsum, "set_block_table: null pointer");
    set_block_table_kernel<<<cdiv(batch, 4), 128, 0, (cudaStream_t)stream>>>(nsab, block_table, cand, seq_ids, need,
                                                                              need_cumsum, batch, mbps);
    return check_launch("set_block_table_and_num_seq_alloc_blocks");
}

int sllm_unset_block_table_and_num_seq_alloc_blocks(int32_t* nsab, const int32_t* block_table, const int32_t* seq_ids,
                                                    uint8_t* is_free, int batch, int mbps, sllm_stream_t stream) {
    if (batch <= 0) return 0;
    SLLM_REQUIRE(nsab && block_table && seq_ids && is_free, "unset_block_table: null pointer");
    unset_block_table_kernel<false><<<cdiv(batch, 4), 128, 0, (cudaStream_t)stream>>>(nsab, block_table, seq_ids, is_free,
                                                                                       nullptr, nullptr, batch, mbps);
    return check_launch("unset_block_table_and_num_seq_alloc_blocks");
}

int sllm_gather_allocated_blocks_and_unset(int32_t* nsab, const int32_t* block_table, const int32_t* seq_ids,
                                           uint8_t* is_free, const int32_t* cumsum, int32_t* gathered, int batch,
                                           int mbps, sllm_stream_t stream) {
    if (batch <= 0) return 0;
    SLLM_REQUIRE(nsab && block_table && seq_ids && is_free && cumsum, "gather_allocated_blocks: null pointer");
    unset_block_table_kernel<true><<<cdiv(batch, 4), 128, 0, (cudaStream_t)stream>>>(nsab, block_table, seq_ids, is_free,
                                                                                      cumsum, gathered, batch, mbps);
    return check_launch("gather_allocated_blocks_and_unset");
}

int sllm_allocate_blocks_for_seqs(int32_t* nsab, int32_t* block_table, uint8_t* is_free, const int32_t* seq_ids,
                                  const int32_t* target_lens, int batch, int block_size, int64_t num_blocks, int mbps,
                                  int64_t* new_out, int64_t cap, int32_t* status, sllm_stream_t stream) {
    SLLM_REQUIRE(batch >= 0 && block_size > 0 && num_blocks >= 0 && mbps > 0, "allocate_blocks: bad sizes");
    SLLM_REQUIRE(status, "allocate_blocks: status_out is required");
    SLLM_REQUIRE(batch == 0 || (nsab && block_table && is_free && seq_ids && target_lens), "allocate_blocks: null pointer");
    const size_t smem = (size_t)(2 * batch + 1) * sizeof(int32_t);
    SLLM_REQUIRE(smem <= 200 * 1024, "allocate_blocks: batch %d too large", batch);
    if (smem > 48 * 1024)
        cudaFuncSetAttribute(allocate_blocks_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    allocate_blocks_kernel<<<1, 1024, smem, (cudaStream_t)stream>>>(nsab, block_table, is_free, seq_ids, target_lens, batch,
                                                                   block_size, num_blocks, mbps, new_out, cap, status);
    return check_launch("allocate_blocks_for_seqs");
}

}  // extern "C"
