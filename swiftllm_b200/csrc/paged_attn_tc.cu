// Paged (decode) attention, generation 2: persistent, warp-specialised tcgen05 + TMA kernel (sm_100a).
//
// Reference being replaced: swiftllm/worker/kernels/paged_attn.py:9-108 (+ :111-149 via the merge kernel of
// paged_attn.cu when a sequence is split).  Shapes served: head_dim 128, block_size 16, GQA group <= 16, fp16/bf16.
//
// One persistent CTA per SM walks a static list of work items (sequence, kv head, split).  Per 128-token tile:
//   warp 0 (1 lane)  TMA producer: reads 8 block-table entries, issues one 4-D tensor-map load per 16-token page
//                    for K and one for V (4 KiB each, SWIZZLE_128B atoms [token-group][d-half][8 tok][64 d]) into a
//                    3-stage ring, completion on mbarriers; Q of the item is a 2-box TMA load.
//   warp 1 (1 lane)  MMA issuer:  S^T[128 tok x 16 heads] = K_tile (A, K-major) . Q^T (B, K-major)         (8 UMMAs)
//                                 O^T[128 d  x 16 heads] = V_tile^T (A, MN-major) . P^T (B, K-major)        (8 UMMAs)
//                    accumulators in TMEM (2 x 16 columns each, double buffered), completion via tcgen05.commit.
//   warps 2-5        softmax: thread = token row of S^T (tcgen05.ld), scale + tail mask, tile max across the 128
//                    threads, p = exp2(s - m) -> bf16/fp16 P^T tile in shared memory (B operand of the PV UMMA);
//                    the same threads then own one d-row of O^T and fold each tile's O into fp32 registers with the
//                    running-max correction (no TMEM read-modify-write), and write o / the split partial.
// Tokens are the MMA M dimension (the transposed problem), so no tensor work is spent on padded query rows and the
// softmax uses all 128 threads; the GQA group is the N dimension (padded to 16 by reading neighbouring heads whose
// columns are simply never read back).
// HBM roofline: algorithmic bytes = sum_i len_i * nkv * 128 * 2 * sizeof(T) + 2 * Bd * nq * 128 * sizeof(T).
#include <mutex>
#include <unordered_map>

#include "tc_helpers.cuh"

namespace sllm {

using namespace tc;

constexpr int TC_D = 128;
constexpr int TC_BS = 16;                 // tokens per page
constexpr int TC_TILE = 128;              // tokens per pipeline stage (8 pages)
constexpr int TC_KSTAGES = 3;             // K ring (freed as soon as S = K.Q^T of the tile has completed)
constexpr int TC_VSTAGES = 3;             // V ring (freed when O = V^T.P^T of the tile has completed)
constexpr int TC_THREADS = 224;           // warp 0: K+Q producer, warp 1: MMA issuer, warps 2..5: softmax, warp 6: V producer
constexpr int TC_NPAD = 16;               // UMMA N (heads, padded)
constexpr int TC_K_BYTES = TC_TILE * TC_D * 2;                 // 32 KiB per K (or V) tile
constexpr int TC_Q_BYTES = 2 * TC_NPAD * 128;                  // [half][16 rows][128 B] = 4 KiB
constexpr int TC_P_BYTES = 2 * TC_NPAD * 128;                  // [token half][16 rows][128 B] = 4 KiB
constexpr int TC_SMEM_BYTES = (TC_KSTAGES + TC_VSTAGES) * TC_K_BYTES + 2 * TC_Q_BYTES + 2 * TC_P_BYTES + 2048;
constexpr int TC_TMEM_COLS = 64;          // S: 2 x 16, O: 2 x 16

struct TcParams {
    const int32_t* block_table; const int32_t* seq_ids; const int32_t* seq_lens;
    void* o; float* part_o; float* part_lse;
    float scale_log2e;
    int split_tokens, num_splits, cur_layer, num_layers, nq, nkv, max_blocks_per_seq, num_seqs, num_items;
    // shared-memory placement of the 1 KiB atoms of a KV tile and the matching descriptor strides
    int page_stride, tg_stride, half_stride, k_sbo, v_lbo, v_sbo, tma_4d;
};

struct Barriers {       // all mbarriers of the CTA (shared memory)
    uint64_t k_full[TC_KSTAGES], k_empty[TC_KSTAGES];
    uint64_t v_full[TC_VSTAGES], v_empty[TC_VSTAGES];
    uint64_t q_full[2], q_empty[2];
    uint64_t s_full[2], s_empty[2];
    uint64_t p_full[2];
    uint64_t o_full[2], o_empty[2];
};

struct Item { int seq, kvh, split, split_start, split_len, ntiles; };

__device__ __forceinline__ bool get_item(const TcParams& p, int idx, Item& it) {
    it.split = idx % p.num_splits;
    it.kvh = (idx / p.num_splits) % p.nkv;
    it.seq = idx / (p.num_splits * p.nkv);
    // every caller is a converged warp: broadcasting the loaded length makes it (and everything derived from it)
    // provably warp-uniform for the compiler
    const int len = __shfl_sync(0xffffffffu, p.seq_lens[it.seq], 0);
    it.split_start = it.split * p.split_tokens;
    if (it.split_start >= len) return false;
    it.split_len = min(p.split_tokens, len - it.split_start);
    it.ntiles = (it.split_len + TC_TILE - 1) / TC_TILE;
    return true;
}

// Walks this CTA's work items tile by tile (every role runs the same deterministic schedule).
struct TileCursor {
    int idx;            // global item index (blockIdx.x + k * gridDim.x)
    Item it;
    int j;              // tile within the item
    uint32_t t, n;      // tile / item counters of this CTA
    bool valid;
    __device__ __forceinline__ void seek(const TcParams& p) {            // position on the next non-empty item
        valid = false;
        while (idx < p.num_items) {
            if (get_item(p, idx, it)) { valid = true; j = 0; return; }
            idx += gridDim.x;
        }
    }
    __device__ __forceinline__ void init(const TcParams& p) { idx = blockIdx.x; t = 0; n = 0; seek(p); }
    __device__ __forceinline__ void advance(const TcParams& p) {
        t++;
        if (++j == it.ntiles) { n++; idx += gridDim.x; seek(p); }
    }
};

// One producer warp streams either the K or the V tiles (plus Q when IS_K): 32 block-table entries are fetched per
// 4 tiles (prefetched one group ahead), lanes 0..7 each issue the TMA of one page.
template <bool IS_K>
__device__ __forceinline__ void producer_loop(const TcParams& p, const CUtensorMap* map, const CUtensorMap* qmap, uint8_t* ring,
                                              uint64_t* full, uint64_t* empty, int nstages, Barriers* bars, uint8_t* q_sm, int g, int lane) {
    uint32_t t = 0, n = 0;
    for (int idx = blockIdx.x; idx < p.num_items; idx += gridDim.x) {
        Item it;
        if (!get_item(p, idx, it)) continue;
        if (IS_K) {       // Q of this item: heads [kvh*g, +16) of sequence seq, 128 d as two 64-wide boxes
            const int qb = n & 1;
            if (lane == 0) {
                mbar_wait(smem_u32(&bars->q_empty[qb]), ((n >> 1) & 1) ^ 1);
                const uint32_t qbar = smem_u32(&bars->q_full[qb]);
                mbar_arrive_expect_tx(qbar, TC_Q_BYTES);
                // q viewed as (128 d | nq heads | Bd sequences): heads past nq are zero-filled by the TMA unit
                tma_load_3d(smem_u32(q_sm + qb * TC_Q_BYTES), qmap, qbar, 0, it.kvh * g, it.seq);
                tma_load_3d(smem_u32(q_sm + qb * TC_Q_BYTES + 2048), qmap, qbar, 64, it.kvh * g, it.seq);
            }
        }
        const int first_page = it.split_start / TC_BS;
        const int npages = (it.split_len + TC_BS - 1) / TC_BS;
        const int32_t* bt = p.block_table + (int64_t)p.seq_ids[it.seq] * p.max_blocks_per_seq + first_page;
        int next_blk = lane < npages ? bt[lane] : 0;                       // group 0
        for (int tile0 = 0; tile0 < it.ntiles; tile0 += 4) {
            const int my_blk = next_blk;
            const int npg = (tile0 + 4) * 8 + lane;                        // prefetch the next group's entries
            next_blk = (tile0 + 4 < it.ntiles && npg < npages) ? bt[npg] : 0;
            const int nt = min(4, it.ntiles - tile0);
            for (int j = 0; j < nt; j++, t++) {
                const int stage = t % nstages;
                const int valid = min(8, npages - (tile0 + j) * 8);
                const uint32_t bar = smem_u32(&full[stage]);
                if (lane == 0) {
                    mbar_wait(smem_u32(&empty[stage]), ((t / nstages) & 1) ^ 1);
                    mbar_arrive_expect_tx(bar, (uint32_t)valid * TC_BS * TC_D * 2);
                }
                __syncwarp();
                const int blk = __shfl_sync(0xffffffffu, my_blk, j * 8 + (lane & 7));
                if (lane < valid) {
                    // row of the cache viewed as [num_blocks*L*nkv*16 rows, 128]: first token row of this page
                    const int64_t row = (((int64_t)blk * p.num_layers + p.cur_layer) * p.nkv + it.kvh) * TC_BS;
                    const uint32_t dst = smem_u32(ring + stage * TC_K_BYTES) + lane * p.page_stride;
                    if (p.tma_4d) {
                        tma_load_4d(dst, map, bar, 0, 0, 0, (int)(row >> 3));
                    } else {
                        tma_load_2d(dst, map, bar, 0, (int)row);
                        tma_load_2d(dst + p.half_stride, map, bar, 64, (int)row);
                    }
                }
            }
        }
        n++;
    }
}

template <typename T, int G>      // G = number of S/O columns read back (GQA group padded to 4, 8 or 16)
__global__ void __launch_bounds__(TC_THREADS, 1) paged_attn_tc_kernel(const __grid_constant__ CUtensorMap kmap,
                                                                      const __grid_constant__ CUtensorMap vmap,
                                                                      const __grid_constant__ CUtensorMap qmap,
                                                                      const TcParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* k_sm = smem;                                            // [KSTAGES][32 KiB]
    uint8_t* v_sm = smem + TC_KSTAGES * TC_K_BYTES;                  // [VSTAGES][32 KiB]
    uint8_t* q_sm = v_sm + TC_VSTAGES * TC_K_BYTES;                  // [2][4 KiB]
    uint8_t* p_sm = q_sm + 2 * TC_Q_BYTES;                           // [2][4 KiB]
    uint8_t* misc = p_sm + 2 * TC_P_BYTES;                           // barriers, tmem base, reduction scratch
    Barriers* bars = reinterpret_cast<Barriers*>(misc);
    uint32_t* tmem_base_s = reinterpret_cast<uint32_t*>(misc + 256);
    float* red = reinterpret_cast<float*>(misc + 320);               // [2 parity][4 warps][16] + [4][16] sums

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = p.nq / p.nkv;

    // ---- one-time setup
    {   // zero the V ring (and K): page slots that a short tail tile never loads must hold finite data (P = 0 there)
        uint4* z = reinterpret_cast<uint4*>(smem);
        for (int i = tid; i < (TC_KSTAGES + TC_VSTAGES) * TC_K_BYTES / 16; i += TC_THREADS) z[i] = make_uint4(0, 0, 0, 0);
    }
    if (tid == 0) {
        for (int i = 0; i < TC_KSTAGES; i++) { mbar_init(smem_u32(&bars->k_full[i]), 1); mbar_init(smem_u32(&bars->k_empty[i]), 1); }
        for (int i = 0; i < TC_VSTAGES; i++) { mbar_init(smem_u32(&bars->v_full[i]), 1); mbar_init(smem_u32(&bars->v_empty[i]), 1); }
        for (int i = 0; i < 2; i++) {
            mbar_init(smem_u32(&bars->q_full[i]), 1); mbar_init(smem_u32(&bars->q_empty[i]), 1);
            mbar_init(smem_u32(&bars->s_full[i]), 1); mbar_init(smem_u32(&bars->s_empty[i]), 128);
            mbar_init(smem_u32(&bars->p_full[i]), 128);
            mbar_init(smem_u32(&bars->o_full[i]), 1); mbar_init(smem_u32(&bars->o_empty[i]), 128);
        }
        mbar_fence_init();
        tma_prefetch_desc(&kmap); tma_prefetch_desc(&vmap); tma_prefetch_desc(&qmap);
    }
    if (warp == 1) tmem_alloc<TC_TMEM_COLS>(smem_u32(tmem_base_s));
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_base_s;

    if (warp == 0) {
        producer_loop<true>(p, &kmap, &qmap, k_sm, bars->k_full, bars->k_empty, TC_KSTAGES, bars, q_sm, g, lane);
    } else if (warp == 6) {
        producer_loop<false>(p, &vmap, &qmap, v_sm, bars->v_full, bars->v_empty, TC_VSTAGES, bars, q_sm, g, lane);
    } else if (warp == 1) {
        // =========================================================== MMA issuer: two in-order queues, polled.  The whole warp
        // runs the loop converged (votes keep the control flow provably uniform) and one elected lane issues, so ptxas keeps
        // the UMMA descriptors in uniform registers instead of wrapping every UTCHMMA in a uniformisation loop.
        {
            constexpr uint32_t IDESC_S = make_instr_desc(128, TC_NPAD, UmmaFmt<T>::value, 0, 0);   // A K-major, B K-major
            constexpr uint32_t IDESC_O = make_instr_desc(128, TC_NPAD, UmmaFmt<T>::value, 1, 0);   // A MN-major (V^T)
            constexpr uint32_t FULL = 0xffffffffu;
            const uint32_t tm = __shfl_sync(FULL, tmem, 0);
            const uint64_t dk0 = make_smem_desc(smem_u32(k_sm), 16, p.k_sbo);
            const uint64_t dq0 = make_smem_desc(smem_u32(q_sm), 16, 1024);
            const uint64_t dv0 = make_smem_desc(smem_u32(v_sm), p.v_lbo, p.v_sbo);
            const uint64_t dp0 = make_smem_desc(smem_u32(p_sm), 16, 1024);
            TileCursor cs, cp;                  // next S tile / next PV tile
            cs.init(p); cp.init(p);
            uint32_t spins = 0;
            while (cp.valid) {
                bool progressed = false;
                if (cs.valid) {
                    const uint32_t t = cs.t;
                    const int stage = t % TC_KSTAGES, b = t & 1, qb = cs.n & 1;
                    const bool ready = (cs.j > 0 || mbar_test_wait(smem_u32(&bars->q_full[qb]), (cs.n >> 1) & 1)) &&
                                       mbar_test_wait(smem_u32(&bars->k_full[stage]), (t / TC_KSTAGES) & 1) &&
                                       mbar_test_wait(smem_u32(&bars->s_empty[b]), ((t >> 1) & 1) ^ 1);
                    if (__all_sync(FULL, ready)) {
                        tc_fence_after();
                        const bool last = cs.j == cs.it.ntiles - 1;
                        if (elect_one()) {
                            const uint64_t kd = dk0 + (uint64_t)((stage * TC_K_BYTES) >> 4);
                            const uint64_t qd = dq0 + (uint64_t)((qb * TC_Q_BYTES) >> 4);
#pragma unroll
                            for (int ks = 0; ks < 8; ks++)      // 16 d per UMMA; d-half = ks / 4
                                umma_ss(tm + b * 16, kd + (uint64_t)(((ks >> 2) * p.half_stride + (ks & 3) * 32) >> 4),
                                        qd + (uint64_t)(((ks >> 2) * 2048 + (ks & 3) * 32) >> 4), IDESC_S, ks > 0);
                            umma_commit(smem_u32(&bars->s_full[b]));
                            umma_commit(smem_u32(&bars->k_empty[stage]));                       // K tile consumed
                            if (last) umma_commit(smem_u32(&bars->q_empty[qb]));                // Q buffer reusable
                        }
                        __syncwarp();
                        cs.advance(p);
                        progressed = true;
                    }
                }
                if (cp.t < cs.t || !cs.valid) {
                    const uint32_t t = cp.t;
                    const int stage = t % TC_VSTAGES, b = t & 1;
                    const bool ready = mbar_test_wait(smem_u32(&bars->p_full[b]), (t >> 1) & 1) &&
                                       mbar_test_wait(smem_u32(&bars->v_full[stage]), (t / TC_VSTAGES) & 1) &&
                                       mbar_test_wait(smem_u32(&bars->o_empty[b]), ((t >> 1) & 1) ^ 1);
                    if (__all_sync(FULL, ready)) {
                        tc_fence_after();
                        if (elect_one()) {
                            const uint64_t vd = dv0 + (uint64_t)((stage * TC_K_BYTES) >> 4);
                            const uint64_t pd = dp0 + (uint64_t)((b * TC_P_BYTES) >> 4);
#pragma unroll
                            for (int kt = 0; kt < 8; kt++)      // 16 tokens (one page) per UMMA
                                umma_ss(tm + 32 + b * 16, vd + (uint64_t)((kt * p.page_stride) >> 4),
                                        pd + (uint64_t)(((kt >> 2) * 2048 + (kt & 3) * 32) >> 4), IDESC_O, kt > 0);
                            umma_commit(smem_u32(&bars->o_full[b]));
                            umma_commit(smem_u32(&bars->v_empty[stage]));
                        }
                        __syncwarp();
                        cp.advance(p);
                        progressed = true;
                    }
                }
                if (progressed) spins = 0;
                else if (++spins > (1u << 24)) {
                    if (lane == 0) printf("sllm: MMA issuer watchdog (block %d, S tile %u, PV tile %u)\n", blockIdx.x, cs.t, cp.t);
                    __trap();
                }
            }
        }
    } else {
        // =========================================================== softmax / accumulate warps (128 threads)
        const int quad = warp & 3;                     // TMEM lane quadrant this warp may access
        const int row = quad * 32 + lane;              // token row of S^T, d row of O^T
        const uint32_t tlane = (uint32_t)(quad * 32) << 16;
        const int sw = warp - 2;                       // 0..3 index for the reduction scratch
        uint32_t t = 0;
        T* const out = reinterpret_cast<T*>(p.o);
        for (int idx = blockIdx.x; idx < p.num_items; idx += gridDim.x) {
            Item it;
            if (!get_item(p, idx, it)) continue;
            float m_run[G], l_part[G], acc[G], alpha_prev[G];
#pragma unroll
            for (int h = 0; h < G; h++) { m_run[h] = -INFINITY; l_part[h] = 0.f; acc[h] = 0.f; alpha_prev[h] = 0.f; }

            auto fold_o = [&](uint32_t tt) {            // acc = acc * alpha(tt) + O_tile(tt)
                const int b = tt & 1;
                mbar_wait(smem_u32(&bars->o_full[b]), (tt >> 1) & 1);
                tc_fence_after();
                uint32_t r[G];
                tmem_ld<G>(tmem + tlane + 32 + b * 16, r);
                tmem_ld_wait();
                tc_fence_before();
                mbar_arrive(smem_u32(&bars->o_empty[b]));
#pragma unroll
                for (int h = 0; h < G; h++) acc[h] = acc[h] * alpha_prev[h] + __uint_as_float(r[h]);
            };

            for (int j = 0; j < it.ntiles; j++, t++) {
                const int b = t & 1;
                mbar_wait(smem_u32(&bars->s_full[b]), (t >> 1) & 1);
                tc_fence_after();
                uint32_t r[G];
                tmem_ld<G>(tmem + tlane + b * 16, r);
                tmem_ld_wait();
                tc_fence_before();
                mbar_arrive(smem_u32(&bars->s_empty[b]));

                const bool valid = j * TC_TILE + row < it.split_len;
                float x[G], wmax[G];
#pragma unroll
                for (int h = 0; h < G; h++) {
                    x[h] = valid ? __uint_as_float(r[h]) * p.scale_log2e : -INFINITY;
                    wmax[h] = warp_max(x[h]);
                }
                float* rp = red + (t & 1) * 64;
                if (lane == 0) {
#pragma unroll
                    for (int h = 0; h < G; h++) rp[sw * 16 + h] = wmax[h];
                }
                named_bar_sync(1, 128);
                float pj[G];
#pragma unroll
                for (int h = 0; h < G; h++) {
                    const float mt = fmaxf(fmaxf(rp[h], rp[16 + h]), fmaxf(rp[32 + h], rp[48 + h]));
                    const float mn = fmaxf(m_run[h], mt);                  // finite: every tile has a valid token
                    const float a = fast_exp2_tc(m_run[h] - mn);
                    pj[h] = fast_exp2_tc(x[h] - mn);
                    l_part[h] = l_part[h] * a + pj[h];
                    m_run[h] = mn;
                    x[h] = a;                                              // x[] now carries alpha(j) = exp2(m(j-1) - m(j))
                }
                // P^T tile (B operand of the PV UMMA): [token half][16 head rows][128 B], SWIZZLE_128B.
                // Buffer b was last read by PV(t-2), whose completion (o_full) these threads observed last iteration.
                {
                    uint8_t* pb = p_sm + b * TC_P_BYTES + (row >> 6) * 2048 + (row & 7) * 2;
                    const int c = (row & 63) >> 3;
#pragma unroll
                    for (int h = 0; h < G; h++)
                        *reinterpret_cast<T*>(pb + h * 128 + ((c ^ (h & 7)) << 4)) = Traits<T>::from_f(pj[h]);
                }
                fence_proxy_async();
                mbar_arrive(smem_u32(&bars->p_full[b]));
                if (j > 0) fold_o(t - 1);
#pragma unroll
                for (int h = 0; h < G; h++) alpha_prev[h] = x[h];
            }
            fold_o(t - 1);

            // ---- finalise the item: l = sum over the 128 token-threads, out[h][d=row] = acc / l
            float* rp = red + 128;                                         // [4 warps][16] scratch for the sums
            float lsum[G];
#pragma unroll
            for (int h = 0; h < G; h++) lsum[h] = warp_sum(l_part[h]);
            // (every thread passed >= 1 per-tile barrier since the previous item's reads of rp)
            if (lane == 0) {
#pragma unroll
                for (int h = 0; h < G; h++) rp[sw * 16 + h] = lsum[h];
            }
            named_bar_sync(1, 128);
#pragma unroll
            for (int h = 0; h < G; h++) {
                if (h < g) {
                    const float L = (rp[h] + rp[16 + h]) + (rp[32 + h] + rp[48 + h]);
                    const float val = acc[h] / L;
                    const int head = it.kvh * g + h;
                    if (p.num_splits == 1) {
                        out[((int64_t)it.seq * p.nq + head) * TC_D + row] = Traits<T>::from_f(val);
                    } else {
                        const int64_t pi = ((int64_t)it.seq * p.nq + head) * p.num_splits + it.split;
                        p.part_o[pi * TC_D + row] = val;
                        if (row == 0) p.part_lse[pi] = log2f(L) + m_run[h];
                    }
                }
            }
        }
    }

    // ---- teardown
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<TC_TMEM_COLS>(tmem);
}

}  // namespace sllm

namespace sllm {

// ------------------------------------------------------------------ host side
namespace {

struct MapKey {
    const void* ptr; uint64_t rows; int kind; uint64_t aux;
    bool operator==(const MapKey& o) const { return ptr == o.ptr && rows == o.rows && kind == o.kind && aux == o.aux; }
};
struct MapKeyHash {
    size_t operator()(const MapKey& k) const {
        return std::hash<const void*>()(k.ptr) ^ (std::hash<uint64_t>()(k.rows) * 1315423911u) ^ (size_t)k.kind ^ (std::hash<uint64_t>()(k.aux) << 1);
    }
};
std::mutex g_map_mutex;
std::unordered_map<MapKey, CUtensorMap, MapKeyHash> g_maps;
int g_kv_4d_ok = -1;            // -1 unknown, 0 the driver rejected the permuted-stride 4-D map, 1 usable (a property of the
                                // driver's tensor-map encoder, identical for every device of the process)

CUtensorMapDataType map_dtype(sllm_dtype_t dt) { return dt == SLLM_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16; }

// kind 0: 2-D [rows, 128] map with (64 x 16) boxes; kind 1: 4-D page map (64 d | 8 tok | 2 halves | rows/8), box (64,8,2,2);
// kind 2: q as (128 d | nq heads | rows sequences) with sequence stride aux>>16 elements and nq = aux & 0xffff, box (64,16,1)
bool encode_map(CUtensorMap* out, const void* ptr, uint64_t rows, int kind, sllm_dtype_t dt, uint64_t aux) {
    TensorMapEncodeFn enc = get_tensor_map_encoder();
    if (!enc) return false;
    CUresult r;
    if (kind == 0) {
        cuuint64_t dims[2] = {128, rows};
        cuuint64_t strides[1] = {256};
        cuuint32_t box[2] = {64, 16}, es[2] = {1, 1};
        r = enc(out, map_dtype(dt), 2, const_cast<void*>(ptr), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    } else if (kind == 2) {
        cuuint64_t dims[3] = {128, aux & 0xffff, rows};
        cuuint64_t strides[2] = {256, (aux >> 16) * 2};
        cuuint32_t box[3] = {64, 16, 1}, es[3] = {1, 1, 1};
        r = enc(out, map_dtype(dt), 3, const_cast<void*>(ptr), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    } else {
        cuuint64_t dims[4] = {64, 8, 2, rows / 8};
        cuuint64_t strides[3] = {256, 128, 2048};
        cuuint32_t box[4] = {64, 8, 2, 2}, es[4] = {1, 1, 1, 1};
        r = enc(out, map_dtype(dt), 4, const_cast<void*>(ptr), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    }
    return r == CUDA_SUCCESS;
}

bool get_map(CUtensorMap* out, const void* ptr, uint64_t rows, int kind, sllm_dtype_t dt, bool cache, uint64_t aux = 0) {
    MapKey key{ptr, rows, kind * 2 + (int)dt, aux};
    if (cache) {
        std::lock_guard<std::mutex> lk(g_map_mutex);
        auto it = g_maps.find(key);
        if (it != g_maps.end()) { *out = it->second; return true; }
    }
    if (!encode_map(out, ptr, rows, kind, dt, aux)) return false;
    if (cache) {
        std::lock_guard<std::mutex> lk(g_map_mutex);
        if (g_maps.size() > 4096) g_maps.clear();
        g_maps[key] = *out;
    }
    return true;
}

}  // namespace

bool tc_paged_supported(int head_dim, int block_size, int nq, int nkv, int64_t num_blocks, int num_layers) {
    if (head_dim != TC_D || block_size != TC_BS) return false;
    if (nkv <= 0 || nq % nkv != 0 || nq / nkv > TC_NPAD) return false;
    if ((int64_t)num_blocks * num_layers * nkv * TC_BS >= (1LL << 31)) return false;     // TMA coordinates are int32
    return get_tensor_map_encoder() != nullptr;
}

// Launches the persistent kernel; returns 0 / error like the other entry points.  `split_tokens`/`num_splits` come from
// the shared planner in paged_attn.cu.
int launch_paged_tc(const void* q, const void* k_cache, const void* v_cache, const int32_t* block_table, const int32_t* seq_ids,
                    const int32_t* seq_lens, void* o, float* part_o, float* part_lse, float scale_log2e, int num_seqs,
                    int split_tokens, int num_splits, int cur_layer, int num_layers, int nq, int nkv, int max_blocks_per_seq,
                    int64_t num_blocks, int64_t q_row_stride, sllm_dtype_t dtype, int num_sms, cudaStream_t stream) {
    const uint64_t rows = (uint64_t)num_blocks * num_layers * nkv * TC_BS;
    CUtensorMap kmap, vmap, qmap;
    if (g_kv_4d_ok != 0) {
        bool ok = get_map(&kmap, k_cache, rows, 1, dtype, true) && get_map(&vmap, v_cache, rows, 1, dtype, true);
        if (g_kv_4d_ok < 0) g_kv_4d_ok = ok ? 1 : 0;
    }
    if (g_kv_4d_ok == 0) {
        SLLM_REQUIRE(get_map(&kmap, k_cache, rows, 0, dtype, true) && get_map(&vmap, v_cache, rows, 0, dtype, true),
                     "paged_attention: cuTensorMapEncodeTiled failed for the KV cache");
    }
    // q is a transient activation (its address changes from call to call): encoded per launch (host-only, ~1 us), never cached
    SLLM_REQUIRE(get_map(&qmap, q, (uint64_t)num_seqs, 2, dtype, false, ((uint64_t)q_row_stride << 16) | (uint64_t)nq),
                 "paged_attention: cuTensorMapEncodeTiled failed for q");

    TcParams p;
    p.block_table = block_table; p.seq_ids = seq_ids; p.seq_lens = seq_lens; p.o = o; p.part_o = part_o; p.part_lse = part_lse;
    p.scale_log2e = scale_log2e; p.split_tokens = split_tokens; p.num_splits = num_splits; p.cur_layer = cur_layer;
    p.num_layers = num_layers; p.nq = nq; p.nkv = nkv; p.max_blocks_per_seq = max_blocks_per_seq; p.num_seqs = num_seqs;
    p.num_items = num_seqs * nkv * num_splits;
    p.tma_4d = g_kv_4d_ok == 1;
    if (p.tma_4d) { p.page_stride = 4096; p.tg_stride = 2048; p.half_stride = 1024; p.k_sbo = 2048; p.v_lbo = 1024; p.v_sbo = 2048; }
    else          { p.page_stride = 2048; p.tg_stride = 1024; p.half_stride = 16384; p.k_sbo = 1024; p.v_lbo = 16384; p.v_sbo = 1024; }
    const int grid = p.num_items < num_sms ? p.num_items : num_sms;
    const int g = nq / nkv;

#define SLLM_TC_LAUNCH(TT, GG)                                                                                          \
    do {                                                                                                                 \
        static unsigned long long configured = 0;                                                                        \
        if (first_use_on_this_device(configured))                                                                        \
            cudaFuncSetAttribute(paged_attn_tc_kernel<TT, GG>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES); \
        paged_attn_tc_kernel<TT, GG><<<grid, TC_THREADS, TC_SMEM_BYTES, stream>>>(kmap, vmap, qmap, p);                  \
    } while (0)
    if (dtype == SLLM_F16) {
        if (g <= 4) SLLM_TC_LAUNCH(__half, 4); else if (g <= 8) SLLM_TC_LAUNCH(__half, 8); else SLLM_TC_LAUNCH(__half, 16);
    } else {
        if (g <= 4) SLLM_TC_LAUNCH(__nv_bfloat16, 4); else if (g <= 8) SLLM_TC_LAUNCH(__nv_bfloat16, 8); else SLLM_TC_LAUNCH(__nv_bfloat16, 16);
    }
#undef SLLM_TC_LAUNCH
    return check_launch("paged_attention(tcgen05)");
}

}  // namespace sllm
