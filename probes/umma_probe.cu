// Standalone probe (no torch): validates, on a real B200, every hardware contract the tcgen05/TMA paged-decode
// kernel relies on, in one run:
//   1. cuTensorMapEncodeTiled with the permuted-stride 4-D map that loads one 16-token KV page as four 1 KiB
//      SWIZZLE_128B atoms laid out [token-group][d-half][8 tokens][64 d]          (variant A, 1 TMA / page)
//      and the plain 2-D map (64 d x 16 tokens boxes, 2 TMAs / page)               (variant B)
//   2. tcgen05.mma kind::f16, A = K tile (128 tokens x 128 d, K-major SW128), B = Q (16 heads x 128 d, K-major
//      SW128 written by threads)                       -> S^T[token, head] in TMEM
//   3. tcgen05.mma, A = V tile as MN-major SW128 (M = d, K = tokens), B = P (16 heads x 128 tokens, K-major SW128
//      written by threads)                             -> O^T[d, head] in TMEM; tries both (LBO,SBO) conventions
//   4. tcgen05.ld 32x32b readback, tcgen05.commit -> mbarrier, alloc/dealloc.
// Prints max-abs errors vs a CPU reference for each variant.  Build: nvcc -gencode arch=compute_100a,code=sm_100a
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, int n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(n)); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
    asm volatile(
        "{\n .reg .pred p;\n WAIT_LOOP:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra DONE;\n bra WAIT_LOOP;\n DONE:\n}\n" ::"r"(smem_u32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;      // version = 1 (Blackwell)
    d |= (uint64_t)2 << 61;      // SWIZZLE_128B
    return d;
}
// kind::f16 instruction descriptor: fp32 accumulate, bf16 inputs
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, int a_mn_major, int b_mn_major) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n"
                 ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

struct Variant { int k_sbo, k_half, v_lbo, v_sbo, page_stride, tg_stride, half_stride; };

// smem byte offset of the 1 KiB atom holding (page p, d-half h, token-group tg) of a 128-token tile
__device__ __forceinline__ int atom_off(const Variant& v, int p, int h, int tg) { return p * v.page_stride + tg * v.tg_stride + h * v.half_stride; }

constexpr int TILE_TOK = 128, D = 128, NH = 16;

__global__ void __launch_bounds__(128, 1) probe_kernel(const __grid_constant__ CUtensorMap kmap4, const __grid_constant__ CUtensorMap vmap4,
                                                       const __grid_constant__ CUtensorMap kmap2, const __grid_constant__ CUtensorMap vmap2,
                                                       int use4d, const __nv_bfloat16* q /*[NH][D]*/, const __nv_bfloat16* p /*[NH][TILE_TOK]*/,
                                                       float* s_out /*[TILE_TOK][NH]*/, float* o_out /*[D][NH]*/, int swap_lbo_sbo,
                                                       uint8_t* smem_dump) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* k_sm = smem;                 // 32 KiB
    uint8_t* v_sm = smem + 32768;         // 32 KiB
    uint8_t* q_sm = smem + 65536;         // 2 halves x 16 rows x 128 B = 4 KiB
    uint8_t* p_sm = smem + 65536 + 4096;  // 4 KiB
    __shared__ uint64_t bar_load, bar_mma;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    Variant var;
    if (use4d) var = Variant{2048, 1024, 1024, 2048, 4096, 2048, 1024};
    else       var = Variant{1024, 16384, 16384, 1024, 2048, 1024, 16384};

    if (tid == 0) {
        mbar_init(&bar_load, 1); mbar_init(&bar_mma, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"(smem_u32(&tmem_base_s)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    // Q and P into K-major SW128 B-operand layout: [half][16 rows][128 B], chunk ^= row & 7
    for (int i = tid; i < NH * D; i += 128) {
        const int h = i / D, d = i % D, half = d / 64, c = (d % 64) / 8;
        *reinterpret_cast<__nv_bfloat16*>(q_sm + half * 2048 + h * 128 + ((c ^ (h & 7)) << 4) + (d % 8) * 2) = q[i];
        const int t = d;   // same index pattern for P: [head][token]
        *reinterpret_cast<__nv_bfloat16*>(p_sm + half * 2048 + h * 128 + ((c ^ (h & 7)) << 4) + (t % 8) * 2) = p[i];
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy smem writes -> visible to tcgen05/TMA
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_s;

    if (tid == 0) {
        mbar_expect_tx(&bar_load, 65536);
        for (int pg = 0; pg < 8; pg++) {
            if (use4d) {
                tma_load_4d(k_sm + atom_off(var, pg, 0, 0), &kmap4, &bar_load, 0, 0, 0, pg * 2);
                tma_load_4d(v_sm + atom_off(var, pg, 0, 0), &vmap4, &bar_load, 0, 0, 0, pg * 2);
            } else {
                for (int h = 0; h < 2; h++) {
                    tma_load_2d(k_sm + atom_off(var, pg, h, 0), &kmap2, &bar_load, h * 64, pg * 16);
                    tma_load_2d(v_sm + atom_off(var, pg, h, 0), &vmap2, &bar_load, h * 64, pg * 16);
                }
            }
        }
    }
    mbar_wait(&bar_load, 0);
    // dump raw smem K tile for host inspection
    for (int i = tid; i < 32768 / 16; i += 128) reinterpret_cast<uint4*>(smem_dump)[i] = reinterpret_cast<uint4*>(k_sm)[i];
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

    if (tid == 0) {
        // S^T (128 x 16) = K (A, K-major) x Q^T (B, K-major): columns [0,16)
        const uint32_t idesc_s = make_idesc(128, NH, 0, 0);
        for (int ks = 0; ks < 8; ks++) {
            const uint64_t a = make_desc(smem_u32(k_sm) + (ks / 4) * var.k_half + (ks % 4) * 32, 16, var.k_sbo);
            const uint64_t b = make_desc(smem_u32(q_sm) + (ks / 4) * 2048 + (ks % 4) * 32, 16, 1024);
            umma(tmem + 0, a, b, idesc_s, ks > 0);
        }
        // O^T (128 x 16) = V^T (A, MN-major) x P^T (B, K-major): columns [16,32)
        const uint32_t idesc_o = make_idesc(128, NH, 1, 0);
        for (int kt = 0; kt < 8; kt++) {       // 16 tokens per instruction = page kt
            const uint32_t lbo = swap_lbo_sbo ? var.v_sbo : var.v_lbo, sbo = swap_lbo_sbo ? var.v_lbo : var.v_sbo;
            const uint64_t a = make_desc(smem_u32(v_sm) + atom_off(var, kt, 0, 0), lbo, sbo);
            const uint64_t b = make_desc(smem_u32(p_sm) + (kt / 4) * 2048 + (kt % 4) * 32, 16, 1024);
            umma(tmem + 16, a, b, idesc_o, kt > 0);
        }
        umma_commit(&bar_mma);
    }
    mbar_wait(&bar_mma, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // read back: thread t of warp w owns TMEM lane 32w + t
    uint32_t r[32];
    const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                   "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
                   "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
                   "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < NH; j++) {
        s_out[tid * NH + j] = __uint_as_float(r[j]);
        o_out[tid * NH + j] = __uint_as_float(r[16 + j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(tmem));
}

static float bf(__nv_bfloat16 x) { return __bfloat162float(x); }

int main() {
    EncodeFn encode = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&encode, cudaEnableDefault, &qres));
    if (!encode) { printf("no cuTensorMapEncodeTiled\n"); return 1; }

    // a small paged cache: [num_blocks=12][L=2][nkv=2][bs=16][D=128]; we gather 8 pages of (layer 1, head 1)
    const int NB = 12, L = 2, NKV = 2, BS = 16;
    const size_t rows = (size_t)NB * L * NKV * BS;
    std::vector<__nv_bfloat16> hk(rows * D), hv(rows * D), hq(NH * D), hp(NH * TILE_TOK);
    srand(1);
    auto rnd = []() { return (float)(rand() % 2001 - 1000) / 1000.f; };
    for (auto& x : hk) x = __float2bfloat16(rnd());
    for (auto& x : hv) x = __float2bfloat16(rnd());
    for (auto& x : hq) x = __float2bfloat16(rnd());
    for (auto& x : hp) x = __float2bfloat16(fabsf(rnd()));
    // to keep the probe simple the 8 pages are consecutive blocks 2..9 => their rows are NOT contiguous (other layers/heads
    // in between), we pass a base pointer at (block 2, layer 1, head 1) and use the page stride via coordinates.
    __nv_bfloat16 *dk, *dv, *dq, *dp; float *ds, *dox; uint8_t* ddump;
    CK(cudaMalloc(&dk, hk.size() * 2)); CK(cudaMalloc(&dv, hv.size() * 2)); CK(cudaMalloc(&dq, hq.size() * 2)); CK(cudaMalloc(&dp, hp.size() * 2));
    CK(cudaMalloc(&ds, TILE_TOK * NH * 4)); CK(cudaMalloc(&dox, D * NH * 4)); CK(cudaMalloc(&ddump, 32768));
    CK(cudaMemcpy(dk, hk.data(), hk.size() * 2, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dv, hv.data(), hv.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dq, hq.data(), hq.size() * 2, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dp, hp.data(), hp.size() * 2, cudaMemcpyHostToDevice));

    // For the probe the tile = 8 consecutive pages of ONE (block-major) stream: use a contiguous region of 128 rows
    // starting at row r0 (pages are then physically consecutive; the gather logic itself is host-side index math).
    const size_t r0 = 5 * 16;
    CUtensorMap kmap4, vmap4, kmap2, vmap2;
    {
        // 4-D: (d_in 64 | tok_in 8 | half 2 | group G); strides: 256 B, 128 B, 2048 B
        cuuint64_t dims[4] = {64, 8, 2, (cuuint64_t)(rows / 8)};
        cuuint64_t strides[3] = {256, 128, 2048};
        cuuint32_t box[4] = {64, 8, 2, 2}, estr[4] = {1, 1, 1, 1};
        CUresult r1 = encode(&kmap4, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, dk + r0 * D, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        CUresult r2 = encode(&vmap4, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, dv + r0 * D, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("encode 4D permuted-stride maps: %d %d (0 = CUDA_SUCCESS)\n", (int)r1, (int)r2);
        cuuint64_t dims2[2] = {128, (cuuint64_t)rows};
        cuuint64_t strides2[1] = {256};
        cuuint32_t box2[2] = {64, 16}, estr2[2] = {1, 1};
        CUresult r3 = encode(&kmap2, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dk + r0 * D, dims2, strides2, box2, estr2, CU_TENSOR_MAP_INTERLEAVE_NONE,
                             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        CUresult r4 = encode(&vmap2, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dv + r0 * D, dims2, strides2, box2, estr2, CU_TENSOR_MAP_INTERLEAVE_NONE,
                             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("encode 2D maps: %d %d\n", (int)r3, (int)r4);
    }
    // CPU reference
    std::vector<float> s_ref(TILE_TOK * NH), o_ref(D * NH);
    for (int t = 0; t < TILE_TOK; t++) for (int h = 0; h < NH; h++) {
        float a = 0; for (int d = 0; d < D; d++) a += bf(hk[(r0 + t) * D + d]) * bf(hq[h * D + d]);
        s_ref[t * NH + h] = a;
    }
    for (int d = 0; d < D; d++) for (int h = 0; h < NH; h++) {
        float a = 0; for (int t = 0; t < TILE_TOK; t++) a += bf(hv[(r0 + t) * D + d]) * bf(hp[h * TILE_TOK + t]);
        o_ref[d * NH + h] = a;
    }
    const int smem_bytes = 65536 + 8192 + 1024;
    CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    for (int use4d = 1; use4d >= 0; use4d--) for (int swap = 0; swap < 2; swap++) {
        CK(cudaMemset(ds, 0xff, TILE_TOK * NH * 4)); CK(cudaMemset(dox, 0xff, D * NH * 4));
        probe_kernel<<<1, 128, smem_bytes>>>(kmap4, vmap4, kmap2, vmap2, use4d, dq, dp, ds, dox, swap, ddump);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("variant use4d=%d swap=%d: kernel error %s\n", use4d, swap, cudaGetErrorString(e)); return 2; }
        std::vector<float> s(TILE_TOK * NH), o(D * NH);
        CK(cudaMemcpy(s.data(), ds, s.size() * 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(o.data(), dox, o.size() * 4, cudaMemcpyDeviceToHost));
        float es = 0, eo = 0, ms = 0, mo = 0;
        for (size_t i = 0; i < s.size(); i++) { es = fmaxf(es, fabsf(s[i] - s_ref[i])); ms = fmaxf(ms, fabsf(s_ref[i])); }
        for (size_t i = 0; i < o.size(); i++) { eo = fmaxf(eo, fabsf(o[i] - o_ref[i])); mo = fmaxf(mo, fabsf(o_ref[i])); }
        printf("variant tma=%s vdesc=%s : S^T max err %.4g (max |ref| %.3g)  O^T max err %.4g (max |ref| %.3g)  -> %s %s\n",
               use4d ? "4D-1perpage" : "2D-2perpage", swap ? "swapped(LBO<->SBO)" : "as-documented", es, ms, eo, mo,
               es < 1e-2 * ms ? "S_OK" : "S_BAD", eo < 1e-2 * mo ? "O_OK" : "O_BAD");
        if (swap == 0) {
            // check raw smem image of the K tile against the expected atom layout
            std::vector<__nv_bfloat16> dump(16384);
            CK(cudaMemcpy(dump.data(), ddump, 32768, cudaMemcpyDeviceToHost));
            int bad = 0;
            for (int t = 0; t < TILE_TOK && bad < 5; t++) for (int d = 0; d < D; d++) {
                const int p = t / 16, tg = (t % 16) / 8, r = t % 8, h = d / 64, c = (d % 64) / 8;
                const int off = use4d ? p * 4096 + tg * 2048 + h * 1024 : h * 16384 + p * 2048 + tg * 1024;
                const int byte = off + r * 128 + ((c ^ r) << 4) + (d % 8) * 2;
                if (bf(dump[byte / 2]) != bf(hk[(r0 + t) * D + d])) { bad++; if (bad < 4) printf("  smem mismatch t=%d d=%d\n", t, d); }
            }
            printf("  K tile smem image (%s): %s\n", use4d ? "4D" : "2D", bad ? "MISMATCH" : "matches expected atom layout");
        }
    }
    return 0;
}
