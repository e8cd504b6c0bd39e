// Microbenchmark: how fast can one SM-resident persistent CTA per SM stream a PAGED KV layout into shared memory?
// Same cache geometry as Llama-3-8B ([blocks][32 layers][8 heads][16 tok][128 d] bf16, 4 KiB pages 1 MiB apart),
// 3-stage ring of 64 KiB stages (8 K pages + 8 V pages), consumer = one thread that releases a stage as soon as it is
// full.  Variants: who issues (1 lane / 8 lanes / 4 warps x 1 lane) and how (4-D page box 4 KiB, 2-D half-page box 2 KiB,
// 1-D bulk 4 KiB, cp.async 16 B by 128 threads).  Prints GB/s per variant.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)
typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t b, int n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(b), "r"(n)); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(b) : "memory"); }
__device__ __forceinline__ bool mbar_try(uint32_t b, uint32_t par) {
    uint32_t ok; asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(b), "r"(par) : "memory"); return ok;
}
__device__ __forceinline__ void mbar_wait(uint32_t b, uint32_t par) { uint32_t n = 0; while (!mbar_try(b, par)) { if (++n > (1u << 26)) __trap(); } }
__device__ __forceinline__ void tma4(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                 ::"r"(dst), "l"(m), "r"(bar), "r"(0), "r"(0), "r"(0), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma2(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst), "l"(m), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void bulk1d(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void cp16(uint32_t dst, const void* src) { asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory"); }
__device__ __forceinline__ void cp_mbar_arrive(uint32_t bar) { asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory"); }

constexpr int L = 32, NKV = 8, STAGES = 3, STAGE_BYTES = 65536, PAGES_PER_SEQ = 256;

// mode 0: 1 lane, 4-D page boxes (16 TMAs / stage)        mode 1: 1 lane, 2-D half-page boxes (32 / stage)
// mode 2: 16 lanes of warp 0, one 4-D box each             mode 3: 1 lane, 1-D bulk 4 KiB per page
// mode 4: 4 producer warps x 8 lanes... (=32 lanes: 16 boxes over lanes 0..15 of warp 0 and warp 1 split K/V)
// mode 5: 128 threads cp.async 16 B (LDGSTS), mbarrier completion
__global__ void __launch_bounds__(192, 1) stream_kernel(const __grid_constant__ CUtensorMap kmap4, const __grid_constant__ CUtensorMap vmap4,
                                                        const __grid_constant__ CUtensorMap kmap2, const __grid_constant__ CUtensorMap vmap2,
                                                        const __nv_bfloat16* kc, const __nv_bfloat16* vc, const int* block_table, int num_seqs,
                                                        int layer, int mode, unsigned long long* sink) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t full[STAGES], empty[STAGES];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nprod = mode == 5 ? 128 : 1;
    if (tid == 0) {
        for (int i = 0; i < STAGES; i++) { mbar_init(smem_u32(&full[i]), mode == 5 ? 128 : (mode == 2 ? 16 : (mode == 4 ? 2 : 1))); mbar_init(smem_u32(&empty[i]), 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int num_items = num_seqs * NKV;
    if (warp == 5) {          // consumer: one thread
        if (lane == 0) {
            uint32_t t = 0; unsigned long long acc = 0;
            for (int item = blockIdx.x; item < num_items; item += gridDim.x)
                for (int tile = 0; tile < PAGES_PER_SEQ / 8; tile++, t++) {
                    const int s = t % STAGES;
                    mbar_wait(smem_u32(&full[s]), (t / STAGES) & 1);
                    acc += *reinterpret_cast<volatile uint32_t*>(smem + s * STAGE_BYTES + 64);
                    mbar_arrive(smem_u32(&empty[s]));
                }
            if (acc == 0x1234567) *sink = acc;
        }
        return;
    }
    if (mode == 5) {
        if (warp >= 4) return;
        uint32_t t = 0;
        for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
            const int seq = item / NKV, h = item % NKV;
            const int* bt = block_table + seq * PAGES_PER_SEQ;
            for (int tile = 0; tile < PAGES_PER_SEQ / 8; tile++, t++) {
                const int s = t % STAGES;
                mbar_wait(smem_u32(&empty[s]), ((t / STAGES) & 1) ^ 1);
                const uint32_t base = smem_u32(smem + s * STAGE_BYTES);
#pragma unroll
                for (int i = 0; i < 16; i++) {           // 2048 chunks of 16 B for K, same for V: 128 threads x 16
                    const int idx = tid + i * 128, pg = idx >> 8, within = idx & 255;
                    const int64_t off = (((int64_t)bt[tile * 8 + pg] * L + layer) * NKV + h) * 2048 + within * 8;
                    cp16(base + idx * 16, kc + off);
                    cp16(base + 32768 + idx * 16, vc + off);
                }
                cp_mbar_arrive(smem_u32(&full[s]));
            }
        }
        return;
    }
    // TMA modes: producers live in warps 0..1
    if (warp > 1) return;
    if (mode != 4 && warp == 1) return;
    uint32_t t = 0;
    for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
        const int seq = item / NKV, h = item % NKV;
        const int* bt = block_table + seq * PAGES_PER_SEQ;
        for (int tile = 0; tile < PAGES_PER_SEQ / 8; tile++, t++) {
            const int s = t % STAGES;
            const uint32_t bar = smem_u32(&full[s]);
            const uint32_t base = smem_u32(smem + s * STAGE_BYTES);
            if (mode == 0 || mode == 1 || mode == 3) {
                if (lane == 0) {
                    mbar_wait(smem_u32(&empty[s]), ((t / STAGES) & 1) ^ 1);
                    mbar_expect_tx(bar, STAGE_BYTES);
                    for (int pg = 0; pg < 8; pg++) {
                        const int64_t row = (((int64_t)bt[tile * 8 + pg] * L + layer) * NKV + h) * 16;
                        if (mode == 0) { tma4(base + pg * 4096, &kmap4, bar, (int)(row >> 3)); tma4(base + 32768 + pg * 4096, &vmap4, bar, (int)(row >> 3)); }
                        else if (mode == 1) {
                            tma2(base + pg * 2048, &kmap2, bar, 0, (int)row); tma2(base + 16384 + pg * 2048, &kmap2, bar, 64, (int)row);
                            tma2(base + 32768 + pg * 2048, &vmap2, bar, 0, (int)row); tma2(base + 49152 + pg * 2048, &vmap2, bar, 64, (int)row);
                        } else { bulk1d(base + pg * 4096, kc + row * 128, 4096, bar); bulk1d(base + 32768 + pg * 4096, vc + row * 128, 4096, bar); }
                    }
                }
                __syncwarp();
            } else if (mode == 2) {      // 16 lanes: lane l < 8 -> K page l, lane 8..15 -> V page l-8
                if (lane < 16) {
                    mbar_wait(smem_u32(&empty[s]), ((t / STAGES) & 1) ^ 1);
                    const int pg = lane & 7;
                    const int64_t row = (((int64_t)bt[tile * 8 + pg] * L + layer) * NKV + h) * 16;
                    mbar_expect_tx(bar, 4096);
                    if (lane < 8) tma4(base + pg * 4096, &kmap4, bar, (int)(row >> 3)); else tma4(base + 32768 + pg * 4096, &vmap4, bar, (int)(row >> 3));
                }
                __syncwarp();
            } else {                     // mode 4: warp 0 lane 0 -> all K pages, warp 1 lane 0 -> all V pages
                if (lane == 0) {
                    mbar_wait(smem_u32(&empty[s]), ((t / STAGES) & 1) ^ 1);
                    mbar_expect_tx(bar, 32768);
                    for (int pg = 0; pg < 8; pg++) {
                        const int64_t row = (((int64_t)bt[tile * 8 + pg] * L + layer) * NKV + h) * 16;
                        if (warp == 0) tma4(base + pg * 4096, &kmap4, bar, (int)(row >> 3)); else tma4(base + 32768 + pg * 4096, &vmap4, bar, (int)(row >> 3));
                    }
                }
                __syncwarp();
            }
        }
    }
}

int main() {
    EncodeFn encode = nullptr; cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&encode, cudaEnableDefault, &q));
    const int num_seqs = 96;                               // 96 x 8 heads x 1 MiB x (K+V) = 1.5 GiB per launch
    const int64_t num_blocks = (int64_t)num_seqs * PAGES_PER_SEQ;
    const int64_t elems = num_blocks * L * NKV * 16 * 128;  // 24 GiB per cache... too big: use L_alloc layers = 4
    (void)elems;
    // allocate a cache with the real block stride but only touch `layer` < L: needs the full 1 MiB block stride -> 24 GiB each
    __nv_bfloat16 *kc, *vc;
    const size_t bytes = (size_t)num_blocks * L * NKV * 4096;
    CK(cudaMalloc(&kc, bytes)); CK(cudaMalloc(&vc, bytes));
    CK(cudaMemset(kc, 1, bytes)); CK(cudaMemset(vc, 1, bytes));
    int* bt; CK(cudaMalloc(&bt, num_blocks * sizeof(int)));
    { int* h = (int*)malloc(num_blocks * sizeof(int)); for (int64_t i = 0; i < num_blocks; i++) h[i] = (int)((i * 7919) % num_blocks); CK(cudaMemcpy(bt, h, num_blocks * sizeof(int), cudaMemcpyHostToDevice)); free(h); }
    unsigned long long* sink; CK(cudaMalloc(&sink, 8));
    const uint64_t rows = (uint64_t)num_blocks * L * NKV * 16;
    CUtensorMap k4, v4, k2, v2;
    {
        cuuint64_t d4[4] = {64, 8, 2, rows / 8}; cuuint64_t s4[3] = {256, 128, 2048}; cuuint32_t b4[4] = {64, 8, 2, 2}, e4[4] = {1, 1, 1, 1};
        cuuint64_t d2[2] = {128, rows}; cuuint64_t s2[1] = {256}; cuuint32_t b2[2] = {64, 16}, e2[2] = {1, 1};
        int r = 0;
        r |= encode(&k4, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, kc, d4, s4, b4, e4, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        r |= encode(&v4, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, vc, d4, s4, b4, e4, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        r |= encode(&k2, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, kc, d2, s2, b2, e2, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        r |= encode(&v2, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, vc, d2, s2, b2, e2, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r) { printf("encode failed %d\n", r); return 1; }
    }
    const int smem_bytes = STAGES * STAGE_BYTES + 1024;
    CK(cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    const char* names[6] = {"1 lane, 4-D page box 4 KiB x16", "1 lane, 2-D half-page box 2 KiB x32", "16 lanes, one 4-D box each",
                            "1 lane, 1-D bulk 4 KiB x16", "2 warps (K | V), 4-D boxes", "128 threads cp.async 16 B"};
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const double gbytes = (double)num_seqs * NKV * PAGES_PER_SEQ * 8192 / 1e9;
    for (int mode = 0; mode < 6; mode++) {
        for (int rep = 0; rep < 2; rep++) stream_kernel<<<148, 192, smem_bytes>>>(k4, v4, k2, v2, kc, vc, bt, num_seqs, rep, mode, sink);
        CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(e0));
        const int iters = 8;
        for (int it = 0; it < iters; it++) stream_kernel<<<148, 192, smem_bytes>>>(k4, v4, k2, v2, kc, vc, bt, num_seqs, 2 + it, mode, sink);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        printf("mode %d  %-40s : %.1f us / launch, %.0f GB/s\n", mode, names[mode], ms / iters * 1e3, gbytes / (ms / iters * 1e-3));
    }
    return 0;
}
