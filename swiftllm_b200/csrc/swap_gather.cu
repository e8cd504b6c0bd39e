// Block swapping with DEVICE-resident id lists.  The reference (csrc/src/block_swapping.cpp:22-85, driven by
// swiftllm/worker/model.py:361-379) reads the gathered source ids and the newly allocated target ids back to the host
// (`.tolist()` x 2: two device syncs per swap) and issues two cudaMemcpyAsync per run of consecutive blocks.  Here the ids
// never leave the GPU: ONE kernel moves all blocks of the swap between the paged cache and the PINNED, device-mapped swap space
// (zero-copy loads / stores over PCIe), so a swap enqueues like any other kernel and the step stays sync-free.
// grid (n blocks, SW_SPLIT slices); every thread keeps four 16-byte accesses in flight (PCIe reads need the parallelism).
// Bytes moved: 2 * n * block_bytes over PCIe (K and V); bound by the link, not by HBM.
#include "common.cuh"

namespace sllm {

constexpr int SW_SPLIT = 8;
constexpr int SW_THREADS = 256;

__global__ void __launch_bounds__(SW_THREADS) swap_blocks_gather_kernel(const int64_t* __restrict__ src_ids,
                                                                        const int64_t* __restrict__ dst_ids,
                                                                        const uint4* __restrict__ k_src, const uint4* __restrict__ v_src,
                                                                        uint4* __restrict__ k_dst, uint4* __restrict__ v_dst,
                                                                        int64_t block_vecs) {
    const int64_t s = src_ids[blockIdx.x] * block_vecs, d = dst_ids[blockIdx.x] * block_vecs;
    const int64_t per = (block_vecs + SW_SPLIT - 1) / SW_SPLIT;
    const int64_t lo = (int64_t)blockIdx.y * per, hi = min(block_vecs, lo + per);
    for (int64_t i = lo + threadIdx.x; i < hi; i += 2 * SW_THREADS) {
        const int64_t j = i + SW_THREADS;
        const bool two = j < hi;
        const uint4 a = k_src[s + i], b = v_src[s + i];
        uint4 c = a, e = b;
        if (two) { c = k_src[s + j]; e = v_src[s + j]; }
        k_dst[d + i] = a; v_dst[d + i] = b;
        if (two) { k_dst[d + j] = c; v_dst[d + j] = e; }
    }
}

}  // namespace sllm

using namespace sllm;

extern "C" int sllm_swap_blocks_gathered(const int64_t* src_ids, const int64_t* dst_ids, int64_t n, int is_swap_in, void* k_cache,
                                         void* v_cache, void* host_k_swap, void* host_v_swap, int64_t block_bytes,
                                         sllm_stream_t stream) {
    SLLM_REQUIRE(n >= 0 && n < (1LL << 31) && block_bytes > 0 && block_bytes % 16 == 0,
                 "swap_blocks_gathered: bad sizes n=%lld block_bytes=%lld", (long long)n, (long long)block_bytes);
    if (n == 0) return 0;
    SLLM_REQUIRE(src_ids && dst_ids && k_cache && v_cache && host_k_swap && host_v_swap, "swap_blocks_gathered: null pointer");
    void *kd = nullptr, *vd = nullptr;
    cudaError_t e1 = cudaHostGetDevicePointer(&kd, host_k_swap, 0), e2 = cudaHostGetDevicePointer(&vd, host_v_swap, 0);
    if (e1 != cudaSuccess || e2 != cudaSuccess) {
        cudaGetLastError();
        set_error("swap_blocks_gathered: the swap space must be pinned, device-mapped host memory (%s)",
                  cudaGetErrorString(e1 != cudaSuccess ? e1 : e2));
        return 1;
    }
    const uint4 *ks, *vs; uint4 *kt, *vt;
    if (is_swap_in) { ks = (const uint4*)kd; vs = (const uint4*)vd; kt = (uint4*)k_cache; vt = (uint4*)v_cache; }
    else            { ks = (const uint4*)k_cache; vs = (const uint4*)v_cache; kt = (uint4*)kd; vt = (uint4*)vd; }
    dim3 grid((unsigned)n, SW_SPLIT);
    swap_blocks_gather_kernel<<<grid, SW_THREADS, 0, (cudaStream_t)stream>>>(src_ids, dst_ids, ks, vs, kt, vt, block_bytes / 16);
    return check_launch("swap_blocks_gathered");
}
