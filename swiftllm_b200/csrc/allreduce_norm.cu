// Tensor-parallel exchange fused with the op that always follows it:
//     x_sum = sum over ranks of the partial GEMM output   (one-shot all-reduce over NVLink peer memory)
//     residual <- h(x_sum) + residual ;  x_out <- rmsnorm(residual) * weight       (rmsnorm.py:39-65 semantics)
// in ONE kernel, instead of ncclAllReduce + fused_add_rmsnorm.  Every rank's partial lives in a symmetric buffer
// that all peers can read through NVLink / NVSwitch (P2P loads); each rank reads all N partials of every row in the
// SAME rank order with fp32 accumulation, so all ranks obtain bit-identical sums (no broadcast step needed).
//
// Synchronisation: a flag per (slot, source rank) in every rank's signal pad.  On entry, block 0 publishes
// "my partial for epoch e is complete" to all peers (system-scope release after the producing GEMM finished in stream
// order); every CTA waits until all N flags of its own pad reach e.  Two symmetric buffers alternate between
// consecutive exchanges, so a rank can only overwrite a buffer after every peer has passed the barrier of the
// exchange in between, i.e. finished reading it.  Epochs live in device memory (CUDA-graph replay safe).
// A watchdog traps instead of hanging if a peer never arrives.
#include "common.cuh"

namespace sllm {

constexpr int AR_MAX_RANKS = 8;

struct ArParams {
    const void* peer_buf[AR_MAX_RANKS];      // partial [T, H] of every rank (peer-mapped device pointers)
    uint32_t* peer_flags[AR_MAX_RANKS];      // signal pad of every rank: uint32 [slots][AR_MAX_RANKS]
    uint32_t* epoch;                         // local device memory: uint32 [slots] epoch, then uint32 [slots] done-counter
    void* x_out; void* residual; const void* weight;
    float eps;
    int rank, nranks, hidden, slot, has_norm;
};

__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// peer data is produced by another GPU while this kernel may already be running: no read-only / L1 caching
template <typename T> __device__ __forceinline__ Vec8<T> ld_vec8_peer(const T* p) {
    uint4 u;
    asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(u.x), "=r"(u.y), "=r"(u.z), "=r"(u.w) : "l"(p) : "memory");
    return *reinterpret_cast<Vec8<T>*>(&u);
}

template <typename T>
__global__ void __launch_bounds__(512) allreduce_add_rmsnorm_kernel(const ArParams p) {
    extern __shared__ uint4 row_smem[];
    __shared__ float red[16];
    using TT = Traits<T>;
    const int64_t t = blockIdx.x;
    const int nvec = p.hidden >> 3;
    const uint32_t e = p.epoch[p.slot] + 1;                 // epoch of this exchange (same value in every CTA)

    // ---- barrier: my partial is complete (stream order) -> tell every peer; wait for every peer's flag
    if (threadIdx.x < p.nranks) {
        if (blockIdx.x == 0) {
            __threadfence_system();
            st_release_sys(p.peer_flags[threadIdx.x] + p.slot * AR_MAX_RANKS + p.rank, e);
        }
        const uint32_t* mine = p.peer_flags[p.rank] + p.slot * AR_MAX_RANKS + threadIdx.x;
        uint32_t spins = 0;
        while ((int32_t)(ld_acquire_sys(mine) - e) < 0) {
            if (++spins > (1u << 27)) { printf("sllm: all-reduce watchdog (rank %d waits for rank %d, epoch %u)\n", p.rank, threadIdx.x, e); __trap(); }
        }
    }
    __syncthreads();

    // ---- one-shot reduction of this token row, then fused add + RMSNorm
    T* rr = reinterpret_cast<T*>(p.residual) + t * p.hidden;
    float ss = 0.f;
    for (int i = threadIdx.x; i < nvec; i += blockDim.x) {
        float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
        for (int r = 0; r < p.nranks; r++) {                 // fixed order on every rank -> identical sums
            Vec8<T> v = ld_vec8_peer(reinterpret_cast<const T*>(p.peer_buf[r]) + t * p.hidden + 8 * i);
#pragma unroll
            for (int j = 0; j < 4; j++) { float2 f = TT::to_f2(v.v[j]); acc[2 * j] += f.x; acc[2 * j + 1] += f.y; }
        }
        Vec8<T> a, b = ld_vec8(rr + 8 * i);
#pragma unroll
        for (int j = 0; j < 4; j++) a.v[j] = __hadd2_rn(TT::from_f2(make_float2(acc[2 * j], acc[2 * j + 1])), b.v[j]);   // h(h(sum) + r)
        st_vec8(rr + 8 * i, a);
        row_smem[i] = *reinterpret_cast<uint4*>(&a);
#pragma unroll
        for (int j = 0; j < 4; j++) { float2 f = TT::to_f2(a.v[j]); ss += f.x * f.x + f.y * f.y; }
    }
    if (p.has_norm) {
        ss = warp_sum(ss);
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
        if (lane == 0) red[warp] = ss;
        __syncthreads();
        float total = 0.f;
        for (int w = 0; w < nwarps; w++) total += red[w];
        const float rstd = 1.0f / sqrtf(total / (float)p.hidden + p.eps);
        T* xo = reinterpret_cast<T*>(p.x_out) + t * p.hidden;
        const T* wt = reinterpret_cast<const T*>(p.weight);
        for (int i = threadIdx.x; i < nvec; i += blockDim.x) {
            Vec8<T> a = *reinterpret_cast<Vec8<T>*>(&row_smem[i]);
            Vec8<T> w = ld_vec8(wt + 8 * i);
            Vec8<T> o;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                float2 f = TT::to_f2(a.v[j]), g = TT::to_f2(w.v[j]);
                o.v[j] = TT::from_f2(make_float2((f.x * rstd) * g.x, (f.y * rstd) * g.y));
            }
            st_vec8(xo + 8 * i, o);
        }
    }

    // ---- the last CTA to finish advances the epoch (all CTAs have read it by then)
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t* done = p.epoch + 16 + p.slot;
        if (atomicAdd(done, 1u) == gridDim.x - 1) { *done = 0; __threadfence(); p.epoch[p.slot] = e; }
    }
}

}  // namespace sllm

using namespace sllm;

// host_peer_bufs / host_peer_flags: HOST arrays of `nranks` device pointers (the symmetric buffer and the signal pad of
// every rank as mapped into this process).  epoch_state: local device memory, >= 32 uint32, zero-initialised once.
// weight == NULL: only residual <- h(sum) + residual (no norm output).
extern "C" int sllm_allreduce_add_rmsnorm(const void* const* host_peer_bufs, void* const* host_peer_flags, int rank, int nranks,
                                          int slot, void* epoch_state, void* x_out, void* residual, const void* weight, float eps,
                                          int64_t num_tokens, int hidden, sllm_dtype_t dtype, sllm_stream_t stream) {
    SLLM_REQUIRE(nranks >= 2 && nranks <= AR_MAX_RANKS && rank >= 0 && rank < nranks, "allreduce: bad rank %d of %d", rank, nranks);
    SLLM_REQUIRE(slot >= 0 && slot < 16, "allreduce: bad slot %d", slot);
    SLLM_REQUIRE(hidden > 0 && hidden % 8 == 0 && num_tokens >= 0, "allreduce: hidden (%d) must be a positive multiple of 8", hidden);
    if (num_tokens == 0) return 0;
    SLLM_REQUIRE(host_peer_bufs && host_peer_flags && epoch_state && residual && (weight == nullptr || x_out), "allreduce: null pointer");
    ArParams p;
    for (int r = 0; r < nranks; r++) { p.peer_buf[r] = host_peer_bufs[r]; p.peer_flags[r] = (uint32_t*)host_peer_flags[r]; }
    p.epoch = (uint32_t*)epoch_state; p.x_out = x_out; p.residual = residual; p.weight = weight; p.eps = eps;
    p.rank = rank; p.nranks = nranks; p.hidden = hidden; p.slot = slot; p.has_norm = weight != nullptr;
    const int nvec = hidden / 8;
    int threads = nvec >= 512 ? 512 : (nvec >= 256 ? 256 : ((nvec + 31) / 32) * 32);
    if (threads < 32) threads = 32;
    const size_t smem = (size_t)nvec * sizeof(uint4);
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == SLLM_F16) {
        if (smem > 48 * 1024) cudaFuncSetAttribute(allreduce_add_rmsnorm_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        allreduce_add_rmsnorm_kernel<__half><<<(unsigned)num_tokens, threads, smem, st>>>(p);
    } else if (dtype == SLLM_BF16) {
        if (smem > 48 * 1024) cudaFuncSetAttribute(allreduce_add_rmsnorm_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        allreduce_add_rmsnorm_kernel<__nv_bfloat16><<<(unsigned)num_tokens, threads, smem, st>>>(p);
    } else {
        SLLM_REQUIRE(false, "allreduce: unknown dtype tag %d", (int)dtype);
    }
    return check_launch("allreduce_add_rmsnorm");
}
