// Shared helpers for the fsem sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "../../include/fsem.h"

namespace fsem {

constexpr int kWarp = 32;
constexpr unsigned kFull = 0xffffffffu;

__host__ __device__ __forceinline__ int64_t round_up(int64_t x, int64_t m) { return (x + m - 1) / m * m; }
__host__ __device__ __forceinline__ int64_t ceil_div(int64_t x, int64_t m) { return (x + m - 1) / m; }

// PESQ frame count: right-pad by n % 256 zeros, frames of 512 hop 256, no centring (PESQ.py:128-133)
__host__ __device__ __forceinline__ int pesq_num_frames(int64_t n) {
    int64_t padded = n + (n % FSEM_PESQ_HOP);
    if (padded < FSEM_PESQ_NFFT) return 0;
    return (int)(1 + (padded - FSEM_PESQ_NFFT) / FSEM_PESQ_HOP);
}

// resampled length: ceil(neu * n / orig)  (torchaudio _apply_sinc_resample_kernel)
__host__ __device__ __forceinline__ int64_t stoi_resampled_len(int64_t n, int orig, int neu) {
    if (orig == neu) return n;
    return (neu * n + orig - 1) / orig;
}

// number of 256/128 analysis frames of a 10 kHz signal of length L (STOI.py:92)
__host__ __device__ __forceinline__ int stoi_num_frames(int64_t L) {
    if (L < FSEM_STOI_WIN) return 0;
    return (int)((L - FSEM_STOI_WIN) / FSEM_STOI_HOP + 1);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(kFull, v, o));
    return v;
}

// ---- TMA bulk copy (cp.async.bulk, global -> shared) with mbarrier completion, sm_90+/sm_100a ----------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t arrivals) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(arrivals) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// 1-D bulk copy: `bytes` multiple of 16, src and dst 16-byte aligned; completes `bytes` on the mbarrier
__device__ __forceinline__ void bulk_copy_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
// bounded wait: a copy that never lands (bad address) traps instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    for (uint32_t spin = 0; !mbar_try_wait(bar, parity); ++spin)
        if (spin > (1u << 26)) __trap();
}
// order earlier generic-proxy shared-memory accesses before later async-proxy (bulk copy) writes
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// 16-byte read-only load that asks L2 to fetch the whole 256-byte block around the address: for row-strided tile
// walks (128 bytes per row and step) every other step then hits L2 and each DRAM page is opened half as often
__device__ __forceinline__ float4 ldg_f4_l2_256(const float* p) {
    float4 v;
    asm volatile("ld.global.nc.L2::256B.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}

__device__ __forceinline__ int item_length(const int32_t* lengths, int64_t item, int64_t n) {
    if (lengths == nullptr) return (int)n;
    int l = lengths[item];
    return l < 0 ? 0 : (l > n ? (int)n : l);
}

}  // namespace fsem
