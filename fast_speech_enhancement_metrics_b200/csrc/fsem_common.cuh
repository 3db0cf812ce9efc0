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

__device__ __forceinline__ int item_length(const int32_t* lengths, int64_t item, int64_t n) {
    if (lengths == nullptr) return (int)n;
    int l = lengths[item];
    return l < 0 ? 0 : (l > n ? (int)n : l);
}

}  // namespace fsem
