// LSD (log-spectral distance) kernels (sm_100a) -- SURVEY.md 8f rank 3, built on the same packed warp FFT.
//
//   lsd_scale_kernel     per item: alpha = <clean, deg> / (<deg, deg> + eps)                       (LSD.py:37-39)
//   lsd_frames_kernel    centred Hann-512/256 STFT of (clean, deg) packed in one complex FFT per frame,
//                        per frame sqrt(mean_k log(|C|^2 / (|alpha||D| + eps)^2 + eps)^2) over the 257 bins
//                        (LSD.py:20-31, 47-50); |STFT(alpha*d)| = |alpha| * |STFT(d)|, so the scaled signal is
//                        never materialised
//   lsd_finalize_kernel  mean over frames in a fixed order                                        (LSD.py:50)
#pragma once
#include "fsem_common.cuh"
#include "fsem_fft.cuh"

namespace fsem {

constexpr float kLsdEps = 1e-8f;

// number of frames of torch.stft(center=True, hop 256): 1 + n / 256
__host__ __device__ __forceinline__ int lsd_num_frames(int64_t n) { return (int)(1 + n / 256); }

__global__ void __launch_bounds__(256)
lsd_scale_kernel(const float* __restrict__ clean, const float* __restrict__ deg, const int32_t* __restrict__ lengths,
                 int64_t batch, int64_t n, int64_t stride, float* __restrict__ alpha) {
    __shared__ double s_cd[8], s_dd[8];
    const int64_t item = blockIdx.x;
    const int len = item_length(lengths, item, n);
    const float* __restrict__ c = clean + item * stride;
    const float* __restrict__ d = deg + item * stride;
    double cd = 0.0, dd = 0.0;
    for (int i = threadIdx.x; i < len; i += 256) {
        const float x = __ldg(c + i), y = __ldg(d + i);
        cd = fma((double)x, (double)y, cd);
        dd = fma((double)y, (double)y, dd);
    }
    cd = warp_sum(cd);
    dd = warp_sum(dd);
    if ((threadIdx.x & 31) == 0) { s_cd[threadIdx.x >> 5] = cd; s_dd[threadIdx.x >> 5] = dd; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0.0, b = 0.0;
        for (int i = 0; i < 8; ++i) { a += s_cd[i]; b += s_dd[i]; }
        alpha[item] = (float)a / ((float)b + kLsdEps);       // float32 division like the reference
    }
}

constexpr int kLsdWarps = 8;

__device__ __forceinline__ float lsd_sqrt(float x) { float y; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float lsd_rcp(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float lsd_lg2(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// kVec2: rows are 8-byte aligned with an even pitch, so a lane fetches its sample pairs (2L, 2L + 1) of an interior frame
// with one 64-bit load and no bounds checks.  The per-bin log ratio goes through the SFU (sqrt / rcp / lg2 .approx, ~1e-6
// relative: two orders of magnitude inside the 5e-4 the float32 FFT round-off of near-eps bins already costs, tests/
// test_lsd.py): 10.2 -> see DESIGN.md at 8192 x 10 s.
template <bool kVec2>
__global__ void __launch_bounds__(kLsdWarps * 32, 2)
lsd_frames_kernel(const float* __restrict__ clean, const float* __restrict__ deg, const int32_t* __restrict__ lengths,
                  int64_t batch, int64_t n, int64_t stride, int tmax, const float* __restrict__ hann,
                  const float* __restrict__ alpha, float* __restrict__ frame_lsd /* [batch][tmax] */) {
    __shared__ __align__(16) float2 s_buf[kLsdWarps][kFftBufElems];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    float2* buf = s_buf[warp];
    FftTwiddles tw;
    tw.init(lane);
    float win[16];                                               // win[8h + j] = hann[2*lane + h + 64 j]
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int j = 0; j < 8; ++j) win[8 * h + j] = hann[fft_in_index(lane, h, j)];

    const int64_t units = batch * (int64_t)tmax;
    const int64_t nwarps = (int64_t)gridDim.x * kLsdWarps;
    const int64_t per = (units + nwarps - 1) / nwarps;
    const int64_t u0 = ((int64_t)blockIdx.x * kLsdWarps + warp) * per;
    const int64_t u1 = min(units, u0 + per);
    if (u0 >= u1) return;
    int64_t item = u0 / tmax;
    int f = (int)(u0 - item * tmax);
    int len = item_length(lengths, item, n);
    int T = lsd_num_frames(len);
    for (int64_t u = u0; u < u1; ++u) {
        if (f < T) {
            const float* __restrict__ c = clean + item * stride;
            const float* __restrict__ d = deg + item * stride;
            const int first = f * 256 - 256;                     // centred frames, zero ("constant") padding
            float re[16], im[16];
            if (kVec2 && first >= 0 && first + 512 <= len) {     // interior frame (warp-uniform)
                const float2* __restrict__ pc2 = reinterpret_cast<const float2*>(c + first) + lane;
                const float2* __restrict__ pd2 = reinterpret_cast<const float2*>(d + first) + lane;
#pragma unroll
                for (int j = 0; j < 8; ++j) {                    // samples 2*lane + 64 j and + 1 = fft_in_index(lane, 0 | 1, j)
                    const float2 cv = __ldg(pc2 + 32 * j), dv = __ldg(pd2 + 32 * j);
                    re[j] = cv.x * win[j];  re[8 + j] = cv.y * win[8 + j];
                    im[j] = dv.x * win[j];  im[8 + j] = dv.y * win[8 + j];
                }
            } else {
#pragma unroll
                for (int h = 0; h < 2; ++h)
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const int i = first + fft_in_index(lane, h, j);
                        const bool ok = i >= 0 && i < len;
                        re[8 * h + j] = ok ? __ldg(c + i) * win[8 * h + j] : 0.f;
                        im[8 * h + j] = ok ? __ldg(d + i) * win[8 * h + j] : 0.f;
                    }
            }
            float ar[8], ai[8], br[8], bi[8];
            warp_fft512<false>(re, im, buf, tw, lane, ar, ai, br, bi);
            float pc[8], pd[8];
            packed_power_regs(ar, ai, br, bi, lane, pc, pd);     // 4x the power of this lane's 8 bins below 256
            const float a = fabsf(alpha[item]);
            float acc = 0.f;                                     // sum of log2(...)^2; ln(2)^2 is applied once per frame
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float den = fmaf(a, lsd_sqrt(pd[j] * kPackedPowerScale), kLsdEps);
                const float l = lsd_lg2(fmaf(pc[j] * kPackedPowerScale, lsd_rcp(den * den), kLsdEps));
                acc = fmaf(l, l, acc);
            }
            if (lane == 0) {                                     // Nyquist bin 256 = A[4] of lane 0: C = Re Z, D = Im Z
                const float den = fmaf(a, fabsf(ai[4]), kLsdEps);
                const float l = lsd_lg2(fmaf(ar[4] * ar[4], lsd_rcp(den * den), kLsdEps));
                acc = fmaf(l, l, acc);
            }
            acc = warp_sum(acc);
            constexpr float kLn2Sq = 0.69314718056f * 0.69314718056f;
            if (lane == 0) frame_lsd[item * tmax + f] = sqrtf(acc * (kLn2Sq / 257.f));
            __syncwarp();
        }
        if (++f == tmax) {
            f = 0;
            ++item;
            if (item < batch) { len = item_length(lengths, item, n); T = lsd_num_frames(len); }
        }
    }
}

__global__ void __launch_bounds__(128)
lsd_finalize_kernel(const float* __restrict__ frame_lsd, const int32_t* __restrict__ lengths, int64_t batch, int64_t n,
                    int tmax, float* __restrict__ out) {
    const int64_t item = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (item >= batch) return;
    const int T = lsd_num_frames(item_length(lengths, item, n));
    double acc = 0.0;
    for (int f = 0; f < T; ++f) acc += (double)frame_lsd[item * tmax + f];
    out[item] = (float)(acc / (double)T);
}

}  // namespace fsem
