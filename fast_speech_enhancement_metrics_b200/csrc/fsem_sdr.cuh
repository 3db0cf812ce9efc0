// SDR kernels (sm_100a) -- SURVEY.md 8f rank 3.  fast_se_metrics/SDR.py:7-97 computes, per item,
//   r[l] = sum_t c[t] c[t+l],  b[l] = sum_t c[t] d[t+l],  l = 0..511   (unit-norm signals; via 2^19-point FFTs there),
//   solves the 512 x 512 symmetric Toeplitz system T(r) x = b (Cholesky there) and maps coh = b.x to dB.
// Here:
//   sdr_norm_kernel   ||x||^2 of every signal (fp64)                                              (SDR.py:65-71)
//   sdr_corr_kernel   the 2 x 512 correlation lags directly in the time domain: 1024 MACs per sample, register-tiled
//                     8 lags x 8 samples per thread from a padded shared-memory tile, fp32 per 2048-sample tile,
//                     fp64 across tiles                                                            (SDR.py:34-49)
//   sdr_solve_kernel  one CTA per item: Levinson-Durbin recursion in fp64 (two barriers per order), coherence, dB
//                                                                                                  (SDR.py:7-31, 88-95)
#pragma once
#include "fsem_common.cuh"

namespace fsem {

constexpr int kSdrLags = 512;
constexpr int kSdrTile = 2048;                 // samples of t per shared-memory tile
constexpr int kSdrSuper = 16;                  // tiles per CTA (32768 samples)
constexpr int kSdrThreads = kSdrLags / 8;      // 64: every thread owns 8 consecutive lags

__global__ void __launch_bounds__(256)
sdr_norm_kernel(const float* __restrict__ clean, const float* __restrict__ deg, const int32_t* __restrict__ lengths,
                int64_t batch, int64_t n, int64_t stride, double* __restrict__ energy /* [2][batch] */) {
    __shared__ double s_part[8];
    const int64_t sig = blockIdx.x;
    const int64_t item = sig < batch ? sig : sig - batch;
    const int len = item_length(lengths, item, n);
    const float* __restrict__ x = (sig < batch ? clean : deg) + item * stride;
    double acc = 0.0;
    for (int i = threadIdx.x; i < len; i += 256) { const float v = __ldg(x + i); acc = fma((double)v, (double)v, acc); }
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0.0;
        for (int i = 0; i < 8; ++i) a += s_part[i];
        energy[sig] = a;
    }
}

// Shared tile layout: float4 q of the staged samples lives at position q + (q >> 1) (one pad float4 after every
// two), so that a thread's lag window (lane pitch 2 float4 -> 3 positions) reads conflict-free LDS.128.
__device__ __forceinline__ int sdr_pos(int q) { return q + (q >> 1); }
constexpr int kSdrQuads = (kSdrTile + kSdrLags) / 4;          // 640 float4 per signal per tile
constexpr int kSdrSmemQuads = kSdrQuads + kSdrQuads / 2 + 2;

__global__ void __launch_bounds__(kSdrThreads)
sdr_corr_kernel(const float* __restrict__ clean, const float* __restrict__ deg, const int32_t* __restrict__ lengths,
                int64_t batch, int64_t n, int64_t stride, int nsuper,
                double* __restrict__ partial /* [batch][nsuper][2][512] */) {
    __shared__ float4 s_c[kSdrSmemQuads];
    __shared__ float4 s_d[kSdrSmemQuads];
    const int tid = threadIdx.x;
    const int64_t item = blockIdx.y;
    const int sup = blockIdx.x;
    const int len = item_length(lengths, item, n);
    const float* __restrict__ c = clean + item * stride;
    const float* __restrict__ d = deg + item * stride;
    double acc_r[8], acc_b[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) { acc_r[q] = 0.0; acc_b[q] = 0.0; }
    const int t_begin = sup * (kSdrSuper * kSdrTile);
    for (int tile = 0; tile < kSdrSuper; ++tile) {
        const int t0 = t_begin + tile * kSdrTile;
        if (t0 >= len) break;                                            // uniform
        __syncthreads();
        for (int q = tid; q < kSdrQuads; q += kSdrThreads) {
            const int i = t0 + 4 * q;
            float4 vc, vd;
            vc.x = (i < len) ? __ldg(c + i) : 0.f;         vd.x = (i < len) ? __ldg(d + i) : 0.f;
            vc.y = (i + 1 < len) ? __ldg(c + i + 1) : 0.f; vd.y = (i + 1 < len) ? __ldg(d + i + 1) : 0.f;
            vc.z = (i + 2 < len) ? __ldg(c + i + 2) : 0.f; vd.z = (i + 2 < len) ? __ldg(d + i + 2) : 0.f;
            vc.w = (i + 3 < len) ? __ldg(c + i + 3) : 0.f; vd.w = (i + 3 < len) ? __ldg(d + i + 3) : 0.f;
            s_c[sdr_pos(q)] = vc;
            s_d[sdr_pos(q)] = vd;
        }
        __syncthreads();
        float r[8], b[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) { r[q] = 0.f; b[q] = 0.f; }
        // 8 samples of t per step: c[t..t+7] (broadcast) against c/d[t + 8*tid .. t + 8*tid + 14]
        for (int t = 0; t < kSdrTile; t += 8) {
            const int qb = t >> 2;                                       // even
            const float4 cb0 = s_c[sdr_pos(qb)], cb1 = s_c[sdr_pos(qb + 1)];
            const float cb[8] = {cb0.x, cb0.y, cb0.z, cb0.w, cb1.x, cb1.y, cb1.z, cb1.w};
            const int ql = qb + 2 * tid;                                 // even: positions ql*3/2 + {0, 1, 3, 4}
            const int base = sdr_pos(ql);
            float cw[16], dw[16];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float4 vc = s_c[base + i + (i >> 1)], vd = s_d[base + i + (i >> 1)];
                cw[4 * i] = vc.x; cw[4 * i + 1] = vc.y; cw[4 * i + 2] = vc.z; cw[4 * i + 3] = vc.w;
                dw[4 * i] = vd.x; dw[4 * i + 1] = vd.y; dw[4 * i + 2] = vd.z; dw[4 * i + 3] = vd.w;
            }
#pragma unroll
            for (int a = 0; a < 8; ++a) {
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    r[q] = fmaf(cb[a], cw[a + q], r[q]);
                    b[q] = fmaf(cb[a], dw[a + q], b[q]);
                }
            }
        }
#pragma unroll
        for (int q = 0; q < 8; ++q) { acc_r[q] += (double)r[q]; acc_b[q] += (double)b[q]; }
    }
    double* out = partial + ((item * nsuper + sup) * 2) * (int64_t)kSdrLags + 8 * tid;
#pragma unroll
    for (int q = 0; q < 8; ++q) { out[q] = acc_r[q]; out[kSdrLags + q] = acc_b[q]; }
}

// Levinson-Durbin in float64, one CTA per item.  128 threads own four lags each (i = tid + 128 q): the recursion's two dot
// products are inherent (2 k multiply-adds in step k), everything else -- the cross-warp sums, lambda, mu -- is computed
// redundantly by every thread, and with 512 threads that redundant float64 work (16-term sums, two divisions per thread
// and step) was 4x what the dot products cost: 16.0 ms at 8192 items.  Four warps per CTA also means 16 CTAs per SM to
// overlap each other's two barriers per step.
constexpr int kSdrSolveThreads = 128;
constexpr int kSdrSolvePer = kSdrLags / kSdrSolveThreads;     // 4
static_assert(kSdrLags % kSdrSolveThreads == 0, "every thread owns the same number of lags");

__global__ void __launch_bounds__(kSdrSolveThreads)
sdr_solve_kernel(const double* __restrict__ partial, int nsuper, const double* __restrict__ energy, int64_t batch,
                 float* __restrict__ sdr_out) {
    __shared__ double s_r[kSdrLags], s_b[kSdrLags], s_a[kSdrLags], s_x[kSdrLags];
    __shared__ double s_p1[2][kSdrSolveThreads / 32], s_p2[2][kSdrSolveThreads / 32];
    const int t = threadIdx.x;
    const int lane = t & 31, warp = t >> 5;
    const int64_t item = blockIdx.x;
    // unit-norm signals (SDR.py:69-70: x / clamp(||x||, 1e-6)): scale the raw correlations instead
    const double nc = fmax(sqrt(energy[item]), 1e-6), nd = fmax(sqrt(energy[batch + item]), 1e-6);
#pragma unroll
    for (int q = 0; q < kSdrSolvePer; ++q) {
        const int i = t + kSdrSolveThreads * q;
        double rr = 0.0, bb = 0.0;
        for (int s = 0; s < nsuper; ++s) {
            const double* p = partial + ((item * nsuper + s) * 2) * (int64_t)kSdrLags;
            rr += p[i];
            bb += p[kSdrLags + i];
        }
        // the reference keeps r and b in float32 (SDR.py:78-79)
        s_r[i] = (double)(float)(rr / (nc * nc));
        s_b[i] = (double)(float)(bb / (nc * nd));
        s_a[i] = (i == 0) ? 1.0 : 0.0;
        s_x[i] = 0.0;
    }
    __syncthreads();
    double E = s_r[0];
    if (t == 0) s_x[0] = s_b[0] / E;
    __syncthreads();
    // a = forward predictor (a[0] = 1), x = solution of the leading k x k system
    for (int k = 1; k < kSdrLags; ++k) {
        double p1 = 0.0, p2 = 0.0;
#pragma unroll
        for (int q = 0; q < kSdrSolvePer; ++q) {
            const int i = t + kSdrSolveThreads * q;
            if (i < k) { const double rk = s_r[k - i]; p1 = fma(s_a[i], rk, p1); p2 = fma(s_x[i], rk, p2); }
        }
        p1 = warp_sum(p1);
        p2 = warp_sum(p2);
        const int slot = k & 1;                                  // two slots: the next step's writes cannot race this step's reads
        if (lane == 0) { s_p1[slot][warp] = p1; s_p2[slot][warp] = p2; }
        __syncthreads();
        double d1 = 0.0, d2 = 0.0;
#pragma unroll
        for (int w = 0; w < kSdrSolveThreads / 32; ++w) { d1 += s_p1[slot][w]; d2 += s_p2[slot][w]; }
        const double lambda = -d1 / E;
        E = E * (1.0 - lambda * lambda);
        const double mu = (s_b[k] - d2) / E;
        double ai[kSdrSolvePer], ak[kSdrSolvePer];
#pragma unroll
        for (int q = 0; q < kSdrSolvePer; ++q) {
            const int i = t + kSdrSolveThreads * q;
            ai[q] = 0.0; ak[q] = 0.0;
            if (i <= k) { ai[q] = s_a[i]; ak[q] = s_a[k - i]; }
        }
        __syncthreads();
#pragma unroll
        for (int q = 0; q < kSdrSolvePer; ++q) {
            const int i = t + kSdrSolveThreads * q;
            if (i <= k) {
                s_a[i] = ai[q] + lambda * ak[q];
                s_x[i] += mu * (ak[q] + lambda * ai[q]);         // a_new[k - i]
            }
        }
    }
    __syncthreads();
    double coh = 0.0;
#pragma unroll
    for (int q = 0; q < kSdrSolvePer; ++q) coh = fma(s_b[t + kSdrSolveThreads * q], s_x[t + kSdrSolveThreads * q], coh);
    coh = warp_sum(coh);
    if (lane == 0) s_p1[0][warp] = coh;
    __syncthreads();
    if (t == 0) {
        double tot = 0.0;
        for (int w = 0; w < kSdrSolveThreads / 32; ++w) tot += s_p1[0][w];
        const float cohf = (float)tot;
        const float ratio = cohf / fmaxf(1.f - cohf, 1e-8f);                  // SDR.py:91-95
        sdr_out[item] = 10.f * log10f(fmaxf(ratio, 1e-8f));
    }
}

}  // namespace fsem
