// Warp-cooperative 512-point complex FFT (radix-8 x 8 x 8) for sm_100a.
//
// One warp transforms one 512-point complex sequence z[n] = c[n] + i*d[n] that packs
// two REAL frames (clean, degraded) of the same utterance; the two real power spectra
// are separated afterwards from Z[k] and Z[512-k].  Each lane owns 16 complex points in
// registers; the three radix-8 passes exchange data through a per-warp shared-memory
// buffer (float2, padded so every 64-bit access is bank-conflict-free).
//
// Replaces on this path: torch.stft / torchaudio Spectrogram -> cuFFT / pocketfft
// (reference call sites PESQ.py:133 and STOI.py:50-61).
//
// Index algebra (decimation in frequency), n = n0 + 8*n1 + 64*n2, k = k0 + 8*k1 + 64*k2:
//   pass 1: a[n', k0]      = W512^(n'*k0) * sum_{n2} z[n' + 64*n2] W8^(n2*k0),  n' = n0 + 8*n1
//   pass 2: b[n0, k1; k0]  = W64^(n0*k1)  * sum_{n1} a[n0 + 8*n1, k0] W8^(n1*k1)
//   pass 3: Z[k0+8k1+64k2] =                sum_{n0} b[n0, k1; k0]    W8^(n0*k2)
#pragma once
#include <cuda_runtime.h>

namespace fsem {

constexpr int kFftN = 512;
// per-warp exchange buffer, in float2 elements (see layouts below)
constexpr int kFftBufElems = 640;

__device__ __forceinline__ void cmul(float& xr, float& xi, float wr, float wi) {
    float tr = xr * wr - xi * wi;
    xi = xr * wi + xi * wr;
    xr = tr;
}

// 4-point DFT (forward), natural-order output
__device__ __forceinline__ void dft4(float& r0, float& i0, float& r1, float& i1,
                                     float& r2, float& i2, float& r3, float& i3) {
    float c0r = r0 + r2, c0i = i0 + i2;
    float c2r = r0 - r2, c2i = i0 - i2;
    float c1r = r1 + r3, c1i = i1 + i3;
    // (x1 - x3) * (-i) = (im, -re)
    float c3r = i1 - i3, c3i = r3 - r1;
    r0 = c0r + c1r; i0 = c0i + c1i;
    r2 = c0r - c1r; i2 = c0i - c1i;
    r1 = c2r + c3r; i1 = c2i + c3i;
    r3 = c2r - c3r; i3 = c2i - c3i;
}

// 8-point forward DFT in place, natural-order input and output.
// kUpperZero: inputs 4..7 are known to be zero (STOI's 256-sample chunk zero-padded to 512).
template <bool kUpperZero>
__device__ __forceinline__ void dft8(float (&r)[8], float (&i)[8]) {
    constexpr float kH = 0.70710678118654752440f;
    float ar[4], ai[4], br[4], bi[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        if (kUpperZero) {
            ar[j] = r[j]; ai[j] = i[j];
            br[j] = r[j]; bi[j] = i[j];
        } else {
            ar[j] = r[j] + r[j + 4]; ai[j] = i[j] + i[j + 4];
            br[j] = r[j] - r[j + 4]; bi[j] = i[j] - i[j + 4];
        }
    }
    // b_j *= W8^j : W8 = (1 - i)/sqrt2, W8^2 = -i, W8^3 = (-1 - i)/sqrt2
    {
        float tr = (br[1] + bi[1]) * kH, ti = (bi[1] - br[1]) * kH;
        br[1] = tr; bi[1] = ti;
        tr = bi[2]; ti = -br[2];
        br[2] = tr; bi[2] = ti;
        tr = (bi[3] - br[3]) * kH; ti = -(br[3] + bi[3]) * kH;
        br[3] = tr; bi[3] = ti;
    }
    dft4(ar[0], ai[0], ar[1], ai[1], ar[2], ai[2], ar[3], ai[3]);   // even outputs 0,2,4,6
    dft4(br[0], bi[0], br[1], bi[1], br[2], bi[2], br[3], bi[3]);   // odd outputs 1,3,5,7
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        r[2 * j] = ar[j];     i[2 * j] = ai[j];
        r[2 * j + 1] = br[j]; i[2 * j + 1] = bi[j];
    }
}

// Lane-resident twiddles, computed once per warp and reused for every frame.
struct FftTwiddles {
    float w1r[2][7], w1i[2][7];   // W512^(n'*k0), n' = lane + 32*h, k0 = 1..7
    float w2r[7], w2i[7];         // W64^(n0*k1),  n0 = lane & 7,    k1 = 1..7
    __device__ __forceinline__ void init(int lane) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            int np = lane + 32 * h;
#pragma unroll
            for (int k0 = 1; k0 < 8; ++k0) {
                float s, c;
                sincospif(-(float)(np * k0) * (1.0f / 256.0f), &s, &c);
                w1r[h][k0 - 1] = c; w1i[h][k0 - 1] = s;
            }
        }
        int n0 = lane & 7;
#pragma unroll
        for (int k1 = 1; k1 < 8; ++k1) {
            float s, c;
            sincospif(-(float)(n0 * k1) * (1.0f / 32.0f), &s, &c);
            w2r[k1 - 1] = c; w2i[k1 - 1] = s;
        }
    }
};

// Forward 512-point FFT.  On entry lane L holds z[L + 32*m] in (re[m], im[m]), m = 0..15.
// On exit the spectrum is in shared memory: Z[k] at buf[fft_out_index(k)].
// `buf` is this warp's private buffer of kFftBufElems float2.  Ends with __syncwarp().
__device__ __forceinline__ int fft_out_index(int k) { return k + 2 * (k >> 3); }

template <bool kUpperZero>
__device__ __forceinline__ void warp_fft512(float (&re)[16], float (&im)[16], float2* buf,
                                            const FftTwiddles& tw, int lane) {
    // ---- pass 1: radix-8 over n2 for n' = lane (even m) and n' = lane + 32 (odd m)
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        float r[8], i[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) { r[j] = re[2 * j + h]; i[j] = im[2 * j + h]; }
        dft8<kUpperZero>(r, i);
#pragma unroll
        for (int k0 = 1; k0 < 8; ++k0) cmul(r[k0], i[k0], tw.w1r[h][k0 - 1], tw.w1i[h][k0 - 1]);
        // exchange 1 layout: a[n', k0] at k0*72 + n'
#pragma unroll
        for (int k0 = 0; k0 < 8; ++k0) buf[k0 * 72 + lane + 32 * h] = make_float2(r[k0], i[k0]);
    }
    __syncwarp();
    // ---- pass 2: radix-8 over n1 for (n0, k0) = (lane & 7, (lane >> 3) + 4*h)
    const int n0 = lane & 7;
    float r2[2][8], i2[2][8];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int k0 = (lane >> 3) + 4 * h;
#pragma unroll
        for (int n1 = 0; n1 < 8; ++n1) {
            float2 v = buf[k0 * 72 + n0 + 8 * n1];
            r2[h][n1] = v.x; i2[h][n1] = v.y;
        }
    }
    __syncwarp();
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int k0 = (lane >> 3) + 4 * h;
        dft8<false>(r2[h], i2[h]);
#pragma unroll
        for (int k1 = 1; k1 < 8; ++k1) cmul(r2[h][k1], i2[h][k1], tw.w2r[k1 - 1], tw.w2i[k1 - 1]);
        // exchange 2 layout: b[n0, k1; k0] at k0*72 + k1*9 + n0
#pragma unroll
        for (int k1 = 0; k1 < 8; ++k1) buf[k0 * 72 + k1 * 9 + n0] = make_float2(r2[h][k1], i2[h][k1]);
    }
    __syncwarp();
    // ---- pass 3: radix-8 over n0 for (k1, k0) = (lane & 7, (lane >> 3) + 4*h)
    const int k1 = lane & 7;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int k0 = (lane >> 3) + 4 * h;
#pragma unroll
        for (int m0 = 0; m0 < 8; ++m0) {
            float2 v = buf[k0 * 72 + k1 * 9 + m0];
            r2[h][m0] = v.x; i2[h][m0] = v.y;
        }
    }
    __syncwarp();
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int k0 = (lane >> 3) + 4 * h;
        dft8<false>(r2[h], i2[h]);
        // fft_out_index(k0 + 8*k1 + 64*k2) = k0 + 10*k1 + 80*k2 (k0 < 8): one base address + immediates
        float2* out = buf + (k0 + 10 * k1);
#pragma unroll
        for (int k2 = 0; k2 < 8; ++k2) out[80 * k2] = make_float2(r2[h][k2], i2[h][k2]);
    }
    __syncwarp();
}

// Power spectra of the two packed real frames from Z[k] (a) and Z[N-k] (b):
//   |C[k]|^2 = ((Zr[k] + Zr[N-k])^2 + (Zi[k] - Zi[N-k])^2) / 4
//   |D[k]|^2 = ((Zi[k] + Zi[N-k])^2 + (Zr[k] - Zr[N-k])^2) / 4
__device__ __forceinline__ void packed_power_pair(float2 a, float2 b, float& pc, float& pd) {
    float sr = a.x + b.x, dr = a.x - b.x;
    float si = a.y + b.y, di = a.y - b.y;
    pc = 0.25f * fmaf(sr, sr, di * di);
    pd = 0.25f * fmaf(si, si, dr * dr);
}

// Lane L gets the power of the 8 CONSECUTIVE bins k = 8L .. 8L+7 of both packed frames.
// With fft_out_index(k) = 10*(k >> 3) + (k & 7) the eight Z[k] are four conflict-free LDS.128 at
// buf + 10L, the mirrored Z[512-k], j = 1..7, are four LDS.128 at group 63-L, and Z[512-8L] is one LDS.64.
__device__ __forceinline__ void packed_power8(const float2* buf, int lane, float (&pc)[8], float (&pd)[8]) {
    float2 a[8], m[8];
    const float4* pa = reinterpret_cast<const float4*>(buf + 10 * lane);
    const float4* pm = reinterpret_cast<const float4*>(buf + 10 * (63 - lane));
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float4 v = pa[i];
        a[2 * i] = make_float2(v.x, v.y); a[2 * i + 1] = make_float2(v.z, v.w);
        float4 w = pm[i];
        m[2 * i] = make_float2(w.x, w.y); m[2 * i + 1] = make_float2(w.z, w.w);
    }
    const float2 b0 = buf[lane == 0 ? 0 : 10 * (64 - lane)];
    packed_power_pair(a[0], b0, pc[0], pd[0]);
#pragma unroll
    for (int j = 1; j < 8; ++j) packed_power_pair(a[j], m[8 - j], pc[j], pd[j]);
}

// Static description of how one lane's 8 consecutive bins map onto contiguous frequency bands.
struct BandPlan {
    unsigned start_mask;   // bit j set: bin 8L + j is the first bin of a band
    int first_band;        // index of the band that starts at the lowest set bit (bands are consecutive)
    unsigned add_mask;     // bit d set: the band open at this lane's right edge also owns the head bins of lane L + d
    // Build from band start bins (ascending, starts[0] == 0 expected so that every bin belongs to a band).
    __device__ __forceinline__ void init(const int32_t* starts, int nbands, int lane) {
        start_mask = 0u;
        first_band = 0x7fffffff;
        for (int b = 0; b < nbands; ++b) {
            int f = starts[b];
            if ((f >> 3) == lane) {
                start_mask |= 1u << (f & 7);
                first_band = min(first_band, b);
            }
        }
        // lanes without a band start are "transparent": the open band of an earlier lane runs through them
        const unsigned transparent = __ballot_sync(0xffffffffu, start_mask == 0u);
        add_mask = 0u;
        bool open = true;
        for (int d = 1; d < 32; ++d) {
            if (!open || lane + d >= 32) break;
            add_mask |= 1u << d;
            open = (transparent >> (lane + d)) & 1u;
        }
    }
};

// Segmented band sums of the two power rows.  Bands that live inside the lane are emitted directly, the band
// that is open at the lane's right edge collects the head partials of the following lanes (at most kSpan of
// them) through shuffles.  emit(band, sum_clean, sum_deg) is called by the lane in which the band STARTS.
template <int kSpan, typename Emit>
__device__ __forceinline__ void band_sums8(const float (&pc)[8], const float (&pd)[8], const BandPlan& plan, int lane,
                                           Emit emit) {
    float run_c = 0.f, run_d = 0.f, head_c = 0.f, head_d = 0.f;
    int nb = -1;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        if ((plan.start_mask >> j) & 1u) {
            if (nb < 0) { head_c = run_c; head_d = run_d; }
            else emit(plan.first_band + nb, run_c, run_d);
            ++nb;
            run_c = 0.f; run_d = 0.f;
        }
        run_c += pc[j];
        run_d += pd[j];
    }
    const bool transparent = nb < 0;           // no band starts here: all 8 bins continue an earlier band
    if (transparent) { head_c = run_c; head_d = run_d; run_c = 0.f; run_d = 0.f; }
#pragma unroll
    for (int d = 1; d <= kSpan; ++d) {
        const float hc = __shfl_down_sync(0xffffffffu, head_c, d);
        const float hd = __shfl_down_sync(0xffffffffu, head_d, d);
        if ((plan.add_mask >> d) & 1u) { run_c += hc; run_d += hd; }
    }
    if (!transparent) emit(plan.first_band + nb, run_c, run_d);
}

}  // namespace fsem
