// Warp-cooperative 512-point complex FFT (radix-8 x 8 x 8) for sm_100a.
//
// One warp transforms one 512-point complex sequence z[n] = c[n] + i*d[n] that packs
// two REAL frames (clean, degraded) of the same utterance; the two real power spectra
// are separated afterwards from Z[k] and Z[512-k].  Each lane owns 16 complex points in
// registers; the three radix-8 passes exchange data through a per-warp shared-memory
// buffer.  Every lane always owns PAIRS of values that are neighbours in the exchange
// layouts, so all exchange traffic is 128-bit (STS.128 / LDS.128, conflict-free by
// padding): 49 shared-memory instructions per transform instead of 89 with 64-bit ones
// (round 1), and the band sums that follow need no per-bin predicates (BandGather).
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
    float w1r[2][7], w1i[2][7];   // W512^(n'*k0), n' = 2*lane + h, k0 = 1..7
    float w2r[7], w2i[7];         // W64^(n0*k1),  n0 = lane & 7,   k1 = 1..7
    __device__ __forceinline__ void init(int lane) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            int np = 2 * lane + h;
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

// Input sample index of register slot (h, j) of lane L:  n = 2L + h + 64 j  (h = 0, 1; j = 0..7).
// A lane therefore reads its inputs as PAIRS of consecutive samples (one 64-bit load per j and signal).
__device__ __forceinline__ int fft_in_index(int lane, int h, int j) { return 2 * lane + h + 64 * j; }

// Forward 512-point FFT.  On entry lane L holds z[2L + h + 64 j] in (re[8h + j], im[8h + j]).
// On exit the spectrum is in shared memory: Z[k] at buf[fft_out_index(k)].
// `buf` is this warp's private buffer of kFftBufElems float2 (16-byte aligned).  Ends with __syncwarp().
//
// Exchange layouts (float2 units; all accesses 128-bit, conflict-free per quarter-warp -- checked by emulation):
//   1: a[n', k0]           at (n' >> 1) * 18 + (n' & 1) * 8 + k0        lane writes rows n' = 2L, 2L+1
//   2: b[n0, k1; k0 pair]  at 2 * ((k0 >> 1) * 72 + k1 * 9 + n0) + (k0 & 1)
//   3: Z[k]                at 10 * (k >> 3) + (k & 7)
// Lane roles: pass 2 lane = (n0 = lane & 7, j = lane >> 3) owns k0 = 2j, 2j+1; pass 3 lane = (k1 = lane & 7, j).
__device__ __forceinline__ int fft_out_index(int k) { return k + 2 * (k >> 3); }

template <bool kUpperZero>
__device__ __forceinline__ void warp_fft512(float (&re)[16], float (&im)[16], float2* buf,
                                            const FftTwiddles& tw, int lane) {
    float4* buf4 = reinterpret_cast<float4*>(buf);
    // ---- pass 1: radix-8 over n2 (the j index) for n' = 2*lane + h
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        float r[8], i[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) { r[j] = re[8 * h + j]; i[j] = im[8 * h + j]; }
        dft8<kUpperZero>(r, i);
#pragma unroll
        for (int k0 = 1; k0 < 8; ++k0) cmul(r[k0], i[k0], tw.w1r[h][k0 - 1], tw.w1i[h][k0 - 1]);
        // exchange 1: row n' = 2*lane + h starts at float2 18*lane + 8*h = float4 9*lane + 4*h
        float4* row = buf4 + 9 * lane + 4 * h;
#pragma unroll
        for (int q = 0; q < 4; ++q) row[q] = make_float4(r[2 * q], i[2 * q], r[2 * q + 1], i[2 * q + 1]);
    }
    __syncwarp();
    // ---- pass 2: radix-8 over n1 for (n0, k0 in {2j, 2j+1}); r2[s][.] is the sequence of k0 = 2j + s
    const int lo3 = lane & 7;          // n0 in pass 2, k1 in pass 3
    const int jq = lane >> 3;
    float r2[2][8], i2[2][8];
    {
        // a[n0 + 8 n1, 2j..2j+1] at float4 ((n0 + 8 n1) >> 1) * 9 + (n0 & 1) * 4 + j
        const float4* src = buf4 + (lo3 >> 1) * 9 + (lo3 & 1) * 4 + jq;
#pragma unroll
        for (int n1 = 0; n1 < 8; ++n1) {
            const float4 v = src[36 * n1];
            r2[0][n1] = v.x; i2[0][n1] = v.y; r2[1][n1] = v.z; i2[1][n1] = v.w;
        }
    }
    __syncwarp();
#pragma unroll
    for (int s = 0; s < 2; ++s) {
        dft8<false>(r2[s], i2[s]);
#pragma unroll
        for (int k1 = 1; k1 < 8; ++k1) cmul(r2[s][k1], i2[s][k1], tw.w2r[k1 - 1], tw.w2i[k1 - 1]);
    }
    {
        // exchange 2: b[n0, k1; 2j..2j+1] at float4 j*72 + k1*9 + n0
        float4* dst = buf4 + jq * 72 + lo3;
#pragma unroll
        for (int k1 = 0; k1 < 8; ++k1) dst[9 * k1] = make_float4(r2[0][k1], i2[0][k1], r2[1][k1], i2[1][k1]);
    }
    __syncwarp();
    // ---- pass 3: radix-8 over n0 for (k1, k0 in {2j, 2j+1})
    {
        const float4* src = buf4 + jq * 72 + lo3 * 9;
#pragma unroll
        for (int n0 = 0; n0 < 8; ++n0) {
            const float4 v = src[n0];
            r2[0][n0] = v.x; i2[0][n0] = v.y; r2[1][n0] = v.z; i2[1][n0] = v.w;
        }
    }
    __syncwarp();
    dft8<false>(r2[0], i2[0]);
    dft8<false>(r2[1], i2[1]);
    {
        // Z[2j + s + 8 k1 + 64 k2] at float2 10 * (k1 + 8 k2) + 2j + s = float4 5 * (k1 + 8 k2) + j
        float4* dst = buf4 + 5 * lo3 + jq;
#pragma unroll
        for (int k2 = 0; k2 < 8; ++k2) dst[40 * k2] = make_float4(r2[0][k2], i2[0][k2], r2[1][k2], i2[1][k2]);
    }
    __syncwarp();
}

// FOUR TIMES the power spectra of the two packed real frames from Z[k] (a) and Z[N-k] (b):
//   4 |C[k]|^2 = (Zr[k] + Zr[N-k])^2 + (Zi[k] - Zi[N-k])^2
//   4 |D[k]|^2 = (Zi[k] + Zi[N-k])^2 + (Zr[k] - Zr[N-k])^2
// The factor 1/4 is a power of two: callers fold it into a constant they apply anyway (band scale, sqrt),
// which is bit-identical to scaling every bin and saves two multiplies per bin.
constexpr float kPackedPowerScale = 0.25f;
__device__ __forceinline__ void packed_power_pair(float2 a, float2 b, float& pc, float& pd) {
    float sr = a.x + b.x, dr = a.x - b.x;
    float si = a.y + b.y, di = a.y - b.y;
    pc = fmaf(sr, sr, di * di);
    pd = fmaf(si, si, dr * dr);
}

// Lane L gets (4x) the power of the 8 CONSECUTIVE bins k = 8L .. 8L+7 of both packed frames.
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
    // Z[512 - 8L] is the first element of group 64 - L, which lane L - 1 has just loaded as ITS m[0]; lane 0 pairs
    // Z[0] with itself.  Two shuffles instead of a 64-bit load that is two-way bank conflicted (row pitch 20 words).
    float2 b0;
    b0.x = __shfl_up_sync(0xffffffffu, m[0].x, 1);
    b0.y = __shfl_up_sync(0xffffffffu, m[0].y, 1);
    if (lane == 0) b0 = a[0];
    packed_power_pair(a[0], b0, pc[0], pd[0]);
#pragma unroll
    for (int j = 1; j < 8; ++j) packed_power_pair(a[j], m[8 - j], pc[j], pd[j]);
}

// ------------------------------------------------------------------------------------------------
// Band sums over contiguous bin runs (the reference's [49,256] Bark einsum, bark.py:203, and [15,257] third-octave
// bmm, STOI.py:123-125, are 0/1 matrices with disjoint contiguous rows: segment sums, not dense contractions).
//
// Step 1 (every lane, no predicates): segmented inclusive scan of the lane's 8 bins,
//     S[8L + j] = keep[j] * S[8L + j - 1] + p[8L + j],   keep[j] = 0 where bin 8L + j starts a band (and at j = 0),
// written to shared memory as two STS.128 per signal.  S[k] is then the sum of the bins from max(band start, lane
// start) to k.
// Step 2 (the lane that OWNS a band): band sum = S[last bin] + sum over the earlier lanes the band covers of their
// tail S[8l + 7] -- at most kPieces extra loads from offsets precomputed once per kernel; unused pieces point at a
// zero word.  The summation order (ascending bins inside a lane, then head + tails) is fixed: deterministic.
// Row layout of S: bin k at float band_s_index(k) = k + 4 * (k >> 5) -- four pad floats after every 32 bins, so that
// the two STS.128 with which a lane stores its 8 scan values are bank-conflict-free per quarter-warp (8 floats per lane
// without the pad put lanes L and L + 4 on the same banks: measured 8 instead of 4 wavefronts per store).
__host__ __device__ constexpr int band_s_index(int k) { return k + 4 * (k >> 5); }
constexpr int kBandSStride = 300;      // floats per signal row of S (288 used + pad; 300 keeps the gather loads of the
                                       // clean and the degraded row of the third-octave kernel on different banks)
// float index (relative to a row) of that row's zero word: S + kBandZeroWord for the clean row and
// S + kBandSStride + kBandZeroWord for the degraded one, so both rows use the SAME offsets (S spans 3 rows)
constexpr int kBandZeroWord = 2 * kBandSStride;
constexpr int kBandSFloats = 3 * kBandSStride + 4;

struct BandScan {
    float keep[8];
    // starts[] ascending with starts[0] == 0: every bin belongs to a band
    __device__ __forceinline__ void init(const int32_t* starts, int nbands, int lane) {
        unsigned start_mask = 1u;
        for (int b = 0; b < nbands; ++b) {
            const int f = starts[b];
            if ((f >> 3) == lane) start_mask |= 1u << (f & 7);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) keep[j] = ((start_mask >> j) & 1u) ? 0.f : 1.f;
    }
    // scan both power rows and store them: S_c = S, S_d = S + kBandSStride
    __device__ __forceinline__ void scan_store(const float (&pc)[8], const float (&pd)[8], float* S, int lane) const {
        float sc[8], sd[8];
        sc[0] = pc[0]; sd[0] = pd[0];
#pragma unroll
        for (int j = 1; j < 8; ++j) { sc[j] = fmaf(keep[j], sc[j - 1], pc[j]); sd[j] = fmaf(keep[j], sd[j - 1], pd[j]); }
        float4* dc = reinterpret_cast<float4*>(S + band_s_index(8 * lane));
        float4* dd = reinterpret_cast<float4*>(S + kBandSStride + band_s_index(8 * lane));
        dc[0] = make_float4(sc[0], sc[1], sc[2], sc[3]); dc[1] = make_float4(sc[4], sc[5], sc[6], sc[7]);
        dd[0] = make_float4(sd[0], sd[1], sd[2], sd[3]); dd[1] = make_float4(sd[4], sd[5], sd[6], sd[7]);
        if (lane < 2) S[kBandZeroWord + lane * kBandSStride] = 0.f;
    }
};

// Offsets (float index relative to a row of S) of the pieces of ONE band: its last bin and the tails of the earlier
// 8-bin groups it covers; unused pieces (and a lane without a band) point at the row's zero word, so the gather has no
// predicates.  (Recomputing the piece predicates per frame instead of holding kPieces offsets was tried: it did not
// remove the last few spilled registers of the 128-register FFT kernels and made their loops 12-16 instructions longer.)
template <int kPieces>
struct BandGather {
    int last;
    int tail[kPieces];
    __device__ __forceinline__ void init(int first_bin, int end_bin /* exclusive */, bool valid) {
        last = kBandZeroWord;
#pragma unroll
        for (int t = 0; t < kPieces; ++t) tail[t] = kBandZeroWord;
        if (valid && end_bin > first_bin) {
            last = band_s_index(end_bin - 1);
            const int l0 = first_bin >> 3, l1 = (end_bin - 1) >> 3;
#pragma unroll
            for (int t = 0; t < kPieces; ++t)
                if (l0 + t < l1) tail[t] = band_s_index(8 * (l0 + t) + 7);
        }
    }
    // sum of the band in the row starting at `row` (S for clean, S + kBandSStride for degraded)
    __device__ __forceinline__ float sum(const float* row) const {
        float acc = row[last];
#pragma unroll
        for (int t = 0; t < kPieces; ++t) acc += row[tail[t]];
        return acc;
    }
};

}  // namespace fsem
