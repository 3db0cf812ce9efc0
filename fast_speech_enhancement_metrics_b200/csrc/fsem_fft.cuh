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
constexpr int kFftBufElems = 752;   // 576 for the two exchanges; the band stage reuses it for P and S rows (1504 floats)

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

// Forward 512-point FFT with the spectrum left IN REGISTERS, mirror-paired.
//
// On entry lane L holds z[2L + h + 64 j] in (re[8h + j], im[8h + j]).  On exit lane L holds two output sequences
//     A[k2] = Z[base_a + 64 k2],   B[k2] = Z[base_b + 64 k2],   k2 = 0..7        (FftLaneBins gives base_a, base_b)
// chosen such that the MIRROR bin of A[k2] is B[7 - k2] (and vice versa): Z[512 - (base_a + 64 k2)] = B[7 - k2].
// The two real power spectra of the packed frames (which need Z[k] and Z[512 - k]) are then formed entirely in
// registers -- no third exchange through shared memory (round 1 / early round 2 wrote the spectrum out and read it
// back: 68 of ~270 shared-memory wavefronts per transform, on kernels that are bound by exactly those wavefronts).
// Lane 0 is the exception: its sequences k = 64 k2 and k = 32 + 64 k2 are their own mirrors (A[k2] <-> A[8 - k2],
// B[k2] <-> B[7 - k2]); packed_power_regs handles it with selects.
//
// Index algebra (decimation in frequency), n = n0 + 8 n1 + 64 n2, k = k0 + 8 k1 + 64 k2.  The mirror of (k0, k1, k2)
// is (8 - k0, 7 - k1, 7 - k2) for k0 != 0 and (0, 8 - k1, 7 - k2) for k0 = 0, k1 != 0, so
//   * pass 2 lane (n0 = lane & 7, j = lane >> 3) owns the k0 PAIR {j, 8 - j} (j = 0: {0, 4}),
//   * pass 3 lane (r = lane & 7, j) owns the sequence pair (k0, k1) = (j, r), (8 - j, 7 - r); for j = 0 the pairs are
//     (0, r), (0, 8 - r) [r = 1..3], (0, 0), (0, 4) [r = 0] and (4, 7 - r), (4, r) [r = 4..7] -- lanes 0..7 replace
//     elements of the stored float4s with selects so that the stores are the same instruction for all lanes.
// Exchange layouts (all accesses 128-bit, conflict-free per quarter-warp -- checked by emulation and with ncu):
//   1: row n' = 2L + h at float4 9 L + 4 h; float4 q of a row = (a[n', k0 = q], a[n', k0 = (8 - q) or 4])
//   2: float4 j * 72 + r * 9 + n0 = (b[n0; A-sequence of reader (r, j)], b[n0; B-sequence])
// `buf`: this warp's private buffer of kFftBufElems float2 (16-byte aligned); the transform uses the first 576.
struct FftLaneBins {
    int base_a, base_b;
    __device__ __forceinline__ void init(int lane) {
        const int r = lane & 7, j = lane >> 3;
        if (j >= 1) { base_a = j + 8 * r; base_b = (8 - j) + 8 * (7 - r); }
        else if (r == 0) { base_a = 0; base_b = 32; }
        else if (r < 4) { base_a = 8 * r; base_b = 8 * (8 - r); }
        else { base_a = 4 + 8 * (7 - r); base_b = 4 + 8 * r; }
    }
};

template <bool kUpperZero>
__device__ __forceinline__ void warp_fft512(float (&re)[16], float (&im)[16], float2* buf, const FftTwiddles& tw, int lane,
                                            float (&ar)[8], float (&ai)[8], float (&br)[8], float (&bi)[8]) {
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
        float4* row = buf4 + 9 * lane + 4 * h;
        row[0] = make_float4(r[0], i[0], r[4], i[4]);
        row[1] = make_float4(r[1], i[1], r[7], i[7]);
        row[2] = make_float4(r[2], i[2], r[6], i[6]);
        row[3] = make_float4(r[3], i[3], r[5], i[5]);
    }
    __syncwarp();
    // ---- pass 2: radix-8 over n1 for n0 = lane & 7 and the k0 pair of group j = lane >> 3
    const int lo3 = lane & 7;          // n0 in pass 2, r in pass 3
    const int jq = lane >> 3;
    float r2[2][8], i2[2][8];
    {
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
        // r2[0][k1] = b[k1; k0 = j], r2[1][k1] = b[k1; k0 = 8 - j] (j = 0: k0 = 0 and 4).  Reader (r, j) wants
        // (b[r; j], b[7 - r; 8 - j]) = (r2[0][r], r2[1][7 - r]); group 0 pairs within one k0 (see above): one element of
        // every float4 r < 4 and both of every float4 r >= 4 of lanes 0..7 are replaced with selects (24 selects).
        const bool g0 = jq == 0;
        float4* dst = buf4 + jq * 72 + lo3;
        // float4 r < 4: (b[r; j], mirror) -- group 0 takes the mirror (0, 8 - r) [r = 0: (0, 4)] from its own first sequence
        dst[0] = make_float4(r2[0][0], i2[0][0], g0 ? r2[0][4] : r2[1][7], g0 ? i2[0][4] : i2[1][7]);
        dst[9] = make_float4(r2[0][1], i2[0][1], g0 ? r2[0][7] : r2[1][6], g0 ? i2[0][7] : i2[1][6]);
        dst[18] = make_float4(r2[0][2], i2[0][2], g0 ? r2[0][6] : r2[1][5], g0 ? i2[0][6] : i2[1][5]);
        dst[27] = make_float4(r2[0][3], i2[0][3], g0 ? r2[0][5] : r2[1][4], g0 ? i2[0][5] : i2[1][4]);
        // float4 r >= 4: (b[r; j], b[7 - r; 8 - j]) -- group 0 stores ((4, 7 - r), (4, r)): this order keeps the scattered
        // power stores of the band stage bank-conflict-free (bins 4 + 8 (7 - r) = 28, 20, 12, 4 in the A slot)
#pragma unroll
        for (int r = 4; r < 8; ++r)
            dst[9 * r] = make_float4(g0 ? r2[1][7 - r] : r2[0][r], g0 ? i2[1][7 - r] : i2[0][r],
                                     g0 ? r2[1][r] : r2[1][7 - r], g0 ? i2[1][r] : i2[1][7 - r]);
    }
    __syncwarp();
    // ---- pass 3: radix-8 over n0 for the two sequences of reader (r = lane & 7, j)
    {
        const float4* src = buf4 + jq * 72 + lo3 * 9;
#pragma unroll
        for (int n0 = 0; n0 < 8; ++n0) {
            const float4 v = src[n0];
            ar[n0] = v.x; ai[n0] = v.y; br[n0] = v.z; bi[n0] = v.w;
        }
    }
    __syncwarp();                      // every lane has read exchange 2: the buffer may be reused by the caller
    dft8<false>(ar, ai);
    dft8<false>(br, bi);
}

// FOUR TIMES the power spectra of the two packed real frames from Z[k] (a) and Z[N-k] (b):
//   4 |C[k]|^2 = (Zr[k] + Zr[N-k])^2 + (Zi[k] - Zi[N-k])^2
//   4 |D[k]|^2 = (Zi[k] + Zi[N-k])^2 + (Zr[k] - Zr[N-k])^2
// The factor 1/4 is a power of two: callers fold it into a constant they apply anyway (band scale, sqrt),
// which is bit-identical to scaling every bin and saves two multiplies per bin.
constexpr float kPackedPowerScale = 0.25f;
__device__ __forceinline__ void packed_power_pair(float zr, float zi, float mr, float mi, float& pc, float& pd) {
    float sr = zr + mr, dr = zr - mr;
    float si = zi + mi, di = zi - mi;
    pc = fmaf(sr, sr, di * di);
    pd = fmaf(si, si, dr * dr);
}

// (4x) power of the lane's 8 bins below 256: slot q < 4 is bin base_a + 64 q (sequence A), slot 4 + q is bin
// base_b + 64 q (sequence B).  Generic lanes pair A[q] with B[7 - q] and B[q] with A[7 - q]; lane 0 pairs A[q] with
// A[(8 - q) & 7] (bins 0, 64, 128, 192; bin 0 with itself) and B[q] with B[7 - q] (bins 32, 96, 160, 224).
__device__ __forceinline__ void packed_power_regs(const float (&ar)[8], const float (&ai)[8], const float (&br)[8],
                                                  const float (&bi)[8], int lane, float (&pc)[8], float (&pd)[8]) {
    const bool l0 = lane == 0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const float mar = l0 ? ar[(8 - q) & 7] : br[7 - q], mai = l0 ? ai[(8 - q) & 7] : bi[7 - q];
        const float mbr = l0 ? br[7 - q] : ar[7 - q],       mbi = l0 ? bi[7 - q] : ai[7 - q];
        packed_power_pair(ar[q], ai[q], mar, mai, pc[q], pd[q]);
        packed_power_pair(br[q], bi[q], mbr, mbi, pc[4 + q], pd[4 + q]);
    }
}

// ------------------------------------------------------------------------------------------------
// Band sums over contiguous bin runs (the reference's [49,256] Bark einsum, bark.py:203, and [15,257] third-octave
// bmm, STOI.py:123-125, are 0/1 matrices with disjoint contiguous rows: segment sums, not dense contractions).
//
// Step 0: the power spectrum leaves the FFT scattered over the lanes (8 bins 64 apart per lane); it is transposed through
// the natural-order rows P (sixteen 32-bit stores, four 128-bit loads per lane, conflict-free).
// Step 1 (every lane, no predicates): segmented inclusive scan of the lane's 8 CONSECUTIVE bins,
//     S[8L + j] = keep[j] * S[8L + j - 1] + p[8L + j],   keep[j] = 0 where bin 8L + j starts a band (and at j = 0),
// written to shared memory as two STS.128 per signal.  S[k] is then the sum of the bins from max(band start, lane
// start) to k.
// Step 2 (the lane that OWNS a band): band sum = S[last bin] + sum over the earlier lanes the band covers of their
// tail S[8l + 7] -- at most kPieces extra loads from offsets precomputed once per kernel; unused pieces point at a
// zero word.  The summation order (ascending bins inside a lane, then head + tails) is fixed: deterministic.
// Row layout of P and S: bin k at float band_s_index(k) = k + 4 * (k >> 5) -- four pad floats after every 32 bins, so
// that (i) the 32-bit stores of the lane-scattered powers (bins base + 64 q) and (ii) the two 128-bit accesses with
// which a lane loads its 8 consecutive bins / stores its 8 scan values are all bank-conflict-free (8 floats per lane
// without the pad put lanes L and L + 4 on the same banks: measured 8 instead of 4 wavefronts per access).
__host__ __device__ constexpr int band_s_index(int k) { return k + 4 * (k >> 5); }
constexpr int kBandSStride = 300;      // floats per row (288 used + pad; 300 keeps the gather loads of the clean and the
                                       // degraded row of the third-octave kernel on different banks)
// Per-warp layout (floats, relative to the warp's buffer): P rows (natural-order powers) at 0 and 300, S rows (scans) at
// 600 and 900, and one zero word per S row at row + kBandZeroWord, so both S rows use the SAME gather offsets.
constexpr int kBandPOffset = 0;
constexpr int kBandSOffset = 2 * kBandSStride;
constexpr int kBandZeroWord = 2 * kBandSStride;          // relative to an S row
constexpr int kBandBufFloats = kBandSOffset + 3 * kBandSStride + 4;

struct BandScan {
    float keep[8];
    int pa, pb;        // float offsets (relative to a P row) of this lane's bins base_a, base_b
    // starts[] ascending with starts[0] == 0: every bin belongs to a band
    __device__ __forceinline__ void init(const int32_t* starts, int nbands, int lane, const FftLaneBins& bins) {
        unsigned start_mask = 1u;
        for (int b = 0; b < nbands; ++b) {
            const int f = starts[b];
            if ((f >> 3) == lane) start_mask |= 1u << (f & 7);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) keep[j] = ((start_mask >> j) & 1u) ? 0.f : 1.f;
        pa = band_s_index(bins.base_a);
        pb = band_s_index(bins.base_b);
    }
    // pc / pd: the lane's 8 scattered bins (packed_power_regs order).  Transposes them through the P rows, scans the
    // lane's 8 CONSECUTIVE bins and stores the scans in the S rows.  band_s_index(base + 64 q) = band_s_index(base) + 72 q.
    // `w` = the warp's buffer as floats.  Ends with __syncwarp(): S is ready for the gathers.
    __device__ __forceinline__ void scan_store(const float (&pc)[8], const float (&pd)[8], float* w, int lane) const {
        float* Pc = w + kBandPOffset;
        float* Pd = Pc + kBandSStride;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            Pc[pa + 72 * q] = pc[q];      Pd[pa + 72 * q] = pd[q];
            Pc[pb + 72 * q] = pc[4 + q];  Pd[pb + 72 * q] = pd[4 + q];
        }
        __syncwarp();
        const int mine = band_s_index(8 * lane);
        const float4 c0 = *reinterpret_cast<const float4*>(Pc + mine), c1 = *reinterpret_cast<const float4*>(Pc + mine + 4);
        const float4 d0 = *reinterpret_cast<const float4*>(Pd + mine), d1 = *reinterpret_cast<const float4*>(Pd + mine + 4);
        const float xc[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
        const float xd[8] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
        float sc[8], sd[8];
        sc[0] = xc[0]; sd[0] = xd[0];
#pragma unroll
        for (int j = 1; j < 8; ++j) { sc[j] = fmaf(keep[j], sc[j - 1], xc[j]); sd[j] = fmaf(keep[j], sd[j - 1], xd[j]); }
        float* S = w + kBandSOffset;
        float4* dc = reinterpret_cast<float4*>(S + mine);
        float4* dd = reinterpret_cast<float4*>(S + kBandSStride + mine);
        dc[0] = make_float4(sc[0], sc[1], sc[2], sc[3]); dc[1] = make_float4(sc[4], sc[5], sc[6], sc[7]);
        dd[0] = make_float4(sd[0], sd[1], sd[2], sd[3]); dd[1] = make_float4(sd[4], sd[5], sd[6], sd[7]);
        if (lane < 2) S[kBandZeroWord + lane * kBandSStride] = 0.f;
        __syncwarp();
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
