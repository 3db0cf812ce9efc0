// STOI / ESTOI kernels (sm_100a).  Stream-ordered launches per call:
//
//   stoi_resample_kernel   polyphase sinc-Hann FIR to 10 kHz (base.py:16-21 -> torchaudio Resample);
//                          skipped when the input already is 10 kHz
//   stoi_energy_kernel     Hann-windowed 256/128 frame energies of the CLEAN signal in dB (STOI.py:92-98)
//   stoi_compact_kernel    max energy, -40 dB mask, warp-ballot prefix scan -> kept-frame list, K (STOI.py:102-107)
//   stoi_tob_kernel        STFT frame u straight from kept frames u, u+1, u+2 (the overlap-add buffer of
//                          STOI.py:71-86 is never materialised), clean+degraded packed in one complex
//                          512-pt FFT, 15 third-octave band sums, sqrt (STOI.py:49-69, 121-125)
//   stoi_segment_kernel    sliding 30-frame segments from a shared-memory tile: equalise + clip +
//                          correlate (STOI), row/column normalise + correlate (ESTOI) (STOI.py:129-198)
//   stoi_finalize_kernel   fixed-order sum of tile partials / number of segments
#pragma once
#include "fsem_common.cuh"
#include "fsem_fft.cuh"

namespace fsem {

struct StoiTables {
    float window[FSEM_STOI_WIN];
    int32_t band_lo[FSEM_STOI_NBANDS];
    int32_t band_hi[FSEM_STOI_NBANDS];
    float clip;
    float dyn_range;
};

// ------------------------------------------------------------------------------------------------
// General polyphase resampler: y[neu*k + p] = sum_j taps[p][j] * xpad[orig*k + j],
// xpad = `width` zeros | x | zeros.  One thread per output sample; 1-D grid of blocks_per_sig blocks per signal
// (no 65535 cap of grid.y on the batch).
template <typename T>
__global__ void __launch_bounds__(256)
stoi_resample_kernel(const T* __restrict__ clean, const T* __restrict__ deg,
                     const int32_t* __restrict__ lengths, int64_t batch, int64_t n, int64_t stride,
                     const float* __restrict__ taps, int orig, int neu, int width, int ntaps,
                     float* __restrict__ y, int64_t ystride, int blocks_per_sig) {
    const int64_t sig = blockIdx.x / blocks_per_sig;
    const int bx = blockIdx.x - (int)sig * blocks_per_sig;
    const int64_t item = sig < batch ? sig : sig - batch;
    const int len = item_length(lengths, item, n);
    const int64_t L = stoi_resampled_len(len, orig, neu);
    const int64_t m = (int64_t)bx * blockDim.x + threadIdx.x;
    if (m >= L) return;
    const T* __restrict__ x = (sig < batch ? clean : deg) + item * stride;
    const int64_t k = m / neu;
    const int p = (int)(m - k * neu);
    const float* __restrict__ h = taps + (int64_t)p * ntaps;
    const int64_t i0 = k * orig - width;
    float acc = 0.f;
    for (int j = 0; j < ntaps; ++j) {
        int64_t i = i0 + j;
        float xv = (i >= 0 && i < len) ? load_sample(x + i) : 0.f;
        acc = fmaf(__ldg(h + j), xv, acc);
    }
    y[sig * ystride + m] = acc;
}

// per-item lengths after resampling: ceil(neu * len / orig)
__global__ void __launch_bounds__(256)
resampled_lengths_kernel(const int32_t* __restrict__ lengths, int64_t batch, int64_t n, int orig, int neu,
                         int32_t* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < batch) out[i] = (int32_t)stoi_resampled_len(item_length(lengths, i, n), orig, neu);
}

// ------------------------------------------------------------------------------------------------
// Specialised 16 kHz -> 10 kHz (8:5, half-width 10, 28 taps) resampler.
//
// Thread = FOUR consecutive polyphase blocks (32 new inputs -> 20 outputs) walked with a sliding window of eight float4s
// (two new ones per block), so every staged input sample is read from shared memory 1.9x instead of 3.5x (one block
// per thread) or 2.75x (two blocks, round 1) -- the kernel is bound by the LSU data pipe (91 %), and its registers do not
// grow with the blocks per thread.  The tile is padded by one float4 per eight (lane pitch 9 float4 = conflict-free).  The
// taps arrive as a kernel argument (constant bank: FFMA reads them for free) and only the statically known non-zero range
// of each phase is evaluated (97 of 140 FMAs per block); the host checks that the taps outside those
// ranges are exactly zero before selecting this kernel.
struct Resample85Taps { float h[5][28]; };
#ifndef FSEM_RS85_BLOCKS
#define FSEM_RS85_BLOCKS 4
#endif
constexpr int kRs85Threads = 128;
constexpr int kRs85Nb = FSEM_RS85_BLOCKS;                       // polyphase blocks (8 in -> 5 out) per thread
constexpr int kRs85TileIn = kRs85Threads * 8 * kRs85Nb;         // new input samples per tile (4096)
constexpr int kRs85TileOut = kRs85Threads * 5 * kRs85Nb;        // outputs per tile (2560 = 20 hops)
constexpr int kRs85Quads = kRs85TileIn / 4 + 10;                // float4s staged per tile (12 floats of history + 28 ahead)
constexpr int kRs85ThreadQuads = 2 * kRs85Nb + 6;               // float4s one thread reads (8*NB + 20 inputs, rounded up)
constexpr int kRs85PadEvery = 2 * kRs85Nb;                      // one pad float4 per 2*NB: lane pitch 2*NB+1 (odd) = conflict-free
constexpr int kRs85SmemQuads = kRs85Quads + kRs85Quads / kRs85PadEvery + 2;
__host__ __device__ constexpr int rs85_lo(int p) { return p == 0 ? 1 : p == 1 ? 2 : p == 2 ? 4 : p == 3 ? 6 : 7; }
__host__ __device__ constexpr int rs85_hi(int p) { return p == 0 ? 19 : p == 1 ? 21 : p == 2 ? 22 : p == 3 ? 24 : 26; }

#ifndef FSEM_RS85_TILES
#define FSEM_RS85_TILES 8
#endif
constexpr int kRs85TilesPerCta = FSEM_RS85_TILES;
constexpr int kRs85Hops = kRs85TileOut / FSEM_STOI_HOP;          // hops of 128 output samples per tile
static_assert(kRs85TileOut % FSEM_STOI_HOP == 0, "tiles must hold whole hops");

// Each CTA walks kRs85TilesPerCta consecutive tiles of one signal; the next tile's input is prefetched into
// registers while the current one is computed.  Outputs go through shared memory so that the global stores are
// coalesced float4 and -- for the CLEAN signal -- so that the silent-frame energies can be fused in:
// per hop h of 128 output samples the kernel writes
//     A_h = sum_r (w[r]       * y[128h + r])^2      (hop as FIRST  half of analysis frame h)
//     B_h = sum_r (w[r + 128] * y[128h + r])^2      (hop as SECOND half of analysis frame h-1)
// (products rounded to fp32 like the reference, sums in fp64), so that ||w * x_t||^2 = A_t + B_{t+1} (STOI.py:92-98).
// five CTAs per SM (<= 102 registers): measured 2.99 ms against 3.30 (no cap, 110 registers), 3.13 (four) and 3.62 (six: spills)
#ifndef FSEM_RS85_MINBLOCKS
#define FSEM_RS85_MINBLOCKS 5
#endif
template <bool kVec4, typename T>
__global__ void __launch_bounds__(kRs85Threads, FSEM_RS85_MINBLOCKS)
stoi_resample85_kernel(const T* __restrict__ clean, const T* __restrict__ deg,
                       const int32_t* __restrict__ lengths, int64_t batch, int64_t n, int64_t stride,
                       const __grid_constant__ Resample85Taps taps, const StoiTables* __restrict__ tab,
                       float* __restrict__ y, int64_t ystride, double2* __restrict__ hop_energy, int hops_max,
                       int blocks_per_sig) {
    __shared__ float4 s_in[kRs85SmemQuads];
    __shared__ __align__(16) float s_out[kRs85TileOut];
    __shared__ __align__(16) float s_win[FSEM_STOI_WIN];
    const int tid = threadIdx.x;
    const int64_t sig = blockIdx.x / blocks_per_sig;                 // 1-D grid: no 65535 cap of grid.y on the batch
    const int bx = blockIdx.x - (int)sig * blocks_per_sig;
    const bool is_clean = sig < batch;
    const int64_t item = is_clean ? sig : sig - batch;
    const int len = item_length(lengths, item, n);
    const int L = (int)stoi_resampled_len(len, 8, 5);
    const int tile0 = bx * kRs85TilesPerCta;
    if (tile0 * kRs85TileOut >= L) return;
    const T* __restrict__ x = (is_clean ? clean : deg) + item * stride;
    float* __restrict__ yrow = y + sig * ystride;
    for (int i = tid; i < FSEM_STOI_WIN; i += kRs85Threads) s_win[i] = tab->window[i];

    constexpr int kPerThread = (kRs85Quads + kRs85Threads - 1) / kRs85Threads;   // float4 per thread per tile fill
    // fill role: a warp still moves 32 consecutive float4, but lanes 0-3 / 4-7 of every 8-lane store phase take quads
    // a and a + 4 (a = group of four float4): with one pad float4 per four the padded positions p = 5a + b are then
    // distinct mod 8 and the STS.128 is conflict-free (consecutive lanes would collide on p and p + 8)
    const int fill_q = kRs85PadEvery == 4 ? (tid & ~31) + 4 * (((tid & 31) >> 3) + 4 * ((tid >> 2) & 1)) + (tid & 3) : tid;
    // sample indices fit 32 bits (n < 2^30, checked on the host): 64-bit index arithmetic and compares cost this
    // LSU / issue bound kernel ~15 % of its instructions
    const int len32 = len;
    // the prefetch registers hold the samples as loaded (RawQuad): they are widened to float when the tile is filled
    auto fetch = [&](int tile, RawQuad<T> (&v)[kPerThread]) {
        const int in0 = tile * kRs85TileIn - 12;                                  // first staged sample (multiple of 4)
        if (kVec4 && in0 >= 0 && in0 + 4 * kRs85Quads <= len32) {                 // interior tile (CTA-uniform): no checks
#pragma unroll
            for (int r = 0; r < kPerThread; ++r) {
                const int q = fill_q + r * kRs85Threads;
                if (r + 1 < kPerThread || q < kRs85Quads) v[r] = load_raw4<false>(x + in0 + 4 * q);
                else v[r] = raw_zero<T>();
            }
            return;
        }
#pragma unroll
        for (int r = 0; r < kPerThread; ++r) {
            const int q = fill_q + r * kRs85Threads;
            const int i = in0 + 4 * q;
            if (q >= kRs85Quads) { v[r] = raw_zero<T>(); continue; }
            if (kVec4 && i >= 0 && i + 4 <= len32) {
                v[r] = load_raw4<false>(x + i);
            } else {
                v[r] = raw_pack(load_raw1(x + i, i >= 0 && i < len32), load_raw1(x + i + 1, i + 1 >= 0 && i + 1 < len32),
                                load_raw1(x + i + 2, i + 2 >= 0 && i + 2 < len32),
                                load_raw1(x + i + 3, i + 3 >= 0 && i + 3 < len32), x);
            }
        }
    };
    RawQuad<T> pre[kPerThread];
    fetch(tile0, pre);
    for (int k = 0; k < kRs85TilesPerCta; ++k) {
        const int tile = tile0 + k;
        const int out0 = tile * kRs85TileOut;
        if (out0 >= L) break;                                                      // uniform
#pragma unroll
        for (int r = 0; r < kPerThread; ++r) {
            const int q = fill_q + r * kRs85Threads;
            if (q < kRs85Quads) s_in[q + q / kRs85PadEvery] = raw_to_f4(pre[r]);
        }
        __syncthreads();
        if (k + 1 < kRs85TilesPerCta && (out0 + kRs85TileOut) < L) fetch(tile + 1, pre);
        // thread t: blocks NB*t .. NB*t+NB-1 of this tile need staged floats [8*NB*t + 2, 8*NB*t + 8*NB + 22)
        // block b reads staged floats 8b + 3 .. 8b + 28 = float4s 2b .. 2b + 7: a sliding window of eight float4s, two new
        // ones per block, so the registers do not grow with the blocks per thread (every staged sample is then read
        // (8 NB + 28) / (8 NB) times from shared memory: 2.75x for NB = 2, 1.9x for NB = 4)
        float xin[4 * kRs85ThreadQuads];
        const float4* src = s_in + (kRs85PadEvery + 1) * tid;
        auto load_quad = [&](int i) {
            float4 v = src[i + i / kRs85PadEvery];
            xin[4 * i] = v.x; xin[4 * i + 1] = v.y; xin[4 * i + 2] = v.z; xin[4 * i + 3] = v.w;
        };
#pragma unroll
        for (int i = 0; i < 8; ++i) load_quad(i);
        float out[5 * kRs85Nb];
#pragma unroll
        for (int b = 0; b < kRs85Nb; ++b) {
            if (b + 1 < kRs85Nb) { load_quad(2 * b + 8); load_quad(2 * b + 9); }   // the next block's two new float4s
#pragma unroll
            for (int p = 0; p < 5; ++p) {
                float acc = 0.f;
#pragma unroll
                for (int j = rs85_lo(p); j <= rs85_hi(p); ++j) acc = fmaf(taps.h[p][j], xin[2 + 8 * b + j], acc);
                out[5 * b + p] = acc;
            }
        }
        if (kRs85Nb % 4 == 0) {                                                    // 5*NB floats = whole float4s, lane pitch 5*NB words
#pragma unroll
            for (int i = 0; i < 5 * kRs85Nb / 4; ++i)
                *reinterpret_cast<float4*>(s_out + 5 * kRs85Nb * tid + 4 * i) =
                    make_float4(out[4 * i], out[4 * i + 1], out[4 * i + 2], out[4 * i + 3]);
        } else {
#pragma unroll
            for (int i = 0; i < 5 * kRs85Nb / 2; ++i)
                *reinterpret_cast<float2*>(s_out + 5 * kRs85Nb * tid + 2 * i) = make_float2(out[2 * i], out[2 * i + 1]);
        }
        __syncthreads();
        // coalesced float4 stores of the tile's outputs
        const int valid = min(kRs85TileOut, L - out0);
        if (valid == kRs85TileOut) {                                               // full tile (CTA-uniform): no checks
#pragma unroll
            for (int q = tid; q < kRs85TileOut / 4; q += kRs85Threads)
                *reinterpret_cast<float4*>(yrow + out0 + 4 * q) = *reinterpret_cast<const float4*>(s_out + 4 * q);
        } else {
            for (int q = tid; q < kRs85TileOut / 4; q += kRs85Threads) {
                const float4 v = *reinterpret_cast<const float4*>(s_out + 4 * q);
                float* d = yrow + out0 + 4 * q;
                if (4 * q + 4 <= valid) {
                    *reinterpret_cast<float4*>(d) = v;
                } else {
                    if (4 * q < valid) d[0] = v.x;
                    if (4 * q + 1 < valid) d[1] = v.y;
                    if (4 * q + 2 < valid) d[2] = v.z;
                }
            }
        }
        if (is_clean) {
            // hop energies: warp w takes hops w, w + 4, w + 8 of this tile; a lane handles 4 samples of every hop and holds
            // up to kV = 2 * 3 fp64 partial sums (A and B of three hops).  They are reduced over the 32 lanes by a
            // reduce-scatter (every step halves the number of sums a lane is responsible for) instead of one butterfly
            // per sum: 8 fp64 shuffles per warp instead of 30 -- shuffles share the LSU data pipe this kernel is bound by.
            const int lane = tid & 31, warp = tid >> 5;
            const float4 wa = *reinterpret_cast<const float4*>(s_win + 4 * lane);
            const float4 wb = *reinterpret_cast<const float4*>(s_win + 128 + 4 * lane);
            constexpr int kWarps = kRs85Threads / 32;
            constexpr int kHopsPerWarp = (kRs85Hops + kWarps - 1) / kWarps;        // 3 (two blocks per thread), 5 (four)
            constexpr int kHopGroups = (kHopsPerWarp + 2) / 3;                     // the reduce-scatter takes three hops (six sums)
#pragma unroll 1
            for (int grp = 0; grp < kHopGroups; ++grp) {
            double v[6];
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                const int h = warp + kWarps * (3 * grp + i);
                double a = 0.0, b = 0.0;
                if (h < kRs85Hops && (h + 1) * FSEM_STOI_HOP <= valid) {              // only complete hops matter (warp-uniform)
                    const float4 x = *reinterpret_cast<const float4*>(s_out + h * FSEM_STOI_HOP + 4 * lane);
                    float f;
                    f = __fmul_rn(x.x, wa.x); a = fma((double)f, (double)f, a);
                    f = __fmul_rn(x.y, wa.y); a = fma((double)f, (double)f, a);
                    f = __fmul_rn(x.z, wa.z); a = fma((double)f, (double)f, a);
                    f = __fmul_rn(x.w, wa.w); a = fma((double)f, (double)f, a);
                    f = __fmul_rn(x.x, wb.x); b = fma((double)f, (double)f, b);
                    f = __fmul_rn(x.y, wb.y); b = fma((double)f, (double)f, b);
                    f = __fmul_rn(x.z, wb.z); b = fma((double)f, (double)f, b);
                    f = __fmul_rn(x.w, wb.w); b = fma((double)f, (double)f, b);
                }
                v[2 * i] = a; v[2 * i + 1] = b;
            }
            // step 16: lanes with bit 4 clear keep sums 0..2, the others 3..5
            const bool b16 = lane & 16, b8 = lane & 8, b4 = lane & 4;
            double u[3];
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                const double send = b16 ? v[i] : v[3 + i], keep = b16 ? v[3 + i] : v[i];
                u[i] = keep + __shfl_xor_sync(kFull, send, 16);
            }
            // step 8: bit 3 clear keeps u[0], u[1]; set keeps u[2] (and a zero)
            double t2[2];
            {
                const double s0 = b8 ? u[0] : u[2], k0 = b8 ? u[2] : u[0];
                const double s1 = b8 ? u[1] : 0.0, k1 = b8 ? 0.0 : u[1];
                t2[0] = k0 + __shfl_xor_sync(kFull, s0, 8);
                t2[1] = k1 + __shfl_xor_sync(kFull, s1, 8);
            }
            // step 4: bit 2 clear keeps t2[0], set keeps t2[1]
            double one;
            {
                const double send = b4 ? t2[0] : t2[1], keep = b4 ? t2[1] : t2[0];
                one = keep + __shfl_xor_sync(kFull, send, 4);
            }
            one += __shfl_xor_sync(kFull, one, 2);
            one += __shfl_xor_sync(kFull, one, 1);
            // lane (b16, b8, b4) now holds the total of sum index (b16 ? 3 : 0) + (b8 ? 2 : (b4 ? 1 : 0)); b8 && b4 is the zero
            const int local = b8 ? 2 : (b4 ? 1 : 0);
            const int sum_idx = (b16 ? 3 : 0) + local;
            const int h = warp + kWarps * (3 * grp + (sum_idx >> 1));
            if ((lane & 3) == 0 && !(b8 && b4) && h < kRs85Hops && (h + 1) * FSEM_STOI_HOP <= valid) {
                double* dst = reinterpret_cast<double*>(hop_energy + item * hops_max + tile * kRs85Hops + h);
                dst[sum_idx & 1] = one;                                           // .x = A_h, .y = B_h
            }
            }
        }
        __syncthreads();
    }
}

// E_t = 20*log10(sqrt(A_t + B_{t+1}) + 1e-9) from the fused hop energies (same fp32 op order as stoi_energy_kernel)
__global__ void __launch_bounds__(256)
stoi_energy_from_hops_kernel(const double2* __restrict__ hop_energy, int hops_max, const int32_t* __restrict__ lengths,
                             int64_t batch, int64_t n, int t0max, float* __restrict__ energy) {
    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= batch * (int64_t)t0max) return;
    const int64_t item = gid / t0max;
    const int t = (int)(gid - item * t0max);
    const int64_t L = stoi_resampled_len(item_length(lengths, item, n), 8, 5);
    if (t >= stoi_num_frames(L)) return;
    const double2 h0 = hop_energy[item * hops_max + t];
    const double2 h1 = hop_energy[item * hops_max + t + 1];
    float nrm = (float)sqrt(h0.x + h1.y);
    float v = __fadd_rn(nrm, 1e-9f);
    float lg = (float)log10((double)v);
    energy[gid] = __fmul_rn(20.f, lg);
}

// ------------------------------------------------------------------------------------------------
// Frame energies of the clean 10 kHz signal: E_t = 20*log10(||w * x_t||_2 + 1e-9) (STOI.py:94-98).
// The products w*x are rounded to fp32 as the reference does; the sum of squares is accumulated in
// fp64 (exactly rounded norm), the remaining ops are fp32 in the reference's order.  One warp per frame.
__global__ void __launch_bounds__(256)
stoi_energy_kernel(const float* __restrict__ sig10k, int64_t sstride, const int32_t* __restrict__ lengths,
                   int64_t batch, int64_t n, int orig, int neu, int t0max,
                   const StoiTables* __restrict__ tab, float* __restrict__ energy /* [batch][t0max] */) {
    const int lane = threadIdx.x & 31;
    const int64_t wid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (wid >= batch * (int64_t)t0max) return;
    const int64_t item = wid / t0max;
    const int t = (int)(wid - item * t0max);
    const int64_t L = stoi_resampled_len(item_length(lengths, item, n), orig, neu);
    if (t >= stoi_num_frames(L)) return;
    const float* __restrict__ x = sig10k + item * sstride + (int64_t)t * FSEM_STOI_HOP;
    double acc = 0.0;
#pragma unroll
    for (int m = 0; m < 8; ++m) {
        int q = lane + 32 * m;
        float f = __fmul_rn(__ldg(x + q), tab->window[q]);
        acc = fma((double)f, (double)f, acc);
    }
    acc = warp_sum(acc);
    if (lane == 0) {
        float nrm = (float)sqrt(acc);
        float v = __fadd_rn(nrm, 1e-9f);
        float lg = (float)log10((double)v);
        energy[item * t0max + t] = __fmul_rn(20.f, lg);
    }
}

// ------------------------------------------------------------------------------------------------
// One warp per item: mask_t = ((max E - 40) - E_t) < 0 in fp32 (STOI.py:102), compaction by
// warp ballot + popcount prefix.  Writes the kept-frame list, K and the mask bits.
__global__ void __launch_bounds__(128)
stoi_compact_kernel(const float* __restrict__ energy, const int32_t* __restrict__ lengths, int64_t batch,
                    int64_t n, int orig, int neu, int t0max, int mask_words, float dyn_range,
                    int32_t* __restrict__ kept_idx /* [batch][t0max] */, int32_t* __restrict__ kept_count,
                    uint32_t* __restrict__ mask_bits /* [batch][mask_words] */) {
    const int lane = threadIdx.x & 31;
    const int64_t item = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (item >= batch) return;
    const int64_t L = stoi_resampled_len(item_length(lengths, item, n), orig, neu);
    const int T0 = stoi_num_frames(L);
    const float* __restrict__ e = energy + item * t0max;
    float mx = -INFINITY;
    for (int t = lane; t < T0; t += 32) mx = fmaxf(mx, e[t]);
    mx = warp_max(mx);
    const float thr = __fsub_rn(mx, dyn_range);
    int count = 0;
    for (int t0 = 0; t0 < mask_words * 32; t0 += 32) {
        const int t = t0 + lane;
        bool keep = false;
        if (t < T0) keep = __fsub_rn(thr, e[t]) < 0.f;
        const unsigned bal = __ballot_sync(kFull, keep);
        if (keep) kept_idx[item * t0max + count + __popc(bal & ((1u << lane) - 1u))] = t;
        if (lane == 0) mask_bits[item * mask_words + (t0 >> 5)] = bal;
        count += __popc(bal);
    }
    if (lane == 0) kept_count[item] = count;
}

// ------------------------------------------------------------------------------------------------
// Diagnostics: how close the silent-frame decisions of an item came to flipping.  One warp per item:
// margin = min_t |(max E - dyn_range) - E_t| over the item's frames, in dB (same fp32 ops as the decision above).
__global__ void __launch_bounds__(128)
stoi_margin_kernel(const float* __restrict__ energy, const int32_t* __restrict__ lengths, int64_t batch, int64_t n,
                   int orig, int neu, int t0max, float dyn_range, float* __restrict__ margin_out) {
    const int lane = threadIdx.x & 31;
    const int64_t item = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (item >= batch) return;
    const int64_t L = stoi_resampled_len(item_length(lengths, item, n), orig, neu);
    const int T0 = stoi_num_frames(L);
    const float* __restrict__ e = energy + item * t0max;
    float mx = -INFINITY;
    for (int t = lane; t < T0; t += 32) mx = fmaxf(mx, e[t]);
    mx = warp_max(mx);
    const float thr = __fsub_rn(mx, dyn_range);
    float mg = INFINITY;
    for (int t = lane; t < T0; t += 32) mg = fminf(mg, fabsf(__fsub_rn(thr, e[t])));
    mg = -warp_max(-mg);
    if (lane == 0) margin_out[item] = mg;
}

// ------------------------------------------------------------------------------------------------
// Exclusive prefix sum of the number of STFT frames per item, U_i = max(K_i - 2, 0): prefix[i] = sum_{j<i} U_j,
// prefix[batch] = total.  One CTA; lets the tob kernel hand every warp an equal, contiguous share of REAL frames
// (K is data dependent: about half of the [item, frame] slots are empty on speech).
__global__ void __launch_bounds__(1024)
stoi_prefix_kernel(const int32_t* __restrict__ kept_count, int64_t batch, int32_t* __restrict__ prefix) {
    __shared__ int s_warp[32];
    __shared__ int s_carry;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_carry = 0;
    __syncthreads();
    for (int64_t base = 0; base < batch; base += 1024) {
        const int64_t i = base + tid;
        const int v = (i < batch) ? max(kept_count[i] - 2, 0) : 0;
        int x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int y = __shfl_up_sync(kFull, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) s_warp[warp] = x;
        __syncthreads();
        if (warp == 0) {
            int w = s_warp[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                int y = __shfl_up_sync(kFull, w, o);
                if (lane >= o) w += y;
            }
            s_warp[lane] = w;
        }
        __syncthreads();
        const int carry = s_carry;
        const int incl = carry + x + (warp > 0 ? s_warp[warp - 1] : 0);
        if (i < batch) prefix[i] = incl - v;
        __syncthreads();
        if (tid == 1023) s_carry = incl;
        __syncthreads();
    }
    if (tid == 0) prefix[batch] = s_carry;
}

// ------------------------------------------------------------------------------------------------
// Third-octave spectrogram of the silence-removed signals.  Persistent warps over (item, u).
// STFT frame u of the reference = FFT512 of the 256-sample chunk
//     c_u[q] = w[q] * ( f_{u+1}[q] + (q < 128 ? f_u[q + 128] : f_{u+2}[q - 128]) ),   f_j[q] = w[q] * x[128*t_j + q]
// zero-padded to 512 (torch.stft centre-pads the 256 window; the circular shift does not change the
// power spectrum), where t_j is the j-th kept frame.  Every product/sum is rounded to fp32 like the reference.
#ifndef FSEM_FFT_WARPS
#define FSEM_FFT_WARPS 8
#endif
#ifndef FSEM_FFT_MINBLOCKS
#define FSEM_FFT_MINBLOCKS 2
#endif
constexpr int kTobWarps = FSEM_FFT_WARPS;

// a third-octave band (at most 48 bins, checked by fsem_stoi_create) reaches back into at most 6 earlier lanes
constexpr int kTobPieces = 6;
static_assert(kBandBufFloats <= 2 * kFftBufElems, "band rows alias the FFT exchange buffer");

// kVec2: the 10 kHz rows are 8-byte aligned with an even pitch (always true for the workspace copy the resampler
// writes), so a lane fetches its sample pairs (2L, 2L + 1) with one 64-bit load
template <bool kVec2>
__global__ void __launch_bounds__(kTobWarps * 32, FSEM_FFT_MINBLOCKS)
stoi_tob_kernel(const float* __restrict__ clean10k, const float* __restrict__ deg10k, int64_t sstride,
                int64_t batch, int t0max, int umax,
                int ustride, const int32_t* __restrict__ kept_idx, const int32_t* __restrict__ frame_prefix,
                const StoiTables* __restrict__ tab, float* __restrict__ tob /* [2][batch][15][ustride] */) {
    __shared__ __align__(16) float2 s_buf[kTobWarps][kFftBufElems];
    __shared__ int32_t s_starts[FSEM_STOI_NBANDS + 2];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    float2* buf = s_buf[warp];
    float* wbuf = reinterpret_cast<float*>(buf);
    // pseudo-bands: 0 = bins below the first band, 1..15 = the third-octave bands (contiguous), 16 = bins above
    if (threadIdx.x == 0) s_starts[0] = 0;
    if (threadIdx.x < FSEM_STOI_NBANDS) s_starts[1 + threadIdx.x] = tab->band_lo[threadIdx.x];
    if (threadIdx.x == FSEM_STOI_NBANDS) s_starts[FSEM_STOI_NBANDS + 1] = tab->band_hi[FSEM_STOI_NBANDS - 1];
    __syncthreads();

    FftTwiddles tw;
    tw.init(lane);
    float win[8];                                            // win[4h + j] = w[2*lane + h + 64 j], j = 0..3
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int j = 0; j < 4; ++j) win[4 * h + j] = tab->window[fft_in_index(lane, h, j)];
    FftLaneBins bins;
    bins.init(lane);
    BandScan scan;
    scan.init(s_starts, FSEM_STOI_NBANDS + 2, lane, bins);
    // lanes 0..14 gather the clean bands, lanes 16..30 the degraded ones
    const int my_band = lane & 15;
    const bool has_band = my_band < FSEM_STOI_NBANDS;
    BandGather<kTobPieces> gather;
    gather.init(s_starts[1 + (has_band ? my_band : 0)], s_starts[2 + (has_band ? my_band : 0)], has_band);
    const float* Srow = wbuf + kBandSOffset + (lane >> 4) * kBandSStride;

    // work = the REAL STFT frames of all items in (item, u) order; every warp takes an equal contiguous share
    const int64_t total = frame_prefix[batch];
    const int64_t nwarps = (int64_t)gridDim.x * kTobWarps;
    const int64_t per = (total + nwarps - 1) / nwarps;
    const int64_t w0 = ((int64_t)blockIdx.x * kTobWarps + warp) * per;
    const int64_t w1 = min(total, w0 + per);
    if (w0 >= w1) return;
    // item containing frame w0: largest i with prefix[i] <= w0 (binary search, once per warp)
    int64_t lo_i = 0, hi_i = batch;
    while (hi_i - lo_i > 1) {
        const int64_t mid = (lo_i + hi_i) >> 1;
        if (frame_prefix[mid] <= w0) lo_i = mid; else hi_i = mid;
    }
    int64_t item = lo_i;
    int item_end = frame_prefix[item + 1];                  // first global frame index of the next item
    int u = (int)(w0 - frame_prefix[item]);
    for (int64_t w = w0; w < w1; ++w, ++u) {
        while (w >= item_end) { ++item; item_end = frame_prefix[item + 1]; u = 0; }   // skips items without frames
        {
            const int32_t* idx = kept_idx + item * t0max + u;
            const int ta = idx[0] * FSEM_STOI_HOP, tb = idx[1] * FSEM_STOI_HOP, tc = idx[2] * FSEM_STOI_HOP;
            // Two thirds of a frame's samples were loaded by the previous frame (L1 hits); the new third is kept frame
            // u + 2, whose latency this warp would wait for.  One prefetch instruction (16 lanes = the 16 lines of kept
            // frame u + 3, both signals) brings the NEXT frame's new lines into L1 while this frame is transformed -- no
            // registers, which a register prefetch does not have in this 128-register kernel (tried: 112 bytes of spills).
            if (w + 1 < item_end && lane < 16) {
                const float* p = ((lane & 8) ? deg10k : clean10k) + item * sstride + idx[3] * FSEM_STOI_HOP + 32 * (lane & 7);
                asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
            }
            const float* __restrict__ xc = clean10k + item * sstride + 2 * lane;
            const float* __restrict__ xd = deg10k + item * sstride + 2 * lane;
            auto pair = [&](const float* p) -> float2 {
                if (kVec2) return __ldg(reinterpret_cast<const float2*>(p));
                return make_float2(__ldg(p), __ldg(p + 1));
            };
            float re[16], im[16];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int q = 64 * j;                        // this lane's samples q + 2L, q + 2L + 1 of the chunk
                // neighbour frame: first half of the chunk overlaps the tail of frame u, second half the head of frame u+2
                const int nb = (j < 2) ? (ta + q + 128) : (tc + q - 128);
                const int jn = (j < 2) ? j + 2 : j - 2;      // window index of the neighbour's sample
                const float2 c0 = pair(xc + tb + q), c1 = pair(xc + nb);
                const float2 d0 = pair(xd + tb + q), d1 = pair(xd + nb);
                const float w0f = win[j], w1f = win[4 + j], n0f = win[jn], n1f = win[4 + jn];
                re[j] = __fmul_rn(w0f, __fadd_rn(__fmul_rn(w0f, c0.x), __fmul_rn(n0f, c1.x)));
                re[8 + j] = __fmul_rn(w1f, __fadd_rn(__fmul_rn(w1f, c0.y), __fmul_rn(n1f, c1.y)));
                im[j] = __fmul_rn(w0f, __fadd_rn(__fmul_rn(w0f, d0.x), __fmul_rn(n0f, d1.x)));
                im[8 + j] = __fmul_rn(w1f, __fadd_rn(__fmul_rn(w1f, d0.y), __fmul_rn(n1f, d1.y)));
            }
#pragma unroll
            for (int j = 4; j < 8; ++j) { re[j] = 0.f; im[j] = 0.f; re[8 + j] = 0.f; im[8 + j] = 0.f; }
            float ar[8], ai[8], br[8], bi[8];
            warp_fft512<true>(re, im, buf, tw, lane, ar, ai, br, bi);
            float pc[8], pd[8];
            packed_power_regs(ar, ai, br, bi, lane, pc, pd);
            scan.scan_store(pc, pd, wbuf, lane);
            const float band = gather.sum(Srow);
            if (has_band) {
                const int64_t sig = (lane >> 4) ? (batch + item) : item;
                // sqrt(sum / 4) = sqrt(sum) / 2 exactly (STOI.py:123-125; the packed power spectrum carries a factor 4)
                tob[(sig * FSEM_STOI_NBANDS + my_band) * (int64_t)ustride + u] = 0.5f * sqrtf(band);
            }
            __syncwarp();
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Segment kernel: CTA = (item, tile of kSegTile consecutive segments).
constexpr int kSegTile = 64;
constexpr int kSegThreads = 256;
constexpr int kSegCols = kSegTile + FSEM_STOI_SEG - 1;  // 93 frames feed 64 segments
constexpr int kSegPitch = 97;                            // odd pitch: conflict-free rows
constexpr int kSegTilesPerCta = 4;
#ifndef FSEM_SEG_RUN
#define FSEM_SEG_RUN 4
#endif
constexpr int kSegRun = FSEM_SEG_RUN;                   // consecutive segments of one band per thread in part A
static_assert(kSegTile % kSegRun == 0, "runs tile the segment tile");                       // consecutive tiles per CTA, next tile prefetched into registers

__device__ __forceinline__ float rsqrt_or_zero(float v) { return v > 0.f ? rsqrtf(v) : 0.f; }
// 1 / sqrt(v) for a centred second moment v = s2 - (sum)^2 / N computed from raw sums; 0 when v is not resolved against
// the raw second moment s2 (float32: 30 terms, ~4e-6 relative)
__device__ __forceinline__ float rsqrt_resolved(float v, float s2) { return v > 4e-6f * s2 ? rsqrtf(v) : 0.f; }

// two CTAs per SM by registers; capping them for three or four (with or without the register copy of the rows in
// part A) spills and measured 1.65 / 1.84 ms against 1.63 ms
__global__ void __launch_bounds__(kSegThreads, 2)
stoi_segment_kernel(const float* __restrict__ tob, int64_t batch, int ustride, int ntiles,
                    const int32_t* __restrict__ kept_count, float clip,
                    float2* __restrict__ partial /* [batch][ntiles] (stoi, estoi) */) {
    __shared__ float s_x[FSEM_STOI_NBANDS][kSegPitch];
    __shared__ float s_y[FSEM_STOI_NBANDS][kSegPitch];
    __shared__ float4 s_row[kSegTile][FSEM_STOI_NBANDS];   // (mean_x, 1/||x-mean||, mean_y, 1/||y-mean||)
    __shared__ float s_red[2][kSegThreads / 32];

    const int tid = threadIdx.x;
    const int groups = (ntiles + kSegTilesPerCta - 1) / kSegTilesPerCta;
    const int64_t item = blockIdx.x / groups;
    const int tile_first = (int)(blockIdx.x - item * groups) * kSegTilesPerCta;
    const int K = kept_count[item];
    const int M = max(K - 31, 0);                     // number of segments (STOI.py:183-186)
    const float* __restrict__ tx0 = tob + (item * FSEM_STOI_NBANDS) * (int64_t)ustride;
    const float* __restrict__ ty0 = tob + ((batch + item) * FSEM_STOI_NBANDS) * (int64_t)ustride;
    // the 2 x 15 x 93 tob values of a tile, spread over the CTA's registers (prefetched one tile ahead)
    constexpr int kTileVals = 2 * FSEM_STOI_NBANDS * kSegCols;
    constexpr int kPerThread = (kTileVals + kSegThreads - 1) / kSegThreads;
    auto fetch = [&](int tile, float (&v)[kPerThread]) {
        const int m0f = tile * kSegTile;
        const int ncolf = min(kSegTile, M - m0f) + FSEM_STOI_SEG - 1;
#pragma unroll
        for (int r = 0; r < kPerThread; ++r) {
            const int i = tid + r * kSegThreads;
            const int sg = i / (FSEM_STOI_NBANDS * kSegCols);
            const int rem = i - sg * (FSEM_STOI_NBANDS * kSegCols);
            const int j = rem / kSegCols, c = rem - j * kSegCols;
            const bool ok = i < kTileVals && c < ncolf;
            v[r] = ok ? __ldg((sg ? ty0 : tx0) + (int64_t)j * ustride + m0f + c) : 0.f;
        }
    };
    float pre[kPerThread];
    if (tile_first * kSegTile < M) fetch(tile_first, pre);
    for (int tk = 0; tk < kSegTilesPerCta; ++tk) {
    const int tile = tile_first + tk;
    if (tile >= ntiles) break;
    const int m0 = tile * kSegTile;
    if (m0 >= M) {                                    // uniform: no segments in this tile (nor in the following ones)
        if (tid == 0) partial[item * ntiles + tile] = make_float2(0.f, 0.f);
        continue;
    }
    const int nseg = min(kSegTile, M - m0);
    __syncthreads();                                  // previous tile fully consumed
#pragma unroll
    for (int r = 0; r < kPerThread; ++r) {
        const int i = tid + r * kSegThreads;
        if (i < kTileVals) {
            const int sg = i / (FSEM_STOI_NBANDS * kSegCols);
            const int rem = i - sg * (FSEM_STOI_NBANDS * kSegCols);
            const int j = rem / kSegCols, c = rem - j * kSegCols;
            (sg ? s_y : s_x)[j][c] = pre[r];
        }
    }
    __syncthreads();
    if (tk + 1 < kSegTilesPerCta && tile + 1 < ntiles && (m0 + kSegTile) < M) fetch(tile + 1, pre);

    // ---- part A: per (segment, band) rows: STOI term + row statistics for ESTOI.
    // A thread takes kSegRun CONSECUTIVE segments of one band: their 30-frame rows overlap in all but kSegRun - 1 frames, so
    // the 30 + kSegRun - 1 values are loaded from shared memory once (8 instead of 60 loads per segment) and the raw sums of
    // the frames common to all of them are formed once; every segment adds its own kSegRun - 1 edge frames.  Additions
    // only (no sliding subtraction): an all-zero or constant row still gives exactly zero moments.  The clipped pass
    // depends on the segment's own alpha and is done in full.
    float stoi_acc = 0.f;
    constexpr int kRunsPerBand = kSegTile / kSegRun;
    for (int id = tid; id < FSEM_STOI_NBANDS * kRunsPerBand; id += kSegThreads) {
        const int j = id / kRunsPerBand, mb = (id - j * kRunsPerBand) * kSegRun;
        if (mb >= nseg) continue;
        float x[FSEM_STOI_SEG + kSegRun - 1], y[FSEM_STOI_SEG + kSegRun - 1];
#pragma unroll
        for (int q = 0; q < FSEM_STOI_SEG + kSegRun - 1; ++q) {      // columns beyond the tile's last segment hold zeros
            x[q] = s_x[j][mb + q];
            y[q] = s_y[j][mb + q];
        }
        float cx = 0.f, cy = 0.f, cxx = 0.f, cyy = 0.f;              // frames kSegRun - 1 .. 29: in every segment of the run
#pragma unroll
        for (int q = kSegRun - 1; q < FSEM_STOI_SEG; ++q) {
            cx += x[q]; cy += y[q];
            cxx = fmaf(x[q], x[q], cxx);
            cyy = fmaf(y[q], y[q], cyy);
        }
#pragma unroll
        for (int r = 0; r < kSegRun; ++r) {
            const int m = mb + r;
            if (m < nseg) {
                float sx = cx, sy = cy, sxx = cxx, syy = cyy;
#pragma unroll
                for (int e = 0; e < kSegRun - 1; ++e) {              // edge frames r .. kSegRun - 2 and 30 .. 29 + r
                    const int q = (e < kSegRun - 1 - r) ? r + e : FSEM_STOI_SEG + e - (kSegRun - 1 - r);
                    sx += x[q]; sy += y[q];
                    sxx = fmaf(x[q], x[q], sxx);
                    syy = fmaf(y[q], y[q], syy);
                }
                // equalize_clip (STOI.py:129-139)
                const float alpha = sqrtf(sxx) / (sqrtf(syy) + 1e-9f);
                const float mx = sx * (1.f / FSEM_STOI_SEG), my = sy * (1.f / FSEM_STOI_SEG);
                // Centred second moments from raw sums (sum (x - mx)^2 = sum x^2 - mx * sum x, ...): one pass over the
                // clipped row instead of two, 10 instead of 18 operations per element -- this kernel is issue-bound.  In
                // float32 the formula is less exact than subtracting the mean first, by ~eps * (1 + mean^2 / variance)
                // per row: measured on speech-like and white-noise third-octave rows <= 2e-5 per (segment, band) term
                // and <= 3e-7 in STOI (the bar is 1e-4; the two-pass form gave 7e-8).  Rows whose relative variance is
                // below what the formula can resolve count as constant rows (contribution 0, like exactly constant rows
                // before).
                float syc = 0.f, sycc = 0.f, sxyc = 0.f;
#pragma unroll
                for (int q = 0; q < FSEM_STOI_SEG; ++q) {
                    const float yc = fminf(y[r + q] * alpha, x[r + q] * clip);
                    syc += yc;
                    sycc = fmaf(yc, yc, sycc);
                    sxyc = fmaf(x[r + q], yc, sxyc);
                }
                const float myc = syc * (1.f / FSEM_STOI_SEG);
                const float vxx = fmaf(-sx, mx, sxx), vyo = fmaf(-sy, my, syy);
                const float vyy = fmaf(-syc, myc, sycc), vxy = fmaf(-sx, myc, sxyc);
                const float rx = rsqrt_resolved(vxx, sxx), ry = rsqrt_resolved(vyy, sycc);
                stoi_acc += vxy * rx * ry;             // correlation of the normalised rows (STOI.py:174-175, 190)
                s_row[m][j] = make_float4(mx, rx, my, rsqrt_resolved(vyo, syy));
            }
        }
    }
    __syncthreads();

    // ---- part B: ESTOI: per (segment, time column): rows normalised, then columns (STOI.py:178-181, 193)
    float estoi_acc = 0.f;
    for (int id = tid; id < kSegTile * FSEM_STOI_SEG; id += kSegThreads) {
        const int m = id / FSEM_STOI_SEG, q = id - m * FSEM_STOI_SEG;
        if (m >= nseg) continue;
        float xh[FSEM_STOI_NBANDS], yh[FSEM_STOI_NBANDS];
        float sx = 0.f, sy = 0.f;
#pragma unroll
        for (int j = 0; j < FSEM_STOI_NBANDS; ++j) {
            const float4 r = s_row[m][j];
            xh[j] = (s_x[j][m + q] - r.x) * r.y;
            yh[j] = (s_y[j][m + q] - r.z) * r.w;
            sx += xh[j]; sy += yh[j];
        }
        const float mx = sx * (1.f / FSEM_STOI_NBANDS), my = sy * (1.f / FSEM_STOI_NBANDS);
        float vxx = 0.f, vyy = 0.f, vxy = 0.f;
#pragma unroll
        for (int j = 0; j < FSEM_STOI_NBANDS; ++j) {
            float dx = xh[j] - mx, dy = yh[j] - my;
            vxx = fmaf(dx, dx, vxx);
            vyy = fmaf(dy, dy, vyy);
            vxy = fmaf(dx, dy, vxy);
        }
        estoi_acc += vxy * rsqrt_or_zero(vxx) * rsqrt_or_zero(vyy);
    }
    stoi_acc = warp_sum(stoi_acc);
    estoi_acc = warp_sum(estoi_acc);
    if ((tid & 31) == 0) { s_red[0][tid >> 5] = stoi_acc; s_red[1][tid >> 5] = estoi_acc; }
    __syncthreads();
    if (tid == 0) {
        float a = 0.f, b = 0.f;
        for (int i = 0; i < kSegThreads / 32; ++i) { a += s_red[0][i]; b += s_red[1][i]; }
        partial[item * ntiles + tile] = make_float2(a, b);
    }
    }   // tiles of this CTA
}

__global__ void __launch_bounds__(128)
stoi_finalize_kernel(const float2* __restrict__ partial, int64_t batch, int ntiles,
                     const int32_t* __restrict__ kept_count, float* __restrict__ stoi_out,
                     float* __restrict__ estoi_out, int32_t* __restrict__ kept_out,
                     int32_t* __restrict__ status_out) {
    const int64_t item = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (item >= batch) return;
    const int K = kept_count[item];
    const int M = max(K - 31, 0);
    double a = 0.0, b = 0.0;
    for (int t = 0; t < ntiles; ++t) {
        float2 p = partial[item * ntiles + t];
        a += (double)p.x;
        b += (double)p.y;
    }
    // correlations / normalisation / num_segments (STOI.py:147-151, 198); M = 0 -> 0/0 = NaN like the reference
    const float fm = (float)M;
    stoi_out[item] = (float)(a / FSEM_STOI_NBANDS) / fm;
    estoi_out[item] = (float)(b / FSEM_STOI_SEG) / fm;
    if (kept_out) kept_out[item] = K;
    if (status_out) status_out[item] = (M > 0) ? FSEM_ITEM_OK : FSEM_ITEM_TOO_SHORT;
}

}  // namespace fsem
