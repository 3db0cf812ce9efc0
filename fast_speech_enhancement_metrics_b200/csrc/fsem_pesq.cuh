// PESQ kernels (sm_100a).  Three stream-ordered launches per call:
//
//   pesq_filter_kernel   x -> sum(y^2) partials (10th-order band-pass, parallel form) and the
//                        tapered + pre-emphasised signal z          (PESQ.py:92-113)
//   pesq_spectrum_kernel z -> Hann-512/256 power spectra of (clean, degraded) packed in ONE
//                        complex FFT per frame -> 49 Bark-band power densities (PESQ.py:123-140,
//                        bark.py:203-204)
//   pesq_bark_kernel     level alignment gain, band/frame equalisation, Zwicker loudness,
//                        symmetric/asymmetric disturbance, L6/L2 aggregation, MOS mapping
//                        (PESQ.py:142-245, loudness.py:48-67, bark.py:169-184)
//
// The level-alignment gain g = sqrt(1e7 / P) needs the whole-signal band power, but everything up
// to the power spectrum is linear, so the filters/FFT run on the UNSCALED signal and the Bark
// powers are multiplied by g^2 in the last kernel.  equalize_ranges (PESQ.py:115-121) cancels
// algebraically against the same normalisation; its only observable effect (NaN for an all-zero
// item) is reproduced through g^2 = inf.
#pragma once
#include "fsem_common.cuh"
#include "fsem_fft.cuh"

namespace fsem {

// ------------------------------------------------------------------------------------------------
// device-side constant tables (one copy per context, in global memory, read through L1/constant path)
struct PesqTables {
    float hann[FSEM_PESQ_NFFT];
    int32_t band_first[FSEM_PESQ_NBANDS];
    int32_t band_count[FSEM_PESQ_NBANDS];
    float pow_dens[FSEM_PESQ_NBANDS];
    float thresh[FSEM_PESQ_NBANDS];
    float zw_exp[FSEM_PESQ_NBANDS];
    float width[FSEM_PESQ_NBANDS];
    float loud_scale[FSEM_PESQ_NBANDS];  // Sl * (2*thresh)^exp
    float width_total;                   // sum_{b>=1} width
};

// filter coefficients travel as a kernel argument (constant bank)
struct PesqFilterCoef {
    float k;
    float c0[FSEM_BP_SECTIONS], c1[FSEM_BP_SECTIONS], a1[FSEM_BP_SECTIONS], a2[FSEM_BP_SECTIONS];
    float pb0, pb1, pb2, pa1, pa2;
};

struct IirState {
    float w1[FSEM_BP_SECTIONS], w2[FSEM_BP_SECTIONS];
    float s1, s2;
};

// one sample of the band-pass (parallel form, 22 flop-instructions): returns y
__device__ __forceinline__ float bandpass_step(const PesqFilterCoef& P, IirState& st, float x) {
    float acc_a = P.k * x;
    float acc_b = 0.f;
#pragma unroll
    for (int s = 0; s < FSEM_BP_SECTIONS; ++s) {
        float w = fmaf(-P.a1[s], st.w1[s], fmaf(-P.a2[s], st.w2[s], x));
        if (s < 3) {
            acc_a = fmaf(P.c0[s], w, acc_a);
            acc_a = fmaf(P.c1[s], st.w1[s], acc_a);
        } else {
            acc_b = fmaf(P.c0[s], w, acc_b);
            acc_b = fmaf(P.c1[s], st.w1[s], acc_b);
        }
        st.w2[s] = st.w1[s];
        st.w1[s] = w;
    }
    return acc_a + acc_b;
}

// one sample of the pre-emphasis biquad (transposed direct form II): returns z
__device__ __forceinline__ float preemph_step(const PesqFilterCoef& P, IirState& st, float x) {
    float z = fmaf(P.pb0, x, st.s1);
    st.s1 = fmaf(-P.pa1, z, fmaf(P.pb1, x, st.s2));
    st.s2 = fmaf(-P.pa2, z, P.pb2 * x);
    return z;
}

// taper weight of sample t in a signal of `len` samples (PESQ.py:108-109); 1 in the interior
__device__ __forceinline__ float taper_weight(int t, int len) {
    float w = 1.f;
    if (t < 15) w *= (float)(t + 1) * 0.0625f;
    if (t >= len - 15) w *= (float)(len - t) * 0.0625f;
    return w;
}

// ------------------------------------------------------------------------------------------------
// Kernel A: one thread per (signal, time chunk).  The thread starts `warm` samples before its
// chunk with zero state (both IIRs have decayed below fp32 resolution by then), walks the chunk
// serially, accumulates y^2 and writes z.  Signals: s < B -> clean[s], else deg[s - B].
template <bool kVec4, typename T>
__global__ void __launch_bounds__(128)
pesq_filter_kernel(const T* __restrict__ clean, const T* __restrict__ deg,
                   const int32_t* __restrict__ lengths, int64_t batch, int64_t n, int64_t stride,
                   int chunk, int nchunks, int warm, const __grid_constant__ PesqFilterCoef P,
                   float* __restrict__ z_out, int64_t zstride, double* __restrict__ partial) {
    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= 2 * batch * nchunks) return;
    const int64_t sig = gid / nchunks;
    const int c = (int)(gid - sig * nchunks);
    const int64_t item = sig < batch ? sig : sig - batch;
    const T* __restrict__ x = (sig < batch ? clean : deg) + item * stride;
    float* __restrict__ z = z_out + sig * zstride;
    const int len = item_length(lengths, item, n);

    const int t_acc = c * chunk;                       // first sample owned by this thread
    if (t_acc >= len) { partial[gid] = 0.0; return; }
    const int t_end = min(len, t_acc + chunk);
    int t = max(0, t_acc - warm);                       // chunk and warm are multiples of 4

    IirState st;
#pragma unroll
    for (int s = 0; s < FSEM_BP_SECTIONS; ++s) { st.w1[s] = 0.f; st.w2[s] = 0.f; }
    st.s1 = 0.f; st.s2 = 0.f;

    auto load4 = [&](int tt, float (&v)[4]) {
        if (kVec4 && tt + 4 <= len) {
            float4 q = load_samples4<false>(x + tt);
            v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) v[j] = (tt + j < len) ? load_sample(x + tt + j) : 0.f;
        }
    };

    // Blocks of 32 samples (8 groups of four): warm-up groups (t < t_acc) only advance the filter states, owned groups
    // also accumulate y^2 and write z.  The next block is loaded before the current one is filtered: this kernel
    // serves tiny batches, where nothing else hides the load latency (0.068 -> 0.060 ms for <= 32 items x 10 s; the rest
    // is the single warp per scheduler issuing ~37 instructions per sample over chunk + warm-up = 896 serial samples).
    double acc_d = 0.0;
    float acc = 0.f;
    int groups = 0;
    float cur[8][4], nxt[8][4];
    auto load_block = [&](int tb, float (&blk)[8][4]) {
#pragma unroll
        for (int g = 0; g < 8; ++g) load4(tb + 4 * g, blk[g]);     // groups beyond len / t_end read as zeros
    };
    load_block(t, cur);                                            // t is a multiple of 32 (chunk % 64 == 0, warm % 32 == 0)
    for (; t < t_end; t += 32) {
        if (t + 32 < t_end) load_block(t + 32, nxt);
#pragma unroll
        for (int g = 0; g < 8; ++g) {
            const int tg = t + 4 * g;
            if (tg >= t_end) break;
            const bool edge = (tg < 16) || (tg + 4 > len - 16);
            float zz[4], yy[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                yy[j] = bandpass_step(P, st, cur[g][j]);
                float xv = edge ? cur[g][j] * taper_weight(tg + j, len) : cur[g][j];
                zz[j] = preemph_step(P, st, xv);
            }
            if (tg >= t_acc) {
                if (tg + 4 <= t_end) {
                    acc += (yy[0] * yy[0] + yy[1] * yy[1]) + (yy[2] * yy[2] + yy[3] * yy[3]);
                    if (kVec4) {
                        *reinterpret_cast<float4*>(z + tg) = make_float4(zz[0], zz[1], zz[2], zz[3]);
                    } else {
#pragma unroll
                        for (int j = 0; j < 4; ++j) z[tg + j] = zz[j];
                    }
                } else {  // ragged tail of the signal: only samples < len count
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        if (tg + j < t_end) { acc = fmaf(yy[j], yy[j], acc); z[tg + j] = zz[j]; }
                    }
                }
                if (++groups == 16) { acc_d += (double)acc; acc = 0.f; groups = 0; }
            }
        }
#pragma unroll
        for (int g = 0; g < 8; ++g)
#pragma unroll
            for (int j = 0; j < 4; ++j) cur[g][j] = nxt[g][j];
    }
    partial[gid] = acc_d + (double)acc;
}

// ------------------------------------------------------------------------------------------------
// Kernel A, tiled variant (16-byte aligned rows): one warp = 32 SIGNALS x one time chunk, every lane
// walking its own signal serially.  Global traffic is staged through a warp-private shared-memory tile of
// 32 rows x 32 samples so that every LDG/STG is a coalesced 128-byte row segment (the direct variant above
// issues 32 scattered 16-byte accesses per instruction and is bound by L1 wavefronts):
//   fill   : 8 x (LDG.128 coalesced -> registers), then 8 x STS.128          (rows = signals)
//   compute: lane l reads row l (8 x LDS.128, pitch 36 floats = conflict-free), runs both IIRs on 32
//            samples, overwrites the row with z in place
//   drain  : 8 x (LDS.128 -> STG.128 coalesced)
// The next tile's loads are issued before the current tile is computed (register prefetch).
// DRAM page locality: the 32 rows of a warp are 640 kB apart, so 128 bytes per row and step would open a DRAM page
// per row and step in both directions.  Reads carry the L2::256B hint (every other step hits L2); z is kept for two
// steps (two tile slots per row) and leaves as two back-to-back 128-byte stores per row, so L2 holds -- and later
// writes back -- whole 256-byte pieces.  Together 4.45 -> 4.0 ms at 8192 x 10 s.
constexpr int kFiltWarps = 4;
#ifndef FSEM_FILTER_DRAIN32
constexpr int kFiltPitch = 68;   // floats per tile row: two 32-sample slots + 4 pad
#else
constexpr int kFiltPitch = 36;   // A/B: one slot, z drained after every tile (4.23 vs 4.00 ms at 8192 x 10 s)
#endif

// Variable-length batches (kHasLengths): `order` is a permutation that puts items of similar length next to each
// other (pesq_order_kernel), so the 32 signals of a warp end together; a warp stops at its longest signal and a lane
// whose signal has ended skips the arithmetic.  Results do not depend on the grouping: every lane runs the same
// recurrence on the same chunk grid whatever its neighbours are.
template <bool kHasLengths, typename T>
__global__ void __launch_bounds__(kFiltWarps * 32)
pesq_filter_tiled_kernel(const T* __restrict__ clean, const T* __restrict__ deg,
                         const int32_t* __restrict__ lengths, const int32_t* __restrict__ order, int64_t batch,
                         int64_t n, int64_t stride, int chunk, int nchunks, int warm,
                         const __grid_constant__ PesqFilterCoef P, float* __restrict__ z_out, int64_t zstride,
                         double* __restrict__ partial) {
    __shared__ __align__(16) float s_tile[kFiltWarps][32 * kFiltPitch];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    float* tile = s_tile[warp];
    // unit = (clean | degraded, group of 32 items, chunk): a warp never straddles the two tensors
    const int64_t groups = ceil_div(batch, 32);
    const int64_t unit = (int64_t)blockIdx.x * kFiltWarps + warp;
    if (unit >= 2 * groups * nchunks) return;
    const int half = (int)(unit / (groups * nchunks));
    const int64_t rem = unit - (int64_t)half * groups * nchunks;
    const int64_t grp = rem / nchunks;
    const int c = (int)(rem - grp * nchunks);
    const int64_t row0 = grp * 32;

    // own signal (compute role): slot row0 + lane of the (possibly permuted) batch
    const bool sig_ok = row0 + lane < batch;
    const int64_t my_item = !sig_ok ? 0 : (kHasLengths && order != nullptr) ? (int64_t)order[row0 + lane] : row0 + lane;
    const int len = sig_ok ? item_length(lengths, my_item, n) : 0;
    // rows this lane helps to move (transfer role): slots (lane >> 3) + 4*i, float4 column lane & 7
    const int col = (lane & 7) * 4;
    const int64_t trow = row0 + (lane >> 3);
    const T* __restrict__ sbase = (half ? deg : clean) + col;
    float* __restrict__ dbase = z_out + (int64_t)half * batch * zstride + col;
    const T* __restrict__ src0 = sbase + trow * stride;
    float* __restrict__ dst0 = dbase + trow * zstride;
    const int64_t sstep = 4 * stride, dstep = 4 * zstride;
    int row_len[kHasLengths ? 8 : 1];
    int row_item[kHasLengths ? 8 : 1];
    if (kHasLengths) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int64_t slot = trow + 4 * i;
            const int64_t it = slot >= batch ? 0 : (order != nullptr ? (int64_t)order[slot] : slot);
            row_item[i] = (int)it;
            row_len[i] = slot < batch ? item_length(lengths, it, n) : 0;
        }
    }
    const int rows_ok = (int)min((int64_t)8, (batch - trow + 3) / 4);      // slots trow + 4i < batch  <=>  i < rows_ok
    auto rlen = [&](int i) -> int { return kHasLengths ? row_len[kHasLengths ? i : 0] : (i < rows_ok ? (int)n : 0); };
    auto src_row = [&](int i) -> const T* {
        return kHasLengths ? sbase + (int64_t)row_item[kHasLengths ? i : 0] * stride : src0 + i * sstep;
    };
    auto dst_row = [&](int i) -> float* {
        return kHasLengths ? dbase + (int64_t)row_item[kHasLengths ? i : 0] * zstride : dst0 + i * dstep;
    };

    const int t_acc = c * chunk;
    // common upper bound of the chunk: nothing to do beyond the longest of the warp's 32 signals
    const int grp_len = kHasLengths ? __reduce_max_sync(kFull, len) : (int)n;
    const int t_stop = min(grp_len, t_acc + chunk);
    const int t_end = min(len, t_stop);                      // this lane's own bound
    int t = max(0, t_acc - warm);                            // multiple of 32
    if (t_acc >= grp_len) t = t_stop;                        // the whole chunk lies beyond every signal: no warm-up either

    IirState st;
#pragma unroll
    for (int s = 0; s < FSEM_BP_SECTIONS; ++s) { st.w1[s] = 0.f; st.w2[s] = 0.f; }
    st.s1 = 0.f; st.s2 = 0.f;
    double acc_d = 0.0;

    // the prefetch registers hold the samples as loaded (RawQuad): they are widened to float when the tile is filled
    auto fetch = [&](int tt, RawQuad<T> (&v)[8]) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int a = tt + col;
            const int rl = rlen(i);
            const T* q = src_row(i) + tt;
            if (a + 4 <= rl) {
                v[i] = load_raw4<true>(q);         // L2::256B hint: 4.46 -> 4.18 ms at 8192 x 10 s against a plain __ldg
            } else {
                v[i] = raw_pack(load_raw1(q, a < rl), load_raw1(q + 1, a + 1 < rl), load_raw1(q + 2, a + 2 < rl),
                                load_raw1(q + 3, a + 3 < rl), q);
            }
        }
    };
    RawQuad<T> pre[8];
    if (t < t_stop) fetch(t, pre);
    for (; t < t_stop; t += 32) {
        // fill the tile with the prefetched segment, then prefetch the next one
#ifndef FSEM_FILTER_DRAIN32
        const int cs = ((t >> 5) & 1) * 32;                  // slot of this tile
#else
        constexpr int cs = 0;
#endif
#pragma unroll
        for (int i = 0; i < 8; ++i)
            *reinterpret_cast<float4*>(tile + ((lane >> 3) + 4 * i) * kFiltPitch + cs + col) = raw_to_f4(pre[i]);
        __syncwarp();
        if (t + 32 < t_stop) fetch(t + 32, pre);
        // compute: lane owns row `lane`
        float* row = tile + lane * kFiltPitch + cs;
        const bool owned = t >= t_acc;
        float acc = 0.f;
        if (t >= 16 && t + 48 <= len) {
            // interior tile: no taper, every sample counts -- no per-sample predicates
            float acc2 = 0.f;
#pragma unroll
            for (int g = 0; g < 8; ++g) {
                float4 q = *reinterpret_cast<const float4*>(row + 4 * g);
                float y0 = bandpass_step(P, st, q.x); float z0 = preemph_step(P, st, q.x);
                float y1 = bandpass_step(P, st, q.y); float z1 = preemph_step(P, st, q.y);
                float y2 = bandpass_step(P, st, q.z); float z2 = preemph_step(P, st, q.z);
                float y3 = bandpass_step(P, st, q.w); float z3 = preemph_step(P, st, q.w);
                acc = fmaf(y0, y0, acc); acc2 = fmaf(y1, y1, acc2);
                acc = fmaf(y2, y2, acc); acc2 = fmaf(y3, y3, acc2);
                *reinterpret_cast<float4*>(row + 4 * g) = make_float4(z0, z1, z2, z3);
            }
            acc += acc2;
        } else if (t < len) {                                // a lane whose signal has ended has nothing left to compute
#pragma unroll
            for (int g = 0; g < 8; ++g) {
                float4 q = *reinterpret_cast<const float4*>(row + 4 * g);
                float v[4] = {q.x, q.y, q.z, q.w};
                float zz[4];
                const int tg = t + 4 * g;
                const bool edge = (tg < 16) || (tg + 4 > len - 16);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    float yv = bandpass_step(P, st, v[j]);
                    float xv = edge ? v[j] * taper_weight(tg + j, len) : v[j];
                    zz[j] = preemph_step(P, st, xv);
                    if (tg + j < t_end) acc = fmaf(yv, yv, acc);
                }
                *reinterpret_cast<float4*>(row + 4 * g) = make_float4(zz[0], zz[1], zz[2], zz[3]);
            }
        }
        if (owned) acc_d += (double)acc;
        __syncwarp();
        // drain z (only owned samples; rows beyond their length keep whatever is there -- never read)
        auto drain = [&](int td, int slot) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int a = td + col;
                const int rl = rlen(i);
                float4 q = *reinterpret_cast<const float4*>(tile + ((lane >> 3) + 4 * i) * kFiltPitch + slot + col);
                float* d = dst_row(i) + td;
                if (a + 4 <= rl) {
                    *reinterpret_cast<float4*>(d) = q;
                } else {
                    if (a < rl) d[0] = q.x;
                    if (a + 1 < rl) d[1] = q.y;
                    if (a + 2 < rl) d[2] = q.z;
                }
            }
        };
#ifndef FSEM_FILTER_DRAIN32
        if (owned) {                                         // owned tiles start on an even slot (chunks are multiples of 64)
            if (cs == 32) { drain(t - 32, 0); drain(t, 32); }
            else if (t + 32 >= t_stop) drain(t, 0);
        }
#else
        if (owned) drain(t, 0);
#endif
        __syncwarp();
    }
    if (sig_ok) partial[((int64_t)half * batch + my_item) * nchunks + c] = acc_d;
}

// ------------------------------------------------------------------------------------------------
// Variable-length batches only (one CTA): `order` = item indices bucket-sorted by length (256 buckets, ascending;
// the order inside a bucket is arbitrary and does not affect any result), and frame_prefix[i] = number of valid
// STFT frames of items 0..i-1 (frame_prefix[batch] = total).
constexpr int kOrderThreads = 1024;

__device__ __forceinline__ int length_bucket(int len, int64_t n) { return (int)(((int64_t)len * 256) / (n + 1)); }

__global__ void __launch_bounds__(kOrderThreads)
pesq_order_kernel(const int32_t* __restrict__ lengths, int64_t batch, int64_t n, int32_t* __restrict__ order,
                  int64_t* __restrict__ frame_prefix) {
    __shared__ int s_hist[256];
    __shared__ long long s_warp[kOrderThreads / 32];
    __shared__ long long s_carry;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid < 256) s_hist[tid] = 0;
    if (tid == 0) s_carry = 0;
    __syncthreads();
    for (int64_t i = tid; i < batch; i += kOrderThreads) atomicAdd(&s_hist[length_bucket(item_length(lengths, i, n), n)], 1);
    __syncthreads();
    if (tid == 0) {                                               // exclusive scan of the 256 bucket counts
        int run = 0;
        for (int k = 0; k < 256; ++k) { const int c = s_hist[k]; s_hist[k] = run; run += c; }
    }
    __syncthreads();
    for (int64_t i = tid; i < batch; i += kOrderThreads)
        order[atomicAdd(&s_hist[length_bucket(item_length(lengths, i, n), n)], 1)] = (int32_t)i;
    // exclusive prefix sum of the frame counts, 1024 items per round
    for (int64_t base = 0; base < batch; base += kOrderThreads) {
        const int64_t i = base + tid;
        const long long T = i < batch ? pesq_num_frames(item_length(lengths, i, n)) : 0;
        long long v = T;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const long long up = __shfl_up_sync(kFull, v, o);
            if (lane >= o) v += up;
        }
        if (lane == 31) s_warp[warp] = v;
        __syncthreads();
        if (warp == 0) {
            long long w = s_warp[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const long long up = __shfl_up_sync(kFull, w, o);
                if (lane >= o) w += up;
            }
            s_warp[lane] = w;                                     // inclusive totals of warps 0..lane
        }
        __syncthreads();
        const long long before = s_carry + (warp > 0 ? s_warp[warp - 1] : 0);
        if (i < batch) frame_prefix[i] = before + v - T;
        __syncthreads();
        if (tid == kOrderThreads - 1) s_carry = before + v;
        __syncthreads();
    }
    if (tid == 0) frame_prefix[batch] = s_carry;
}

// ------------------------------------------------------------------------------------------------
// Kernel B: persistent warps; one warp = one (item, frame) unit at a time.
#ifndef FSEM_FFT_WARPS
#define FSEM_FFT_WARPS 8
#endif
#ifndef FSEM_FFT_MINBLOCKS
#define FSEM_FFT_MINBLOCKS 2
#endif
constexpr int kSpecWarps = FSEM_FFT_WARPS;

// Shared-memory plan of the spectrum kernel (dynamic): per warp an FFT exchange buffer, a band staging row and a
// 3-slot ring of half-frames (256 samples of the clean and of the degraded signal per slot) filled by TMA bulk
// copies: frame f needs halves f and f+1, half f+2 is in flight while frame f is transformed, so the global-load
// latency is off the critical path and every sample is fetched from L2/HBM once (frames overlap by 50 %).
// Row pitch of the Bark buffer: 49 bands + 1 pad float.  The Bark kernel reads a tile as (frame = lane / 2, band = lane % 2
// + 2k): with 50 floats per frame the 32 lanes of a warp hit 32 different banks (18 * frame mod 32 runs through the even
// banks); with 49 two of them collide on every access (20 % of that kernel's shared-memory wavefronts).
constexpr int kBarkRow = FSEM_PESQ_NBANDS + 1;

// Level alignment (PESQ.py:97-100): band-pass energy of one signal = fixed-order sum of the IIR pass's chunk partials,
// g^2 = 1e7 / (sum(y^2) / (len + 5120) / 1.04684).  One definition for the spectrum kernel (which applies g^2 to the
// Bark powers it stores) and the Bark kernel (which decides NaN / all-zero items on it).
// called by all 32 lanes of a warp; every lane returns the same value (strided partial sums, xor-butterfly: a fixed order).
// Tiny batches cut the signal into hundreds of chunks, so a serial sum per frame-warp would dominate their spectrum kernel.
__device__ __forceinline__ double pesq_band_power(const double* __restrict__ partial, int nchunks, int64_t batch, int sig,
                                                  int64_t item, int lane) {
    const double* p = partial + ((int64_t)sig * batch + item) * nchunks;
    double acc = 0.0;
    for (int c = lane; c < nchunks; c += 32) acc += p[c];
    return warp_sum(acc);
}
__device__ __forceinline__ float pesq_level_gain(double power, int len) {
    return (float)(1.0e7 * ((double)len + 5120.0) * 1.04684 / power);
}

struct SpecWarpSmem {
    float2 fft[kFftBufElems];          // 6016 B: FFT exchanges, then the band stage's P and S rows
    float half_c[3][FSEM_PESQ_HOP];    // 3072 B
    float half_d[3][FSEM_PESQ_HOP];    // 3072 B
    unsigned long long bar[3];         //   24 B (+8 pad)
    float gain[2];                     //    8 B: level-alignment gains g^2 of the current item (clean, degraded)
};
static_assert(kBandBufFloats <= 2 * kFftBufElems, "band rows alias the FFT exchange buffer");
// extra lanes a Bark band may reach back into: bands 0..31 (<= 4 bins at 16 kHz) at most one, bands 32..48 at most four
// (fsem_pesq_create checks the design against these bounds)
constexpr int kBarkPiecesLow = 1;
constexpr int kBarkPiecesHigh = 4;
constexpr size_t kSpecDynSmem = sizeof(SpecWarpSmem) * kSpecWarps;
static_assert(sizeof(SpecWarpSmem) % 16 == 0, "per-warp shared block must keep 16-byte alignment");

__global__ void __launch_bounds__(kSpecWarps * 32, FSEM_FFT_MINBLOCKS)
pesq_spectrum_kernel(const float* __restrict__ z, int64_t zstride, const int32_t* __restrict__ lengths,
                     const int64_t* __restrict__ frame_prefix, int64_t batch, int64_t n, int tmax, int tpitch,
                     const double* __restrict__ partial /* [2][batch][nchunks] band-pass energies of the IIR pass */,
                     int nchunks, const PesqTables* __restrict__ tab,
                     float* __restrict__ bark /* [2][batch][tpitch][kBarkRow], level-aligned */) {
    extern __shared__ __align__(16) unsigned char s_raw[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    SpecWarpSmem& sm = reinterpret_cast<SpecWarpSmem*>(s_raw)[warp];
    float2* buf = sm.fft;
    float* wbuf = reinterpret_cast<float*>(sm.fft);
    const float* S = wbuf + kBandSOffset;
    const uint32_t bar0 = smem_u32(&sm.bar[0]);
    const uint32_t hc0 = smem_u32(&sm.half_c[0][0]), hd0 = smem_u32(&sm.half_d[0][0]);
    constexpr uint32_t kHalfBytes = FSEM_PESQ_HOP * sizeof(float);
    if (lane == 0) {
        for (int i = 0; i < 3; ++i) mbar_init(bar0 + 8 * i, 1);
        mbar_fence_init();
    }
    __syncwarp();

    FftTwiddles tw;
    tw.init(lane);
    // First half of the periodic Hann window only: win[4h + j] = hann[2*lane + h + 64 j], j < 4.  The second half is
    // hann[n + 256] = 1 - hann[n] (0.5 -+ 0.5 cos), applied as y - y * hann[n] with one FMA: eight registers fewer in a
    // kernel that sits at the 128-register occupancy limit.  torch's float32 table deviates from this identity by at
    // most 1.8e-7 (its own cosine rounding), three orders of magnitude below the float32 FFT's round-off.
    float win[8];
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int j = 0; j < 4; ++j) win[4 * h + j] = tab->hann[fft_in_index(lane, h, j)];
    FftLaneBins bins;
    bins.init(lane);
    BandScan scan;
    scan.init(tab->band_first, FSEM_PESQ_NBANDS, lane, bins);
    // this lane gathers Bark bands `lane` and `lane + 32`
    BandGather<kBarkPiecesLow> g_lo;
    BandGather<kBarkPiecesHigh> g_hi;
    g_lo.init(tab->band_first[lane], tab->band_first[lane] + tab->band_count[lane], true);
    {
        const bool has_hi = lane + 32 < FSEM_PESQ_NBANDS;
        const int bh = has_hi ? lane + 32 : 0;
        g_hi.init(tab->band_first[bh], tab->band_first[bh] + tab->band_count[bh], has_hi);
    }
    // power-density correction * Sp of the two bands this lane stores (bark.py:132, 204), times the 1/4 of the packed
    // power spectrum (exact: a power of two)
    const float scale0 = tab->pow_dens[lane] * kPackedPowerScale;
    const float scale1 = (lane + 32 < FSEM_PESQ_NBANDS) ? tab->pow_dens[lane + 32] * kPackedPowerScale : 0.f;

    // Work units = VALID (item, frame) pairs in item-major order; frame_prefix[i] = number of units before item i
    // (nullptr: every item has tmax frames).  Every warp owns a contiguous, equally long range of them, so ragged
    // batches are balanced by frames, not by padded length.
    const int64_t units = frame_prefix ? frame_prefix[batch] : batch * (int64_t)tmax;
    const int64_t nwarps = (int64_t)gridDim.x * kSpecWarps;
    const int64_t per = (units + nwarps - 1) / nwarps;
    const int64_t v0 = ((int64_t)blockIdx.x * kSpecWarps + warp) * per;
    int remaining = (int)(min(units, v0 + per) - v0);            // per-warp share: far below 2^31
    if (remaining <= 0) return;
    int64_t item;
    int f;
    if (frame_prefix) {                                          // last item with frame_prefix[item] <= v0
        int64_t lo = 0, hi = batch;
        while (hi - lo > 1) {
            const int64_t mid = (lo + hi) >> 1;
            if (frame_prefix[mid] <= v0) lo = mid; else hi = mid;
        }
        item = lo;
        f = (int)(v0 - frame_prefix[lo]);
    } else {
        item = v0 / tmax;
        f = (int)(v0 - item * tmax);
    }
    int len = item_length(lengths, item, n);
    int T = pesq_num_frames(len);

    // Ring bookkeeping: half h of an item = samples [256 h, 256 h + 256) of both signals.  Slots rotate 0,1,2;
    // `par` bit s is the mbarrier phase parity the latest copy into slot s completes (flipped at every issue).
    unsigned par = 7u;
    auto issue_half = [&](int slot, int64_t it, int h) {
        par ^= 1u << slot;
        if (lane == 0) {
            fence_proxy_async();
            mbar_arrive_expect_tx(bar0 + 8 * slot, 2 * kHalfBytes);
            bulk_copy_g2s(hc0 + slot * kHalfBytes, z + it * zstride + h * FSEM_PESQ_HOP, kHalfBytes, bar0 + 8 * slot);
            bulk_copy_g2s(hd0 + slot * kHalfBytes, z + (batch + it) * zstride + h * FSEM_PESQ_HOP, kHalfBytes,
                          bar0 + 8 * slot);
        }
    };
    int sa = 0, sb = 1;                                          // slots of the current frame's two halves
    issue_half(0, item, f);
    issue_half(1, item, f + 1);
    // level alignment is linear up to the power spectrum: g^2 of the item multiplies the Bark powers as they are stored
    // (an all-zero item has g^2 = inf and stores NaN rows; the Bark kernel never reads them)
    auto set_gain = [&](int64_t it, int ln) {
        const float gc = pesq_level_gain(pesq_band_power(partial, nchunks, batch, 0, it, lane), ln);
        const float gd = pesq_level_gain(pesq_band_power(partial, nchunks, batch, 1, it, lane), ln);
        if (lane == 0) { sm.gain[0] = gc; sm.gain[1] = gd; }
        __syncwarp();
    };
    set_gain(item, len);

    while (true) {
        const int sn = 3 - sa - sb;                              // the free slot
        // look ahead: next valid unit of this warp's range; fast path = next frame of the same item
        const bool has_next = remaining > 1;
        const bool same_item = has_next && f + 1 < T;
        int64_t item2 = item;
        int len2 = len, T2 = T;
        if (has_next && !same_item) {                            // next item that has frames (one exists: remaining > 1)
            do {
                ++item2;
                len2 = item_length(lengths, item2, n);
                T2 = pesq_num_frames(len2);
            } while (T2 == 0 && item2 + 1 < batch);
        }
        // wait for the two halves of the current frame
        mbar_wait(bar0 + 8 * sa, (par >> sa) & 1u);
        mbar_wait(bar0 + 8 * sb, (par >> sb) & 1u);
        // lane L owns samples 2L + h + 64 j of the frame: pairs (h = 0, 1) are one LDS.64; j < 4 lie in the first half
        const float2* ac = reinterpret_cast<const float2*>(sm.half_c[sa]) + lane;
        const float2* ad = reinterpret_cast<const float2*>(sm.half_d[sa]) + lane;
        const float2* bc = reinterpret_cast<const float2*>(sm.half_c[sb]) + lane;
        const float2* bd = reinterpret_cast<const float2*>(sm.half_d[sb]) + lane;
        float re[16], im[16];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float2 c0 = ac[32 * j], d0 = ad[32 * j], c1 = bc[32 * j], d1 = bd[32 * j];
            re[j] = c0.x * win[j];                    re[8 + j] = c0.y * win[4 + j];
            im[j] = d0.x * win[j];                    im[8 + j] = d0.y * win[4 + j];
            re[4 + j] = fmaf(-win[j], c1.x, c1.x);    re[12 + j] = fmaf(-win[4 + j], c1.y, c1.y);
            im[4 + j] = fmaf(-win[j], d1.x, d1.x);    im[12 + j] = fmaf(-win[4 + j], d1.y, d1.y);
        }
        if (f * FSEM_PESQ_HOP + FSEM_PESQ_NFFT > len) {          // warp-uniform, last frame(s) only:
            const int room = len - f * FSEM_PESQ_HOP;            // zero padding beyond the signal (PESQ.py:128-130)
#pragma unroll
            for (int h = 0; h < 2; ++h)
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    if (fft_in_index(lane, h, j) >= room) { re[8 * h + j] = 0.f; im[8 * h + j] = 0.f; }
        }
        // Refill the free slot now: the frame's samples are in registers, the copy has the whole transform to land, and
        // the proxy fence inside issue_half only has this frame's shared-memory loads to wait for (at the top of the loop
        // it also waited for the previous frame's global stores).  The free slot was last read one frame ago.
        __syncwarp();
        if (has_next) {
            if (same_item) issue_half(sn, item, f + 2);
            else issue_half(sn, item2, 0);
        }
        float ar[8], ai[8], br[8], bi[8];
        warp_fft512<false>(re, im, buf, tw, lane, ar, ai, br, bi);
        float pc[8], pd[8];
        packed_power_regs(ar, ai, br, bi, lane, pc, pd);
        if (lane == 0) { pc[0] = 0.f; pd[0] = 0.f; }             // bin 0: "we won't use energy feature" (PESQ.py:136)
        scan.scan_store(pc, pd, wbuf, lane);
        float* __restrict__ out_c = bark + (item * tpitch + f) * kBarkRow;
        float* __restrict__ out_d = bark + ((batch + item) * tpitch + f) * kBarkRow;
        const float lo_c = g_lo.sum(S), lo_d = g_lo.sum(S + kBandSStride);
        const float hi_c = g_hi.sum(S), hi_d = g_hi.sum(S + kBandSStride);     // all lanes: no divergence around the loads
        const float2 g2 = *reinterpret_cast<const float2*>(sm.gain);
        out_c[lane] = (lo_c * scale0) * g2.x;
        out_d[lane] = (lo_d * scale0) * g2.y;
        if (lane + 32 < FSEM_PESQ_NBANDS) {
            out_c[lane + 32] = (hi_c * scale1) * g2.x;
            out_d[lane + 32] = (hi_d * scale1) * g2.y;
        }
        if (!has_next) break;
        --remaining;
        if (same_item) {                                         // frame f+1 = halves (f+1, f+2) = slots (sb, sn)
            sa = sb; sb = sn;
            ++f;
        } else {                                                 // new item: its first half sits in sn, fetch the second
            item = item2; f = 0; len = len2; T = T2;
            const int freed = sa;                                // both old slots are free; take the older one
            sa = sn; sb = freed;
            __syncwarp();
            issue_half(sb, item, f + 1);
            set_gain(item, len);                                 // every lane has read the old pair (syncwarp above)
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Kernel C: one CTA per item; Bark-domain model.
// Two shapes of the same code: (128 threads, 64-frame tiles) for throughput -- 8 CTAs per SM hide each other's
// barriers (128-frame tiles were slower) -- and (640 threads, 320-frame tiles) for batches that cannot fill the SMs
// anyway, where the kernel's latency is the number of serial tile rounds (2 instead of 10 for a 10 s utterance).
// Frame tiles live in dynamic shared memory: 2 x kTile x 49 Bark powers + the per-frame flags.
#ifndef FSEM_BARK_THREADS
#define FSEM_BARK_THREADS 128
#endif
constexpr int kBarkThreads = FSEM_BARK_THREADS;
constexpr int kBarkTile = FSEM_BARK_THREADS / 2;
constexpr int kBarkThreadsWide = 640;
constexpr int kBarkTileWide = 320;
__host__ __device__ constexpr size_t bark_dyn_smem(int tile) { return tile <= 64 ? 0 : sizeof(float) * (size_t)tile * (2 * kBarkRow + 2); }

// x^y for x > 0 through the SFU (lg2.approx / ex2.approx, one MUFU each: the .ftz forms skip the denormal range fix-ups,
// and every argument here is a normal number): relative error ~1e-6 for the exponents used (|y * log2 x| < 10), three
// orders of magnitude inside the PESQ budget and ~20x cheaper than powf.  NaN propagates.
__device__ __forceinline__ float mufu_lg2(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float mufu_ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float mufu_rcp(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float fast_pow(float x, float y) { return mufu_ex2(y * mufu_lg2(x)); }

__device__ __forceinline__ float zwicker_loudness(float p, float thr, float inv_thr, float e, float scale) {
    // loudness.py:62-67: Sl*(2 thr)^e * ((0.5 + 0.5 p/thr)^e - 1), 0 where p <= thr
    float l = scale * (fast_pow(fmaf(0.5f * p, inv_thr, 0.5f), e) - 1.f);
    return (p <= thr) ? 0.f : l;   // NaN p: comparison false -> l (NaN) propagates like the reference
}

// The Bark rows arrive LEVEL-ALIGNED (the spectrum kernel applies g^2) in rows of `tpitch` = round4(tmax) frames per
// (signal, item), so every frame tile is one contiguous, 16-byte aligned run: tiles are fetched with two bulk copies
// (cp.async.bulk + mbarrier, no register pass), 1.40 -> see DESIGN.md.
template <int kThreads, int kTile>
__global__ void __launch_bounds__(kThreads)
pesq_bark_kernel(const float* __restrict__ bark, const double* __restrict__ partial, int nchunks,
                 const int32_t* __restrict__ lengths, const int32_t* __restrict__ order, int64_t batch, int64_t n, int tpitch,
                 const PesqTables* __restrict__ tab, float* __restrict__ dist_ws /* [2][batch][tpitch] */,
                 float* __restrict__ mos_out, int32_t* __restrict__ status_out,
                 double* __restrict__ power_out /* [2][batch] */) {
    static_assert(kThreads == 2 * kTile && kThreads % 32 == 0 && kThreads >= 2 * FSEM_PESQ_NBANDS, "two threads per frame of a tile");
    static_assert(kTile % 4 == 0, "tile rows are fetched in 16-byte units");
    // the throughput shape keeps its tiles in static arrays (the compiler schedules the unrolled band loops better
    // around provably distinct arrays); only the wide shape needs dynamic memory
    constexpr bool kDyn = bark_dyn_smem(kTile) > 0;
    extern __shared__ __align__(16) float s_bark_dyn[];
    __shared__ __align__(16) float s_tile_st[kDyn ? 1 : 2][kDyn ? 1 : kTile][kBarkRow];
    __shared__ float s_live_st[kDyn ? 1 : kTile];
    __shared__ float s_fr_st[kDyn ? 1 : kTile];
    float (*s_tile)[kTile][kBarkRow] =
        kDyn ? reinterpret_cast<float (*)[kTile][kBarkRow]>(s_bark_dyn)
             : reinterpret_cast<float (*)[kTile][kBarkRow]>(&s_tile_st[0][0][0]);
    float* s_live = kDyn ? s_bark_dyn + 2 * kTile * kBarkRow : s_live_st;     // 1 - silent flag of a frame
    float* s_fr = kDyn ? s_bark_dyn + 2 * kTile * kBarkRow + kTile : s_fr_st;
    __shared__ float s_thr[FSEM_PESQ_NBANDS], s_thr100[FSEM_PESQ_NBANDS], s_ithr[FSEM_PESQ_NBANDS], s_exp[FSEM_PESQ_NBANDS],
        s_lsc[FSEM_PESQ_NBANDS], s_w[FSEM_PESQ_NBANDS], s_ratio[FSEM_PESQ_NBANDS];
    __shared__ float s_g2[2];
    __shared__ float s_carry[2];                          // last frame ratio of the previous tile, by tile parity
    __shared__ float s_red[2][kThreads / 32];
    __shared__ __align__(8) unsigned long long s_bar;

    const int tid = threadIdx.x;
    // ragged batches: CTAs are scheduled in launch order, so the longest items go first (no long item left for the tail)
    const int64_t item = order ? (int64_t)order[batch - 1 - blockIdx.x] : (int64_t)blockIdx.x;
    const int len = item_length(lengths, item, n);
    const int T = pesq_num_frames(len);
    const uint32_t bar = smem_u32(&s_bar);

    if (tid < FSEM_PESQ_NBANDS) {
        float thr = tab->thresh[tid];
        s_thr[tid] = thr;
        s_thr100[tid] = thr * 100.f;
        s_ithr[tid] = 1.f / thr;
        s_exp[tid] = tab->zw_exp[tid];
        s_lsc[tid] = tab->loud_scale[tid];
        s_w[tid] = tab->width[tid];
    }
    if (tid < 64) {
        // level alignment (PESQ.py:97-100): g^2 = 1e7 / (sum(y^2) / (len + 5120) / 1.04684); applied by the spectrum
        // kernel (pesq_level_gain, the same expression), recomputed here for the NaN / all-zero decision (warp = signal)
        const int sig = tid >> 5;
        const double acc = pesq_band_power(partial, nchunks, batch, sig, item, tid & 31);
        if ((tid & 31) == 0) {
            power_out[(int64_t)sig * batch + item] = acc;
            s_g2[sig] = pesq_level_gain(acc, len);
        }
    }
    if (tid == 0) {
        mbar_init(bar, 1);
        mbar_fence_init();
    }
    __syncthreads();
    if (T < 20 || !(isfinite(s_g2[0]) && isfinite(s_g2[1]))) {
        if (tid == 0) {
            mos_out[item] = nanf("");
            if (status_out) status_out[item] = (T < 20) ? FSEM_ITEM_TOO_SHORT : FSEM_ITEM_NAN;
        }
        return;
    }
    const float* __restrict__ bc = bark + item * (int64_t)tpitch * kBarkRow;
    const float* __restrict__ bd = bark + (batch + item) * (int64_t)tpitch * kBarkRow;
    const uint32_t tile_c = smem_u32(&s_tile[0][0][0]), tile_d = smem_u32(&s_tile[1][0][0]);
    unsigned parity = 0;

    // all threads have finished with the previous tile (barrier), then one thread starts the two copies; rows beyond
    // nf (up to three, inside the item's own pitch) ride along and are never read
    auto load_tile = [&](int f0, int nf) {
        __syncthreads();
        if (tid == 0) {
            const uint32_t bytes = (uint32_t)((nf + 3) & ~3) * (kBarkRow * sizeof(float));
            fence_proxy_async();
            mbar_arrive_expect_tx(bar, 2 * bytes);
            bulk_copy_g2s(tile_c, bc + (int64_t)f0 * kBarkRow, bytes, bar);
            bulk_copy_g2s(tile_d, bd + (int64_t)f0 * kBarkRow, bytes, bar);
        }
        mbar_wait(bar, parity);
        parity ^= 1u;
    };

    // ---- phase 1: silent-frame flags and mean audible band power (PESQ.py:144-147)
    double band_acc = 0.0;   // thread t < 98: signal t / 49, band t % 49
    const int my_sig = tid / FSEM_PESQ_NBANDS;
    const int my_band = tid - my_sig * FSEM_PESQ_NBANDS;
    for (int f0 = 0; f0 < T; f0 += kTile) {
        const int nf = min(kTile, T - f0);
        load_tile(f0, nf);
        {   // two threads per frame (even / odd bands), combined with one shuffle
            const int fs = tid >> 1, half = tid & 1;
            float a = 0.f;
            if (fs < nf) {
#pragma unroll 5
                for (int b = half; b < FSEM_PESQ_NBANDS; b += 2) {
                    float p = s_tile[0][fs][b];
                    a += p * ((p > s_thr100[b]) ? 1.f : 0.f);
                }
            }
            a += __shfl_xor_sync(kFull, a, 1);
            if (fs < nf && half == 0) s_live[fs] = (a < 1.0e7f) ? 0.f : 1.f;
        }
        __syncthreads();
        if (tid < 2 * FSEM_PESQ_NBANDS) {
            const float thr100 = s_thr100[my_band];
            const float* col = &s_tile[my_sig][0][my_band];
            float acc = 0.f;                                   // one tile in float32, tiles in float64
#pragma unroll 4
            for (int f = 0; f < nf; ++f) {
                const float p = col[f * kBarkRow];
                acc += (p > thr100) ? p * s_live[f] : 0.f;
            }
            band_acc += (double)acc;
        }
    }
    __syncthreads();                                         // the last tile is dead: its memory carries the 98 band means
    double* s_mean = reinterpret_cast<double*>(&s_tile[0][0][0]);
    if (tid < 2 * FSEM_PESQ_NBANDS) s_mean[tid] = band_acc / (double)T;
    __syncthreads();
    if (tid < FSEM_PESQ_NBANDS) {
        float r = (float)((s_mean[FSEM_PESQ_NBANDS + tid] + 1000.0) / (s_mean[tid] + 1000.0));
        s_ratio[tid] = fminf(fmaxf(r, 0.01f), 100.f);
    }
    if (tid == 0) s_carry[0] = 0.f;

    const float wtot = tab->width_total;
    float* __restrict__ dsym = dist_ws + item * (int64_t)tpitch;
    float* __restrict__ dasym = dist_ws + (batch + item) * (int64_t)tpitch;

    // ---- phase 2: frame equalisation, loudness, disturbances (PESQ.py:149-224)
    for (int f0 = 0; f0 < T; f0 += kTile) {
        const int nf = min(kTile, T - f0);
        load_tile(f0, nf);                                  // its leading barrier also publishes s_ratio / s_carry
        // two threads per frame: thread 2*fs handles the even bands, 2*fs+1 the odd ones
        const int fs = tid >> 1, half = tid & 1;
        const bool live = fs < nf;
        float afp_c = 0.f, afp_d = 0.f;
        if (live) {
#pragma unroll 5
            for (int b = half; b < FSEM_PESQ_NBANDS; b += 2) {
                float c = s_ratio[b] * s_tile[0][fs][b];
                float d = s_tile[1][fs][b];
                afp_c += c * ((c > s_thr[b]) ? 1.f : 0.f);
                afp_d += d * ((d > s_thr[b]) ? 1.f : 0.f);
            }
        }
        afp_c += __shfl_xor_sync(kFull, afp_c, 1);
        afp_d += __shfl_xor_sync(kFull, afp_d, 1);
        if (live && half == 0) s_fr[fs] = (afp_c + 5.0e3f) / (afp_d + 5.0e3f);
        __syncthreads();
        float sym_acc = 0.f, asym_acc = 0.f;
        if (live) {
            const int f = f0 + fs;
            float fr = s_fr[fs];
            if (f > 0) {
                float prev = (fs == 0) ? s_carry[(f0 / kTile) & 1] : s_fr[fs - 1];
                fr = 0.8f * fr + 0.2f * prev;               // non-recursive smoothing (PESQ.py:159)
            }
            fr = fminf(fmaxf(fr, 3.0e-4f), 5.f);
            for (int b = half; b < FSEM_PESQ_NBANDS; b += 2) {
                const float c = s_ratio[b] * s_tile[0][fs][b];
                const float d = fr * s_tile[1][fs][b];
                const float thr = s_thr[b], ithr = s_ithr[b], e = s_exp[b], lsc = s_lsc[b];
                const float lc = zwicker_loudness(c, thr, ithr, e, lsc);
                const float ld = zwicker_loudness(d, thr, ithr, e, lsc);
                const float dead = 0.25f * fminf(lc, ld);
                float diff = ld - lc;
                float mag = fmaxf(fabsf(diff) - dead, 0.f);
                float dist = copysignf(mag, diff);
                if (diff != diff) dist = diff;             // keep NaN
                // band 0 excluded (bark.py:184): its width weight is dropped here
                const float wd = (b >= 1) ? s_w[b] * dist : 0.f;
                sym_acc = fmaf(wd, wd, sym_acc);
                const float ratio = (d + 50.f) * mufu_rcp(c + 50.f);   // c + 50 >= 50: a normal number
                // ratio^1.2 < 3  <=>  ratio < 3^(1/1.2): decide on the ratio itself (no pow error in the decision);
                // branch-free: the pow is two MUFU operations
                const float scale = (ratio < 2.49804953f) ? 0.f : fminf(fast_pow(ratio, 1.2f), 12.f);
                asym_acc += fabsf(wd * scale);
            }
        }
        sym_acc += __shfl_xor_sync(kFull, sym_acc, 1);
        asym_acc += __shfl_xor_sync(kFull, asym_acc, 1);
        if (live && half == 0) {
            const int f = f0 + fs;
            float sym = fmaxf(sqrtf(wtot * sym_acc), 1.0e-20f);
            float asym = fmaxf(asym_acc, 1.0e-20f);
            const float weight = fast_pow((afp_c + 1.0e5f) * 1.0e-7f, 0.04f);
            dsym[f] = fminf(sym / weight, 45.f);
            dasym[f] = fminf(asym / weight, 45.f);
        }
        if (tid == 2 * (nf - 1)) s_carry[(f0 / kTile + 1) & 1] = s_fr[nf - 1];   // for the next tile (other slot: thread 0
                                                                                 // may still be reading this tile's)
    }
    __syncthreads();   // dsym/dasym written by this CTA are visible to it after the barrier

    // ---- phase 3: L6 over 20-frame windows (hop 10), L2 across windows, MOS (PESQ.py:168-172, 240-243)
    const int W = (T - 20) / 10 + 1;
    float acc_s = 0.f, acc_a = 0.f;
    for (int w = tid; w < W; w += kThreads) {
        float s6 = 0.f, a6 = 0.f;
#pragma unroll 4
        for (int j = 0; j < 20; ++j) {
            float s = dsym[10 * w + j], a = dasym[10 * w + j];
            float s2 = s * s, a2 = a * a;
            s6 = fmaf(s2 * s2, s2, s6);
            a6 = fmaf(a2 * a2, a2, a6);
        }
        float ps = powf(s6 * 0.05f, 1.f / 6.f), pa = powf(a6 * 0.05f, 1.f / 6.f);
        acc_s = fmaf(ps, ps, acc_s);
        acc_a = fmaf(pa, pa, acc_a);
    }
    acc_s = warp_sum(acc_s);
    acc_a = warp_sum(acc_a);
    if ((tid & 31) == 0) { s_red[0][tid >> 5] = acc_s; s_red[1][tid >> 5] = acc_a; }
    __syncthreads();
    if (tid == 0) {
        float ts = 0.f, ta = 0.f;
        for (int i = 0; i < kThreads / 32; ++i) { ts += s_red[0][i]; ta += s_red[1][i]; }
        float d_sym = sqrtf(ts / (float)W), d_asym = sqrtf(ta / (float)W);
        float mos = 4.5f - 0.1f * d_sym - 0.0309f * d_asym;
        mos = 0.999f + 4.f / (1.f + expf(-1.3669f * mos + 3.8224f));
        mos_out[item] = mos;
        if (status_out) status_out[item] = (mos == mos) ? FSEM_ITEM_OK : FSEM_ITEM_NAN;
    }
}

}  // namespace fsem
