// Single-read first pass of the two-metric entry point: every input sample is fetched from
// HBM ONCE and feeds both PESQ's IIR pass (pesq_filter_tiled_kernel: band-pass power partials + tapered,
// pre-emphasised z, PESQ.py:92-113) and STOI's 16 k -> 10 k polyphase resampler with its fused hop energies
// (stoi_resample85_kernel: y and A_h / B_h, base.py:19-20 + STOI.py:92-98).  The two separate kernels are both
// HBM-bound (74 % and 70 % of the measured copy bandwidth); together they move 38.5 GB per 8192 x 10 s step, this
// kernel moves 27.7 GB.
//
// Work decomposition is the IIR pass's: one warp = 32 signals x one time chunk, lane = signal, serial in time,
// `warm` samples of IIR warm-up before the chunk.  Chunks are multiples of 1024 input samples = 640 outputs = 5
// hops, so every output sample and every hop energy has exactly one owner.  Per iteration a 32-sample raw tile of
// the 32 signals lands in a warp-private three-slot ring (rows = signals, pitch 100 floats: conflict-free
// LDS.128 for lane = row) through 16-byte cp.async copies issued one tile ahead (no staging registers); then
//   * the resampler produces the four polyphase blocks whose 28-tap windows end inside the new tile
//     (inputs [t-16, t+16), i.e. it needs x[t-26 .. t+25]: the previous and the current slot),
//   * the IIRs run over the PREVIOUS tile in place (its raw samples are no longer needed) and its z is drained,
// so z needs no staging of its own: the three slots hold the previous, the current and the in-flight tile.  z, the band-power partials and y are bit-identical
// to the two separate kernels' (same recurrences, same FMA order); the hop energies are the same fp64 sums of the
// same fp32-rounded products accumulated serially per lane instead of by a warp tree (differences ~1e-16 relative,
// far below the float rounding of the frame norm they feed).
#pragma once
#include "fsem_pesq.cuh"
#include "fsem_stoi.cuh"

namespace fsem {

constexpr int kFpWarps = 2;           // warps (= independent work units) per CTA: 31 KB of shared memory, 7 CTAs per SM
constexpr int kFpRingPitch = 100;    // floats per ring row: 3 slots x 32 samples + 4 pad (100 mod 32 = 4: conflict-free LDS.128)
constexpr int kFpOutPitch = 22;      // floats per staged output row: 20 outputs + 2 pad (conflict-free STS.64)
constexpr int kFpQuantum = 1024;     // chunk granularity in input samples (5 hops of 128 outputs)
constexpr int kFpWarpFloats = 32 * kFpRingPitch + 32 * kFpOutPitch;
constexpr size_t kFpDynSmem = sizeof(float) * kFpWarps * kFpWarpFloats;
constexpr size_t kFpIirDynSmem = sizeof(float) * kFpWarps * 32 * kFpRingPitch;   // kResample = false: no output staging

// 16-byte asynchronous global -> shared copy (LDGSTS); bytes beyond `src_bytes` are zero-filled
__device__ __forceinline__ void cp_async16(float* dst, const float* src, int src_bytes) {
    asm volatile("cp.async.cg.shared.global.L2::256B [%0], [%1], 16, %2;" ::"r"(smem_u32(dst)), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int kPending>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(kPending) : "memory"); }

// analysis window of the silent-frame energies as a kernel argument: indexed with warp-uniform offsets (LDC)
struct StoiWindowArg { float w[FSEM_STOI_WIN]; };

#ifndef FSEM_FP_MINBLOCKS
#define FSEM_FP_MINBLOCKS 7
#endif
// kResample = false: the same tile ring and IIR code without the STOI half (no FIR, no y, no hop energies) -- the
// IIR pass on its own with the cp.async prefetch (the raw tile lands two iterations before the IIRs reach it).
template <bool kHasLengths, bool kResample = true>
__global__ void __launch_bounds__(kFpWarps * 32, FSEM_FP_MINBLOCKS)
pesq_stoi_first_pass_kernel(const float* __restrict__ clean, const float* __restrict__ deg,
                            const int32_t* __restrict__ lengths, const int32_t* __restrict__ order, int64_t batch,
                            int64_t n, int64_t stride, int chunk, int nchunks, int warm,
                            const __grid_constant__ PesqFilterCoef P, const __grid_constant__ Resample85Taps taps,
                            const __grid_constant__ StoiWindowArg win, float* __restrict__ z_out, int64_t zstride,
                            double* __restrict__ partial, float* __restrict__ y_out, int64_t ystride,
                            double2* __restrict__ hop_energy, int hops_max) {
    extern __shared__ __align__(16) float s_dyn[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    float* ring = s_dyn + warp * (kResample ? kFpWarpFloats : 32 * kFpRingPitch);
    float* stage = ring + 32 * kFpRingPitch;

    const int64_t groups = ceil_div(batch, 32);
    const int64_t unit = (int64_t)blockIdx.x * kFpWarps + warp;
    if (unit >= 2 * groups * nchunks) return;
    const int half = (int)(unit / (groups * nchunks));
    const int64_t rem = unit - (int64_t)half * groups * nchunks;
    const int64_t grp = rem / nchunks;
    const int c = (int)(rem - grp * nchunks);
    const int64_t row0 = grp * 32;

    // own signal (compute role)
    const bool sig_ok = row0 + lane < batch;
    const int64_t my_item = !sig_ok ? 0 : (kHasLengths && order != nullptr) ? (int64_t)order[row0 + lane] : row0 + lane;
    const int len = sig_ok ? item_length(lengths, my_item, n) : 0;
    const int L = (int)stoi_resampled_len(len, 8, 5);
    const int L0 = (int)stoi_resampled_len(n, 8, 5);         // every row's L when there are no per-item lengths
    const int L_min = L0;                                    // used on the fast path without per-item lengths only
    // rows this lane helps to move (transfer role): slots (lane >> 3) + 4*k, float4 column lane & 7
    const int col = (lane & 7) * 4;
    const int64_t trow = row0 + (lane >> 3);
    const float* __restrict__ sbase = (half ? deg : clean) + col;
    float* __restrict__ dbase = z_out + (int64_t)half * batch * zstride + col;
    const float* __restrict__ src0 = sbase + trow * stride;
    float* __restrict__ dst0 = dbase + trow * zstride;
    const int64_t sstep = 4 * stride, dstep = 4 * zstride;
    int row_len[kHasLengths ? 8 : 1];
    int row_item[kHasLengths ? 8 : 1];
    if (kHasLengths) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int64_t slot = trow + 4 * k;
            const int64_t it = slot >= batch ? 0 : (order != nullptr ? (int64_t)order[slot] : slot);
            row_item[k] = (int)it;
            row_len[k] = slot < batch ? item_length(lengths, it, n) : 0;
        }
    }
    const int rows_ok = (int)min((int64_t)8, (batch - trow + 3) / 4);
    auto rlen = [&](int k) -> int { return kHasLengths ? row_len[kHasLengths ? k : 0] : (k < rows_ok ? (int)n : 0); };
    auto src_row = [&](int k) -> const float* {
        return kHasLengths ? sbase + (int64_t)row_item[kHasLengths ? k : 0] * stride : src0 + k * sstep;
    };
    auto dst_row = [&](int k) -> float* {
        return kHasLengths ? dbase + (int64_t)row_item[kHasLengths ? k : 0] * zstride : dst0 + k * dstep;
    };

    const int t_acc = c * chunk;                             // multiple of kFpQuantum
    const int t_own_end = t_acc + chunk;                     // end of the owned range (not clipped)
    const int grp_len = kHasLengths ? __reduce_max_sync(kFull, len) : (int)n;
    const int t_stop = min(grp_len, t_own_end);              // the IIRs run up to here
    if (t_acc >= grp_len) {                                  // nothing of this chunk exists in any of the 32 signals
        if (sig_ok) partial[((int64_t)half * batch + my_item) * nchunks + c] = 0.0;
        return;
    }
    const int i_first = max(0, t_acc - warm) >> 5;           // first tile loaded (and filtered)
    const int i_end = (t_stop + 31) >> 5;                    // one past the last filtered tile: loaded for the FIR tail

    IirState st;
#pragma unroll
    for (int s = 0; s < FSEM_BP_SECTIONS; ++s) { st.w1[s] = 0.f; st.w2[s] = 0.f; }
    st.s1 = 0.f; st.s2 = 0.f;
    double acc_d = 0.0;
    double hop_a = 0.0, hop_b = 0.0;                         // running A_h / B_h of this lane's signal (clean half only)

    // asynchronous fill of one ring slot with the tile starting at sample tt: 8 x 16 bytes per lane, coalesced
    // 128-byte row segments; samples beyond a row's length arrive as zeros (src-size < 16)
    float* tbase = ring + (lane >> 3) * kFpRingPitch + col;
    const int grp_min = kHasLengths ? __reduce_min_sync(kFull, sig_ok ? len : 0x7fffffff) : (int)n;
    const bool rows_full = row0 + 32 <= batch;               // all 32 slots of the group hold a signal
    auto issue_tile = [&](int tt, int slot_off) {
        if (rows_full && tt + 32 <= grp_min) {               // interior tile of a full group: every row has all 32 samples
#pragma unroll
            for (int k = 0; k < 8; ++k) cp_async16(tbase + 4 * k * kFpRingPitch + slot_off, src_row(k) + tt, 16);
        } else {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int bytes = max(0, min(16, 4 * (rlen(k) - (tt + col))));
                const float* q = src_row(k);
                cp_async16(tbase + 4 * k * kFpRingPitch + slot_off, bytes > 0 ? q + tt : q, bytes);
            }
        }
        cp_async_commit();
    };
    int sp = 0, sc = 32, sn = 64;                            // ring offsets of the previous / current / next tile
    // the slot "before" the first tile: zeros (x[-10..-1] of the resampler's padding when i_first == 0; never used
    // by an owned block otherwise, because warm >= 32)
#pragma unroll
    for (int k = 0; k < 8; ++k)
        *reinterpret_cast<float4*>(tbase + 4 * k * kFpRingPitch + sp) = make_float4(0.f, 0.f, 0.f, 0.f);
    issue_tile(i_first * 32, sc);
    float* row = ring + lane * kFpRingPitch;
    float* srow = stage + lane * kFpOutPitch;
    // drain role for the staged outputs: lanes 0-9 / 10-19 / 20-29 take the ten float2 of rows 3k / 3k+1 / 3k+2
    const int dr_row = lane / 10, dr_c2 = lane - 10 * dr_row;
    const int dr_b0 = (2 * dr_c2) / 5, dr_b1 = (2 * dr_c2 + 1) / 5;     // polyphase blocks of the two outputs of the pair
    // The loop body is kept compact on purpose (rolled loops over sample groups / block pairs / drain rows): with
    // every stage unrolled the hot loop exceeds the instruction cache and the warps stall on instruction fetch.
    for (int i = i_first; i <= i_end; ++i) {
        const int t = i * 32;
        // ---- (a) tile i + 1 goes into the slot drained in the previous iteration; wait for tile i
        if (kResample) {                                     // the FIR needs tile i now
            if (i < i_end) { issue_tile(t + 32, sn); cp_async_wait<1>(); }
            else cp_async_wait<0>();
        } else {                                             // IIR only: tile i - 1 is needed, tiles i and i + 1 may be in flight
            if (i + 1 < i_end) { issue_tile(t + 32, sn); cp_async_wait<2>(); }
            else if (i < i_end) cp_async_wait<1>();
            else cp_async_wait<0>();
        }
        __syncwarp();
        // ---- (b) resampler: blocks kb = 4i-2 .. 4i+1 (inputs [t-16, t+16)), outputs o = 20i-10 .. 20i+9
        const bool rs_any = kResample && (t + 16 > t_acc) && (t - 16 < t_own_end);
        if (rs_any) {
#pragma unroll 1
            for (int bp = 0; bp < 2; ++bp) {
                // two blocks per round: window xw[k] = x[t - 28 + 16 bp + k], k < 40.  x[t-32 .. t-1] sit in the previous
                // slot, x[t .. t+31] in the current one; every aligned group of four lies inside one slot.
                float xw[40];
#pragma unroll
                for (int g = 0; g < 10; ++g) {
                    const int so = -28 + 16 * bp + 4 * g;                     // sample offset of the group relative to t
                    const float4 v = *reinterpret_cast<const float4*>(row + (so < 0 ? sp + 32 + so : sc + so));
                    xw[4 * g] = v.x; xw[4 * g + 1] = v.y; xw[4 * g + 2] = v.z; xw[4 * g + 3] = v.w;
                }
                float o[10];
#pragma unroll
                for (int b = 0; b < 2; ++b) {
#pragma unroll
                    for (int p = 0; p < 5; ++p) {
                        float acc = 0.f;
#pragma unroll
                        for (int j = rs85_lo(p); j <= rs85_hi(p); ++j) acc = fmaf(taps.h[p][j], xw[8 * b + 2 + j], acc);
                        o[5 * b + p] = acc;
                    }
                }
                float* sdst = srow + 10 * bp;                                  // conflict-free STS.64 (pitch 22)
#pragma unroll
                for (int q = 0; q < 5; ++q) *reinterpret_cast<float2*>(sdst + 2 * q) = make_float2(o[2 * q], o[2 * q + 1]);
            }
            if (half == 0) {
                // hop energies of the CLEAN signal from the staged outputs of this lane's own row (same products as
                // stoi_resample85_kernel: rounded to fp32, squares summed in fp64)
#pragma unroll 1
                for (int b = 0; b < 4; ++b) {
                    const int kb = 4 * i - 2 + b;
                    if ((8 * kb < t_acc) || (8 * kb >= t_own_end)) continue;
                    const int o0 = 5 * kb;
                    const int r0 = o0 & (FSEM_STOI_HOP - 1);
                    const float* ov = srow + 5 * b;
                    if (r0 + 5 < FSEM_STOI_HOP) {                          // 24 of 25.6 blocks: no hop boundary in or after the block
#pragma unroll
                        for (int p = 0; p < 5; ++p) {
                            const float v = ov[p];
                            const float fa = __fmul_rn(v, win.w[r0 + p]);
                            const float fb = __fmul_rn(v, win.w[r0 + p + FSEM_STOI_HOP]);
                            hop_a = fma((double)fa, (double)fa, hop_a);
                            hop_b = fma((double)fb, (double)fb, hop_b);
                        }
                    } else {
#pragma unroll 1
                        for (int p = 0; p < 5; ++p) {
                            const int r = (r0 + p) & (FSEM_STOI_HOP - 1);
                            const float v = ov[p];
                            const float fa = __fmul_rn(v, win.w[r]);
                            const float fb = __fmul_rn(v, win.w[r + FSEM_STOI_HOP]);
                            hop_a = fma((double)fa, (double)fa, hop_a);
                            hop_b = fma((double)fb, (double)fb, hop_b);
                            if (r == FSEM_STOI_HOP - 1) {                  // hop complete (warp-uniform): only whole hops inside L count
                                if (sig_ok && o0 + p < L)
                                    hop_energy[my_item * hops_max + ((o0 + p) >> 7)] = make_double2(hop_a, hop_b);
                                hop_a = 0.0; hop_b = 0.0;
                            }
                        }
                    }
                }
            }
            __syncwarp();
            // drain the staged outputs: 32 rows x 20 floats, 80 contiguous bytes per row, three rows per instruction
            {
                const int o = 20 * i - 10 + 2 * dr_c2;
                const float* sv = stage + dr_row * kFpOutPitch + 2 * dr_c2;
                const bool all_own = (32 * i - 16 >= t_acc) && (32 * i + 16 <= t_own_end) && (20 * i + 10 <= L_min);
                if (!kHasLengths && rows_full && all_own) {
                    // interior iteration of a full group: every output is owned and inside its row
                    float* d = y_out + ((int64_t)half * batch + row0 + dr_row) * ystride + o;
                    if (lane < 30) {
#pragma unroll
                        for (int k = 0; k < 10; ++k)
                            *reinterpret_cast<float2*>(d + 3 * k * ystride) =
                                *reinterpret_cast<const float2*>(sv + 3 * k * kFpOutPitch);
                        if (dr_row < 2)
                            *reinterpret_cast<float2*>(d + 30 * ystride) = *reinterpret_cast<const float2*>(sv + 30 * kFpOutPitch);
                    }
                } else {
                    const int kb0 = 4 * i - 2 + dr_b0, kb1 = 4 * i - 2 + dr_b1;
                    const bool blk0 = (8 * kb0 >= t_acc) && (8 * kb0 < t_own_end);
                    const bool blk1 = (8 * kb1 >= t_acc) && (8 * kb1 < t_own_end);
#pragma unroll 1
                    for (int k = 0; k < 11; ++k) {
                        const int r = min(3 * k + dr_row, 31);
                        const bool act = (lane < 30) && (3 * k + dr_row < 32);
                        const float2 v = *reinterpret_cast<const float2*>(stage + r * kFpOutPitch + 2 * dr_c2);
                        int rL;
                        long long ritem;
                        if (kHasLengths) {
                            rL = __shfl_sync(kFull, L, r);
                            ritem = __shfl_sync(kFull, (long long)my_item, r);
                        } else {
                            rL = (row0 + r < batch) ? L0 : 0;
                            ritem = row0 + r;
                        }
                        const bool own0 = act && blk0 && o < rL;
                        const bool own1 = act && blk1 && o + 1 < rL;
                        float* d = y_out + ((int64_t)half * batch + ritem) * ystride + o;
                        if (own0 && own1) *reinterpret_cast<float2*>(d) = v;
                        else if (own0) d[0] = v.x;
                        else if (own1) d[1] = v.y;
                    }
                }
            }
            __syncwarp();
        }
        // ---- (c) IIRs over the previous tile (index i - 1) in place, then drain its z
        const int tp = t - 32;
        if (i > i_first) {
            float* prow = row + sp;
            const bool owned = tp >= t_acc;
            const int t_end = min(len, t_stop);
            float acc = 0.f;
            if (tp >= 16 && tp + 48 <= len) {
                float acc2 = 0.f;
                float4 q = *reinterpret_cast<const float4*>(prow);
#pragma unroll 1
                for (int g = 0; g < 8; ++g) {
                    const float4 qn = *reinterpret_cast<const float4*>(prow + 4 * min(g + 1, 7));   // next group: hides the LDS latency
                    float y0 = bandpass_step(P, st, q.x); float z0 = preemph_step(P, st, q.x);
                    float y1 = bandpass_step(P, st, q.y); float z1 = preemph_step(P, st, q.y);
                    float y2 = bandpass_step(P, st, q.z); float z2 = preemph_step(P, st, q.z);
                    float y3 = bandpass_step(P, st, q.w); float z3 = preemph_step(P, st, q.w);
                    acc = fmaf(y0, y0, acc); acc2 = fmaf(y1, y1, acc2);
                    acc = fmaf(y2, y2, acc); acc2 = fmaf(y3, y3, acc2);
                    *reinterpret_cast<float4*>(prow + 4 * g) = make_float4(z0, z1, z2, z3);
                    q = qn;
                }
                acc += acc2;
            } else if (tp < len) {
#pragma unroll 1
                for (int g = 0; g < 8; ++g) {
                    float4 q = *reinterpret_cast<const float4*>(prow + 4 * g);
                    float v[4] = {q.x, q.y, q.z, q.w};
                    float zz[4];
                    const int tg = tp + 4 * g;
                    const bool edge = (tg < 16) || (tg + 4 > len - 16);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        float yv = bandpass_step(P, st, v[j]);
                        float xv = edge ? v[j] * taper_weight(tg + j, len) : v[j];
                        zz[j] = preemph_step(P, st, xv);
                        if (tg + j < t_end) acc = fmaf(yv, yv, acc);
                    }
                    *reinterpret_cast<float4*>(prow + 4 * g) = make_float4(zz[0], zz[1], zz[2], zz[3]);
                }
            }
            if (owned) acc_d += (double)acc;
            __syncwarp();
            if (owned) {
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const int a = tp + col;
                    const int rl = rlen(k);
                    float4 q = *reinterpret_cast<const float4*>(tbase + 4 * k * kFpRingPitch + sp);
                    float* d = dst_row(k) + tp;
                    if (a + 4 <= rl) {
                        *reinterpret_cast<float4*>(d) = q;
                    } else {
                        if (a < rl) d[0] = q.x;
                        if (a + 1 < rl) d[1] = q.y;
                        if (a + 2 < rl) d[2] = q.z;
                    }
                }
            }
            __syncwarp();
        }
        const int freed = sp;                                // rotate: the drained slot receives tile i + 2
        sp = sc; sc = sn; sn = freed;
    }
    if (sig_ok) partial[((int64_t)half * batch + my_item) * nchunks + c] = acc_d;
}

}  // namespace fsem
