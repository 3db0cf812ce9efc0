// SDR correlation lags on the 5th-generation tensor cores (tcgen05.mma, accumulators in TMEM) -- sm_100a only.
//
// SDR.py:34-49 needs, per item, r[l] = sum_t c[t] c[t+l] and b[l] = sum_t c[t] d[t+l] for l = 0..511: 1024 MACs per
// sample, a genuinely dense contraction (unlike the Bark / third-octave projections of the hot path, which are disjoint
// 0/1 segment sums and stay on CUDA cores).  Cut time into rows of 128 samples, t = 128 m + i:
//     G[i][j] = sum_m c[128 m + i] * x[128 m + j]        (x = c or d; i < 128, j < 640)
// is a GEMM with M = 128 (i), N = j, K = m (n / 128 rows), and r[l] = sum_i G[i][i + l] (sum along diagonals).
//
// Precision: the operands are split into two bf16 terms, x = hi + lo (16 significant bits), and three products are
// accumulated in fp32: hi*hi + hi*lo + lo*hi (the dropped lo*lo is 2^-16 of the result).  Relative to r[0] the error of
// a lag is ~2^-17 / sqrt(n) from the split plus the fp32 accumulation of 1250 row products -- the same order as the
// float32 correlations the reference itself keeps (SDR.py:78-79); the SIMT kernel this replaces summed fp32 per
// 2048-sample tile and fp64 across tiles.
//
// One CTA = (item, correlation r | b, lag half): 128 x 384 accumulator (lags 256 h .. 256 h + 255 need columns
// j = 256 h + [0, 384)) in 384 of the SM's 512 TMEM columns.  Per k-step (16 rows = 2048 samples) all 256 threads build
// the operand tiles in shared memory in the canonical MN-major, no-swizzle UMMA layout (8 k-rows x 16 bytes core
// matrices): 8 consecutive samples are read with two 16-byte loads, split into hi / lo bf16 and stored as one 16-byte
// row of a core matrix of each tile.  One thread then issues six tcgen05.mma (3 products x N = 256 + 128) that read the
// tiles through shared-memory descriptors and accumulate in TMEM; tcgen05.commit on an mbarrier releases the stage.  Four
// stages keep the tensor pipe busy while the next tiles are being built.  Epilogue: every warp reads its 32 TMEM lanes
// (tcgen05.ld 32x32b) and adds G[i][j] into its private lag array at l = j - i; the eight arrays are summed in a fixed
// order (deterministic) into the double partials the Levinson kernel reads.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>

#include "fsem_common.cuh"

namespace fsem {

constexpr int kTcM = 128;                 // i values per row of time (UMMA M)
constexpr int kTcK = 16;                  // rows m per k-step (UMMA K for bf16)
constexpr int kTcN = 384;                 // columns j per CTA: 256 lags + 128
constexpr int kTcLagsPerCta = 256;
constexpr int kTcStages = 3;
constexpr int kTcProducers = 512;         // 16 producer warps build the operand tiles (and run the epilogue)
constexpr int kTcUnits = 1024 / kTcProducers;   // 8-sample units per producer thread and k-step
constexpr int kTcThreads = kTcProducers + 64;   // + one warp whose lane 0 issues the tcgen05.mma + one warp that issues the bulk copies
constexpr int kTcCore = 128;              // bytes of a core matrix: 8 k-rows x 16 bytes (8 bf16 along MN)
constexpr int kTcABytes = 2 * (kTcM / 8) * kTcCore;          // [k1 = 2][mn1 = 16]      4096
constexpr int kTcBBytes = 2 * (kTcN / 8) * kTcCore;          // [k1 = 2][mn1 = 48]     12288
constexpr int kTcStageBytes = 2 * kTcABytes + 2 * kTcBBytes; // A_hi | A_lo | B_hi | B_lo   32768
constexpr int kTcTmemCols = 512;

__device__ __forceinline__ uint64_t umma_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    // cute::UMMA::SmemDescriptor: start address [0,14), leading byte offset [16,30), stride byte offset [32,46) (all
    // without their 4 LSB), version = 1 at [46,48), base offset 0, layout type SWIZZLE_NONE = 0 at [61,64)
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}
__host__ __device__ constexpr uint32_t umma_idesc_bf16_mn(int m, int n) {
    // cute::UMMA::InstrDescriptor: c_format F32 = 1 at [4,6), a/b format BF16 = 1 at [7,10) / [10,13), a/b major MN = 1 at
    // bits 15 / 16, n >> 3 at [17,23), m >> 4 at [24,29)
    return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(n >> 3) << 17) |
           ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void umma_bf16_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }

// eight consecutive samples -> (hi, lo) bf16 rows of 16 bytes each
__device__ __forceinline__ void split_bf16x8(const float (&x)[8], uint4& hi, uint4& lo) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        const __nv_bfloat162 hh = __floats2bfloat162_rn(x[2 * p], x[2 * p + 1]);
        const float2 back = __bfloat1622float2(hh);
        const __nv_bfloat162 ll = __floats2bfloat162_rn(x[2 * p] - back.x, x[2 * p + 1] - back.y);
        h[p] = *reinterpret_cast<const uint32_t*>(&hh);
        l[p] = *reinterpret_cast<const uint32_t*>(&ll);
    }
    hi = make_uint4(h[0], h[1], h[2], h[3]);
    lo = make_uint4(l[0], l[1], l[2], l[3]);
}

// Raw float32 staging of one k-step, filled by THREE TMA tensor copies (cp.async.bulk.tensor.3d: no LSU traffic, one
// elected thread): the 16 rows of the A segment (128 samples) and of the two halves of the B segment (2 x 192 samples).
// The tensor map views a signal buffer as [item][row m][sample t] with OVERLAPPING rows (row pitch 128 samples); the
// boxes are 4 samples wider than needed (132 / 196 floats) so that the row pitch in shared memory is = 4 (mod 32) words:
// the eight lanes of a quarter-warp, which read eight different rows, then hit different banks.
constexpr int kTcBoxA = kTcM + 4;                              // 132 floats
constexpr int kTcBoxB = kTcN / 2 + 4;                          // 196 floats
constexpr int kTcRawAPitch = kTcBoxA * 4;                      // 528 bytes
constexpr int kTcRawBPitch = kTcBoxB * 4;                      // 784 bytes
constexpr int kTcRawABytes = kTcK * kTcRawAPitch;              // 8448   (multiple of 128: TMA destination alignment)
constexpr int kTcRawBBytes = kTcK * kTcRawBPitch;              // 12544
constexpr int kTcRawBytes = kTcRawABytes + 2 * kTcRawBBytes;   // 33536
static_assert(kTcRawABytes % 128 == 0 && kTcRawBBytes % 128 == 0, "TMA destinations must be 128-byte aligned");
constexpr int kTcRawStages = 3;
constexpr int kTcRaccBytes = (kTcProducers / 32) * kTcLagsPerCta * 4;
constexpr size_t kTcDynSmem = (size_t)kTcStages * kTcStageBytes + (size_t)kTcRawStages * kTcRawBytes + kTcRaccBytes + 1024;
constexpr int kTcMapDim0 = 256 + kTcN / 2 + kTcBoxB;           // 644: largest sample coordinate a box may touch + 1

__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
        ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(bar) : "memory");
}

__global__ void __launch_bounds__(kTcThreads, 1)
sdr_corr_tc_kernel(const float* __restrict__ clean, const float* __restrict__ deg, const int32_t* __restrict__ lengths,
                   int64_t batch, int64_t n, int64_t stride, int use_tma,
                   const __grid_constant__ CUtensorMap map_a,      // clean signal, box 132 x 16
                   const __grid_constant__ CUtensorMap map_bc,     // clean signal, box 196 x 16
                   const __grid_constant__ CUtensorMap map_bd,     // degraded signal, box 196 x 16
                   double* __restrict__ partial /* [batch][1][2][512] */) {
    extern __shared__ __align__(1024) unsigned char s_tc[];
    __shared__ __align__(8) unsigned long long s_empty[kTcStages];   // tile stage free again: tcgen05.commit of the MMAs that read it
    __shared__ __align__(8) unsigned long long s_full[kTcStages];    // tile stage built: one arrival per producer warp
    __shared__ __align__(8) unsigned long long s_raw_full[kTcRawStages];    // raw rows landed (bulk-copy transaction bytes)
    __shared__ __align__(8) unsigned long long s_raw_empty[kTcRawStages];   // raw rows consumed: one arrival per producer warp
    __shared__ __align__(8) unsigned long long s_done;
    __shared__ uint32_t s_tmem;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t item = blockIdx.x >> 2;
    const int corr = (blockIdx.x >> 1) & 1;            // 0: r (clean x clean), 1: b (clean x degraded)
    const int half = blockIdx.x & 1;                   // lags 256 * half .. 256 * half + 255
    const int len = item_length(lengths, item, n);
    const float* __restrict__ c = clean + item * stride;
    const float* __restrict__ x = (corr ? deg : clean) + item * stride;
    unsigned char* raw0 = s_tc + (size_t)kTcStages * kTcStageBytes;
    float* racc = reinterpret_cast<float*>(raw0 + (size_t)kTcRawStages * kTcRawBytes);     // [producer warps][256 lags]
    const uint32_t smem0 = smem_u32(s_tc);

    if (tid == 0) {
        for (int s = 0; s < kTcStages; ++s) {
            mbar_init(smem_u32(&s_empty[s]), 1);
            mbar_init(smem_u32(&s_full[s]), kTcProducers / 32);
        }
        for (int s = 0; s < kTcRawStages; ++s) {
            mbar_init(smem_u32(&s_raw_full[s]), 1);
            mbar_init(smem_u32(&s_raw_empty[s]), kTcProducers / 32);
        }
        mbar_init(smem_u32(&s_done), 1);
        mbar_fence_init();
    }
    const bool is_mma_warp = warp == kTcProducers / 32;
    const bool is_copy_warp = warp == kTcProducers / 32 + 1;
    if (is_mma_warp) {                                 // the MMA warp allocates (and later frees) the accumulator columns
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(kTcTmemCols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    for (int i = tid; i < (kTcProducers / 32) * kTcLagsPerCta; i += kTcThreads) racc[i] = 0.f;
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem = s_tmem;

    const int rows = (len + kTcM - 1) / kTcM;                          // rows m with at least one sample
    const int ksteps = (rows + kTcK - 1) / kTcK;
    // leading k-steps whose 16 rows lie completely inside the signal (A and B segment): staged by bulk copies; the
    // last one or two k-steps (and everything, for rows that are not 16-byte aligned) read global memory directly
    int bulk_steps = 0;
    if (use_tma) {
        // largest admissible first row m0: all 16 rows inside the item's valid samples AND inside the tensor map (rows
        // beyond its dim 1 would be zero-filled)
        const int last_ok = min((len - kTcLagsPerCta * half - kTcN) / kTcM, (int)((n - kTcMapDim0) / kTcM)) - (kTcK - 1);
        if (len >= kTcLagsPerCta * half + kTcN && last_ok >= 0) bulk_steps = min(ksteps, last_ok / kTcK + 1);
    }
    constexpr uint32_t kIdesc256 = umma_idesc_bf16_mn(kTcM, 256);
    constexpr uint32_t kIdesc128 = umma_idesc_bf16_mn(kTcM, 128);

    if (is_copy_warp) {
        // ---- copy warp: one lane issues the three tensor copies of every staged k-step
        if (lane == 0) {
            const CUtensorMap* mb = corr ? &map_bd : &map_bc;
            for (int s = 0; s < bulk_steps; ++s) {
                const int rs = s % kTcRawStages;
                if (s >= kTcRawStages) mbar_wait(smem_u32(&s_raw_empty[rs]), (uint32_t)((s / kTcRawStages - 1) & 1));
                const uint32_t bar = smem_u32(&s_raw_full[rs]);
                mbar_arrive_expect_tx(bar, kTcRawBytes);
                const uint32_t dst = smem_u32(raw0 + (size_t)rs * kTcRawBytes);
                const int m0 = s * kTcK;
                tma_load_3d(dst, &map_a, 0, m0, (int)item, bar);
                tma_load_3d(dst + kTcRawABytes, mb, kTcLagsPerCta * half, m0, (int)item, bar);
                tma_load_3d(dst + kTcRawABytes + kTcRawBBytes, mb, kTcLagsPerCta * half + kTcN / 2, m0, (int)item, bar);
            }
        }
    } else if (is_mma_warp) {
        if (lane == 0) {
            for (int s = 0; s < ksteps; ++s) {
                const int stage = s % kTcStages;
                mbar_wait(smem_u32(&s_full[stage]), (uint32_t)((s / kTcStages) & 1));
                tcgen05_fence_after();
                const uint32_t sa = smem0 + stage * kTcStageBytes;
                const uint32_t a_hi = sa, a_lo = sa + kTcABytes, b_hi = sa + 2 * kTcABytes, b_lo = b_hi + kTcBBytes;
                constexpr uint32_t kALbo = (kTcM / 8) * kTcCore, kBLbo = (kTcN / 8) * kTcCore, kSbo = kTcCore;
                const uint32_t pa[3] = {a_hi, a_hi, a_lo};
                const uint32_t pb[3] = {b_hi, b_lo, b_hi};
#pragma unroll
                for (int p = 0; p < 3; ++p) {
                    const uint32_t acc = (s > 0 || p > 0) ? 1u : 0u;
                    const uint64_t ad = umma_smem_desc(pa[p], kALbo, kSbo);
                    umma_bf16_ss(tmem, ad, umma_smem_desc(pb[p], kBLbo, kSbo), kIdesc256, acc);
                    umma_bf16_ss(tmem + 256, ad, umma_smem_desc(pb[p] + 32 * kTcCore, kBLbo, kSbo), kIdesc128, acc);
                }
                umma_commit(smem_u32(&s_empty[stage]));                 // arrives when these MMAs have finished
                if (s + 1 == ksteps) umma_commit(smem_u32(&s_done));
            }
        }
    } else {
        // ---- producers.  Units of a thread (fixed): unit q -> (A | B, core-matrix row k0, group mn1, k1)
        int u_off[kTcUnits];   // byte offset of the unit's 16-byte row inside its tile
        int u_raw[kTcUnits];   // byte offset of the unit's 8 samples inside a raw stage
        int u_t[kTcUnits];     // sample index of the unit's first sample relative to 128 * m0
        bool u_a[kTcUnits];
#pragma unroll
        for (int q = 0; q < kTcUnits; ++q) {
            const int u = tid + kTcProducers * q;                       // 0..1023: A units first, then B units
            const bool is_a = u < 256;
            const int ub = is_a ? u : u - 256;
            const int k0 = ub & 7;
            const int rest = ub >> 3;
            const int groups = is_a ? kTcM / 8 : kTcN / 8;              // core matrices along MN
            const int k1 = rest / groups, mn1 = rest - k1 * groups;
            // tile layout [k1][mn1] core matrices, 16-byte row k0 inside: the 8 lanes of a quarter-warp (k0 = 0..7, same
            // mn1) write one contiguous core matrix -> conflict-free
            u_off[q] = (k1 * groups + mn1) * kTcCore + k0 * 16;
            u_t[q] = kTcM * (8 * k1 + k0) + 8 * mn1 + (is_a ? 0 : kTcLagsPerCta * half);
            u_raw[q] = is_a ? (8 * k1 + k0) * kTcRawAPitch + 32 * mn1
                            : (mn1 < kTcN / 16 ? kTcRawABytes + (8 * k1 + k0) * kTcRawBPitch + 32 * mn1
                                               : kTcRawABytes + kTcRawBBytes + (8 * k1 + k0) * kTcRawBPitch + 32 * (mn1 - kTcN / 16));
            u_a[q] = is_a;
        }
        for (int s = 0; s < ksteps; ++s) {
            const int stage = s % kTcStages;
            float v[kTcUnits][8];
            if (s < bulk_steps) {
                const int rs = s % kTcRawStages;
                mbar_wait(smem_u32(&s_raw_full[rs]), (uint32_t)((s / kTcRawStages) & 1));
                const unsigned char* raw = raw0 + (size_t)rs * kTcRawBytes;
#pragma unroll
                for (int q = 0; q < kTcUnits; ++q) {
                    const float4 p0 = *reinterpret_cast<const float4*>(raw + u_raw[q]);
                    const float4 p1 = *reinterpret_cast<const float4*>(raw + u_raw[q] + 16);
                    v[q][0] = p0.x; v[q][1] = p0.y; v[q][2] = p0.z; v[q][3] = p0.w;
                    v[q][4] = p1.x; v[q][5] = p1.y; v[q][6] = p1.z; v[q][7] = p1.w;
                }
            } else {
                const int m0 = s * kTcK;
#pragma unroll
                for (int q = 0; q < kTcUnits; ++q) {
                    const int t0 = kTcM * m0 + u_t[q];
                    const float* __restrict__ src = (u_a[q] ? c : x) + t0;
#pragma unroll
                    for (int e = 0; e < 8; ++e) v[q][e] = (t0 + e < len) ? __ldg(src + e) : 0.f;
                }
            }
            uint4 hi[kTcUnits], lo[kTcUnits];
#pragma unroll
            for (int q = 0; q < kTcUnits; ++q) split_bf16x8(v[q], hi[q], lo[q]);
            if (s >= kTcStages)                                         // the MMAs that read this stage have completed
                mbar_wait(smem_u32(&s_empty[stage]), (uint32_t)((s / kTcStages - 1) & 1));
            unsigned char* st = s_tc + (size_t)stage * kTcStageBytes;
#pragma unroll
            for (int q = 0; q < kTcUnits; ++q) {
                unsigned char* base = st + (u_a[q] ? 0 : 2 * kTcABytes);
                const int lo_off = u_a[q] ? kTcABytes : kTcBBytes;
                *reinterpret_cast<uint4*>(base + u_off[q]) = hi[q];
                *reinterpret_cast<uint4*>(base + lo_off + u_off[q]) = lo[q];
            }
            // One proxy fence serves both hand-overs: the generic-proxy tile stores become visible to the tensor core, and
            // the generic-proxy READS of the raw rows are ordered before the TMA writes that refill them (without it the
            // refill raced with the reads: results changed from run to run).
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(smem_u32(&s_full[stage]));
                if (s < bulk_steps) mbar_arrive(smem_u32(&s_raw_empty[s % kTcRawStages]));
            }
        }
    }
    __syncwarp();
    if (ksteps > 0 && warp < kTcProducers / 32) {
        mbar_wait(smem_u32(&s_done), 0);
        tcgen05_fence_after();
        // ---- epilogue: warp w reads TMEM lanes 32 (w % 4) .. +31 (rows i) and the columns [kCols (w / 4), + kCols)
        constexpr int kParts = kTcProducers / 128;                      // column parts (4 with 16 producer warps)
        constexpr int kCols = kTcN / kParts;                            // 96
        static_assert(kCols % 32 == 0, "column parts are read in chunks of 32");
        const int quad = warp & 3;
        const int i_row = 32 * quad + lane;
        float* mine = racc + warp * kTcLagsPerCta;
#pragma unroll 1
        for (int cb = 0; cb < kCols / 32; ++cb) {
            const int c0 = kCols * (warp >> 2) + 32 * cb;
            uint32_t r[32];
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                  "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
                  "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
                  "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                : "r"(tmem + ((uint32_t)(32 * quad) << 16) + (uint32_t)c0));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int q = 0; q < 32; ++q) {
                const int l = c0 + q - i_row;                           // lag (relative to 256 * half) of G[i][c0 + q]
                if (l >= 0 && l < kTcLagsPerCta) mine[l] += __uint_as_float(r[q]);   // lanes hit distinct addresses
                __syncwarp();
            }
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (tid < kTcLagsPerCta) {   // fixed-order sum of the eight per-warp arrays -> the double partials of sdr_solve_kernel
        float acc = 0.f;
#pragma unroll
        for (int w = 0; w < kTcProducers / 32; ++w) acc += racc[w * kTcLagsPerCta + tid];
        partial[(item * 2 + corr) * (int64_t)512 + kTcLagsPerCta * half + tid] = (double)acc;
    }
    if (is_mma_warp) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kTcTmemCols));
    }
}

}  // namespace fsem
