// "Sliced-integer mode" of the predictive variance: the FP64 contraction V = K* L^-T evaluated EXACTLY-ROUNDED on the
// 5th-generation tensor cores' integer path (tcgen05.mma kind::i8, exact int32 accumulators in TMEM).
//
// Each operand is scaled by a power of two (L^-1: one per row; K*: one for the whole matrix, k <= outputscale) and cut
// into S signed 7-bit slices by repeated round-to-nearest
//       x * scale = sum_{j<S} I_j 2^{-7(j+1)} + rho,      |I_j| <= 64,  |rho| <= 2^{-7S-1},
// so that            V[m][i] = (1 / scale_i scale_K) sum_{g<S} 2^{-7(g+2)} sum_{a+b=g} <I_a(L^-1 row i), I_b(K* row m)>
// up to the dropped products a+b >= S.  Every inner product is an integer GEMM whose int32 accumulation is exact
// (|sum| <= 64*64*K*S < 2^31 for K <= 32768); the S diagonals are combined in FP64 in the epilogue.  With S = 5 the
// predictive std agrees with the FP64 path to ~5e-7 relative (measured, N = 8192, noise 1e-2; 7e-6 at noise 1e-4), S = 6 to
// ~5e-9 (4e-8) -- inside the 1e-6 bar of the FP64 mode, at the integer tensor-core rate instead of the DMMA rate
// (the variance contraction is the N^2 FLOP/point term that bounds predict).
//
// Two digit formats (template parameter BITS):
//   BITS = 7: |I_j| <= 64 by repeated round-to-nearest (the formulas above); int32 sums safe for K <= 32768.
//   BITS = 8: balanced radix-256 digits I_j in [-128, 127] of  V = rint(x * scale * 2^{8S})  (one rounding, 8S bits kept):
//             x * scale = sum_j I_j 2^{-8(j+1)},  bytes of (V + sum_j 128 * 256^j) xor 0x80.  The scale keeps |x * scale| <= 127.4/256
//             so the top digit fits.  |sum| <= 128*128*K*S < 2^31 for K <= 16384 whatever the digits are -- above, up to 32768, when
//             the digits of the sliced L^-1 allow it (i8_digit_bound below).  Five 8-bit slices (40 bits, 15 products)
//             replace six 7-bit ones (42 bits, 21 products); six 8-bit slices carry 48 bits.
// Plane layout of the predict operands (K* chunk and L^-1): inside every 128-block of the contraction index the entry of
// column c sits at byte (c % 16) * 8 + c / 16 -- the same permutation on both operands leaves every inner product unchanged
// and lets the K* builder (gpx_kmat.cu), whose threads own columns tx + 16 b, store 8 consecutive bytes per slice straight
// from registers.
//
// Contents: slice_i8_kernel / slice_i8_rows_kernel (FP64 -> int8 planes), trmm_sumsq_i8_kernel / trmm_sumsq_i8p_kernel (predict:
// sum of squares epilogue; one tile per cluster / persistent), gemm_i8_kernel (general NT product with a store, mirror or accumulate
// epilogue: K_hat^-1 for the gradient, the merge levels of the triangular inverse, the Cholesky trailing updates and the
// validation covariance of training loops), i8_digit_bound_kernel, peak_i8mma_kernel (roofline probe).
//
// Tile: 128 test points (MMA M = TMEM lanes) x 64 rows of L^-1 (per slice), S accumulators of 64 TMEM columns, one per
// diagonal g.  The S slices of the L^-1 tile are STACKED along MMA N in shared memory, so for the K* slice a one
// instruction with N = 64 (S - a) (cut at 256) multiplies it with the L^-1 slices b = 0 .. S-1-a and lands in the
// accumulators g = a .. S-1, which are contiguous in TMEM: S = 5 needs 6 MMAs per 32-deep k step instead of 15, and the
// K* slices are read from shared memory 6 times instead of 15.  The first k step of K* slice 0 covers all S accumulators
// and overwrites them; every later MMA accumulates.  K runs over the training index in BKB-byte slices (SWIZZLE_64B: three stages of 60-72 KB).
// Warp roles as in gpx_tf32.cu: warp 0 TMA producer, warp 1 TMEM allocator + MMA issuer, warps 2..9 epilogue.
#include "gpx_common.cuh"
#include "gpx_tma.cuh"
#include "gpx_tc.cuh"
#include <stdlib.h>

namespace gpx {

constexpr int I_M = 128;        // test points per tile
constexpr int I_N = 64;         // L^-1 rows per tile (and per slice in the stacked B operand)
constexpr int I_BITS = 7;
constexpr int I_THREADS = 320;
constexpr int I_BAND = 8;       // m tiles that share L^-1 tiles through L2
constexpr int I_TMEM_COLS = 512;
constexpr int I_SMAX = 6;
constexpr int I_RING = 8;       // work-item slots of the dynamic scheduler (no role is ever 8 items behind the scheduler)

template <int S, int BKB>
struct I8Cfg {
    static constexpr int STAGES = (BKB == 64) ? 3 : 1;
    static constexpr int A_BYTES = I_M * BKB;
    static constexpr int B_BYTES = I_N * BKB;
    static constexpr int STAGE_BYTES = S * (A_BYTES + B_BYTES);
    // tiles | full[STAGES] empty[STAGES] acc_full acc_empty | tmem slot | inverse scales [64] | partial sums [2][128] |
    // item_full[I_RING] | item ring [I_RING]
    static constexpr size_t SMEM_BYTES = 1024 + size_t(STAGES) * STAGE_BYTES + (2 * STAGES + 3) * 8 + I_N * 8 + 2 * I_M * 8 + I_RING * 12 + 32;
};

// power-of-two scale that brings |x| <= amax into [-1/2, 1/2]:  returns e with amax < 2^e, scale = 2^-(e+1)
__device__ __forceinline__ int scale_exponent(double amax) {
    if (!(amax > 0.0)) return 0;
    int e;
    frexp(amax, &e);            // amax = f 2^e, f in [1/2, 1)
    return e;
}
// ---- FP64 -> S int8 slices -------------------------------------------------------------------------------
// One CTA per row.  planes[j][row][perm128(col)] (row pitch = cols bytes, plane pitch = plane_rows * cols bytes; cols % 128 == 0).
// tri != 0: `src` is the mirrored inverse S (L^-1 below, L^-T above the diagonal blocks): entries right of the diagonal
// are sliced as zero and the row's own maximum sets its scale, whose inverse is stored in inv_scale[row].
// tri == 0: one scale for the whole matrix from *amax_ptr (K* <= outputscale).
template <int S, int BITS>
__global__ void __launch_bounds__(256) slice_i8_kernel(const double* __restrict__ src, int cols, long ld_src,
                                                       int8_t* __restrict__ planes, long plane_rows, int tri,
                                                       double* __restrict__ inv_scale, const double* __restrict__ amax_ptr) {
    const int row = blockIdx.x;
    const double* x = src + (long)row * ld_src;
    const int kmax = tri ? row + 1 : cols;          // entries [0, kmax) are live
    __shared__ double red[8];
    __shared__ int s_e;
    if (tri) {
        double a = 0.0;
        for (int c = threadIdx.x; c < kmax; c += 256) a = fmax(a, fabs(x[c]));
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) a = fmax(a, __shfl_xor_sync(0xffffffffu, a, o));
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = a;
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int w = 1; w < 8; ++w) a = fmax(a, red[w]);
            s_e = scale_exponent_b<BITS>(a);
            inv_scale[row] = ldexp(1.0, s_e + 1);
        }
        __syncthreads();
    } else {
        if (threadIdx.x == 0) s_e = scale_exponent_b<BITS>(*amax_ptr);
        __syncthreads();
    }
    const double scale = ldexp(1.0, -(s_e + 1) + (BITS == 7 ? I_BITS : 0));   // 7 bit: first slice = rint(x * scale * 2^7)
    const long pitch = plane_rows * (long)cols;
    int8_t* out = planes + (long)row * cols;
    const int nblk = (tri ? min(cols, round_up(row + 1, TILE)) : cols) / TILE;   // the MMA tiles never read beyond the row's 128-block
    // work item = (128-block, tx): the entries of columns blk*128 + tx + 16 u, u < 8, land in 8 consecutive bytes
    for (int w = threadIdx.x; w < nblk * 16; w += 256) {
        const int c0 = (w >> 4) * TILE + (w & 15);
        double r[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) r[u] = (c0 + 16 * u < kmax) ? x[c0 + 16 * u] * scale : 0.0;
        int8_t* dst = out + (w >> 4) * TILE + (w & 15) * 8;
        if (BITS == 7) {
#pragma unroll
            for (int j = 0; j < S; ++j) {
                uint32_t lo = 0, hi = 0;
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const double q = rint(r[u]);
                    r[u] = (r[u] - q) * 128.0;                       // exact: |r - q| <= 1/2
                    const uint32_t b = (uint32_t)(__double2int_rn(q) & 0xff);
                    if (u < 4) lo |= b << (8 * u); else hi |= b << (8 * (u - 4));
                }
                *reinterpret_cast<uint2*>(dst + j * pitch) = make_uint2(lo, hi);
            }
        } else {
            unsigned long long dg[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) dg[u] = balanced_digits<S>(r[u]);
#pragma unroll
            for (int j = 0; j < S; ++j) {
                uint32_t lo = 0, hi = 0;
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const uint32_t b = (uint32_t)(dg[u] >> (8 * (S - 1 - j))) & 0xffu;
                    if (u < 4) lo |= b << (8 * u); else hi |= b << (8 * (u - 4));
                }
                *reinterpret_cast<uint2*>(dst + j * pitch) = make_uint2(lo, hi);
            }
        }
    }
}

// ---- descriptors --------------------------------------------------------------------------------------------
template <int BKB>
__device__ __forceinline__ uint64_t umma_desc_kmajor(uint32_t smem_addr) {
    // cute::UMMA::SmemDescriptor, K-major, rows of BKB bytes: 8-row groups 8*BKB bytes apart (SBO), LBO unused (=1),
    // version 1, layout_type 2 = SWIZZLE_128B / 4 = SWIZZLE_64B
    return (uint64_t)((smem_addr & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)((8 * BKB) >> 4) << 32) |
           ((uint64_t)1 << 46) | ((uint64_t)(BKB == 128 ? 2 : 4) << 61);
}
// cute::UMMA::InstrDescriptor: D = S32 (2 at bit 4), A = B = signed int8 (1 at bits 7 and 10), both K-major
__device__ __forceinline__ uint32_t i8_idesc(int n) {
    return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(I_M >> 4) << 24);
}
__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}

// ---- the kernel ---------------------------------------------------------------------------------------------
// grid.x = (Np/64) * (Mc/128) tiles; qpart[it][m] (ld = Mc) = sum over the 64 rows of tile `it` of V[m][i]^2.
// tmA: K* planes as one [S*Mc][Np] byte matrix (box BKB x 128); tmB: L^-1 planes as [S*Np][Np] (box BKB x 64).
//
// MC = true (default): CTAs run as clusters of two on the SAME 128 test points and on two neighbouring row tiles
// (it = 2p, 2p+1).  Each CTA fetches half (64 points) of every K* slice tile and TMA-multicasts it into both CTAs, so the
// L2 serves the K* bytes -- two thirds of a stage -- once per pair: 60 -> 40 KB of L2 reads per CTA and stage (S = 5).
// Both CTAs run the k extent of the odd tile (one extra stage of zero slices for the even one); a stage is refilled
// only when both tensor cores have retired it (tcgen05.commit multicast onto both `empty` barriers, count 2).
template <int S, int BKB, bool MC, int BITS>
__global__ void __launch_bounds__(I_THREADS, 1)
trmm_sumsq_i8_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, int Np, int Mc,
                     const __grid_constant__ CUtensorMap tmAh, const double* __restrict__ inv_scale,
                     const double* __restrict__ amax_ptr, double* __restrict__ qpart) {
    using C = I8Cfg<S, BKB>;
    extern __shared__ unsigned char smem_raw[];
    unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(base + size_t(C::STAGES) * C::STAGE_BYTES);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * C::STAGES + 1);
    double* sinv = reinterpret_cast<double*>(bars + 2 * C::STAGES + 2);      // [64]
    double* red = sinv + I_N;                                                 // [2][128]
    const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + C::STAGES), acc_full = smem_u32(bars + 2 * C::STAGES);
    const uint32_t tile0 = smem_u32(base);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    // bands of I_BAND m tiles; inside a band the row tiles run heaviest (largest k extent) first
    const uint32_t crank = MC ? cluster_ctarank() : 0u;
    int it, mb, kend;
    {
        const int n_mb = Mc / I_M, n_it = MC ? Np / (2 * I_N) : Np / I_N;      // MC: n_it counts PAIRS of row tiles
        const int id = MC ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
        const int per_band = I_BAND * n_it;
        const int band = id / per_band;
        const int left = n_mb - band * I_BAND;
        const int width = left < I_BAND ? left : I_BAND;
        const int rem = id - band * per_band;
        it = n_it - 1 - rem / width;
        mb = band * I_BAND + rem % width;
        // entries right of the diagonal are zero slices up to the end of the row's 128-block, so a pair runs the odd tile's extent
        // (the k index is permuted inside 128-blocks, so extents are whole 128-blocks)
        kend = MC ? (it + 1) * 2 * I_N : round_up((it + 1) * I_N, TILE);
        if (MC) it = 2 * it + (int)crank;
    }
    const int kstart = 0;
    const int nk = (kend - kstart) / BKB;

    if (tid == 0) {
        for (int s = 0; s < C::STAGES; ++s) {
            mbar_init(full0 + 8 * s, 1);
            mbar_init(empty0 + 8 * s, MC ? 2 : 1);
        }
        mbar_init(acc_full, 1);
        mbar_fence_init();
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(tmem_slot)), "n"(I_TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    if (tid < I_N) {
        // V = (sum_g acc_g 2^{-BITS g}) * 2^{-2 BITS} / (scale_i * scale_K):  all powers of two, exact
        const int ek = scale_exponent_b<BITS>(*amax_ptr);
        sinv[tid] = inv_scale[it * I_N + tid] * ldexp(1.0, ek + 1 - 2 * BITS);
    }
    tc_fence_before();
    __syncthreads();
    if (MC) cluster_sync_all();        // the peer's barriers must be initialised before anything is multicast at them
    tc_fence_after();
    const uint32_t tmem_d = *tmem_slot;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmB);
            for (int kt = 0; kt < nk; ++kt) {
                const int s = kt % C::STAGES, round = kt / C::STAGES;
                if (round > 0) mbar_wait(empty0 + 8 * s, (round - 1) & 1);
                mbar_arrive_expect_tx(full0 + 8 * s, C::STAGE_BYTES);
                const uint32_t dst = tile0 + s * C::STAGE_BYTES;
                const int k0 = kstart + kt * BKB;
                if (MC) {   // my 64 points of every K* slice tile, delivered to both CTAs of the pair
#pragma unroll
                    for (int j = 0; j < S; ++j)
                        tma_load_2d_multicast(dst + j * C::A_BYTES + crank * (C::A_BYTES / 2), &tmAh, k0,
                                              j * Mc + mb * I_M + (int)crank * (I_M / 2), full0 + 8 * s, 3);
                } else {
#pragma unroll
                    for (int j = 0; j < S; ++j) tma_load_2d(dst + j * C::A_BYTES, &tmA, k0, j * Mc + mb * I_M, full0 + 8 * s);
                }
#pragma unroll
                for (int j = 0; j < S; ++j)
                    tma_load_2d(dst + S * C::A_BYTES + j * C::B_BYTES, &tmB, k0, j * Np + it * I_N, full0 + 8 * s);
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            for (int kt = 0; kt < nk; ++kt) {
                const int s = kt % C::STAGES, round = kt / C::STAGES;
                mbar_wait(full0 + 8 * s, round & 1);
                tc_fence_after();
                const uint32_t st = tile0 + s * C::STAGE_BYTES;
#pragma unroll
                for (int kk = 0; kk < BKB / 32; ++kk) {
                    const uint64_t adv = (uint64_t)((kk * 32) >> 4);     // 32 int8 along K inside the swizzle atom
#pragma unroll
                    for (int a = 0; a < S; ++a) {
                        const uint64_t da = umma_desc_kmajor<BKB>(st + a * C::A_BYTES) + adv;
                        const int ncols = (S - a) * I_N;                  // L^-1 slices 0 .. S-1-a -> diagonals a .. S-1
#pragma unroll
                        for (int c0 = 0; c0 < ncols; c0 += 256) {
                            const int w = (ncols - c0) < 256 ? (ncols - c0) : 256;
                            const uint64_t db = umma_desc_kmajor<BKB>(st + S * C::A_BYTES + (c0 / I_N) * C::B_BYTES) + adv;
                            // the very first k step of K* slice 0 touches every accumulator once: it overwrites (no zero fill)
                            umma_i8(tmem_d + (uint32_t)(a * I_N + c0), da, db, i8_idesc(w), (kt | kk | a) != 0 ? 1u : 0u);
                        }
                    }
                }
                if (MC) umma_commit_multicast(empty0 + 8 * s, 3);     // both producers wait for both tensor cores
                else umma_commit(empty0 + 8 * s);
            }
            umma_commit(acc_full);
        }
    } else {
        // ===== epilogue warps 2..9: TMEM lane quadrant = warp % 4, columns [32 half, 32 half + 32) of every diagonal =====
        const int quad = warp & 3, half = (warp - 2) >> 2;
        const uint32_t lane_base = (uint32_t)(quad * 32) << 16;
        mbar_wait(acc_full, 0);
        tc_fence_after();
        double t[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) t[j] = 0.0;
#pragma unroll
        for (int g = S - 1; g >= 0; --g) {          // smallest diagonal first
            uint32_t v[32];
            tmem_ld_32x32b_x32(tmem_d + lane_base + (uint32_t)(g * I_N + half * 32), v);
            tmem_wait_ld();
            const double wg = 1.0 / (double)(1ull << (BITS * g));
#pragma unroll
            for (int j = 0; j < 32; ++j) t[j] = fma((double)(int)v[j], wg, t[j]);
        }
        double ssum = 0.0;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            const double x = t[j] * sinv[half * 32 + j];
            ssum = fma(x, x, ssum);
        }
        red[half * I_M + quad * 32 + lane] = ssum;
        named_bar_sync(1, 256);
        if (half == 0) {
            const int m = mb * I_M + quad * 32 + lane;
            qpart[(long)it * Mc + m] = red[quad * 32 + lane] + red[I_M + quad * 32 + lane];
        }
        tc_fence_before();
    }
    __syncthreads();
    if (MC) cluster_sync_all();        // nobody leaves while the peer may still multicast into / arrive on this CTA
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_d), "n"(I_TMEM_COLS) : "memory");
    }
}

// ---- persistent edition (default) ---------------------------------------------------------------------------
// One cluster of two CTAs per SM pair walks the list of (row-tile pair, point tile) work items with a fixed stride (the
// list is band-major, heaviest rows first, so the clusters resident at any time share ~9 L^-1 tile pairs and 8 K* tiles
// through L2).  The TMA ring, its phases and the TMEM allocation live across tiles: the producer runs ahead into the next
// tile while the epilogue warps drain the accumulators, so neither the pipeline ramp (first L2/HBM round trip) nor the
// prologue (barrier init, TMEM allocation, cluster sync) is paid per tile -- at N = 2048 those were ~20 % of a tile.
// Pipelines: smem full/empty (TMA <-> MMA, count 2 on `empty`: both tensor cores of the pair must have retired a stage before
// either producer multicasts into it), TMEM acc_full/acc_empty (MMA <-> 8 epilogue warps).
// Epilogue: diagonals are combined 16 columns at a time (registers: the kernel must leave room for a K* builder CTA of the
// next chunk on the same SM), int32 -> FP64 by the 2^52 trick (one LOP + one DADD instead of I2F.F64).
// Work distribution: sched != null -> items are handed out IN ORDER from an atomic counter (sched[0]); the producer thread of the
// cluster's first CTA fetches the next item and publishes it into an 8-slot ring in BOTH CTAs' shared memory (st.shared::cluster
// + mbarrier.arrive.release.cluster; every role waits on its own CTA's slot barrier).  Ordered hand-out keeps the clusters that are
// resident at any time inside one band of point tiles whatever their items weigh (the static stride let them drift over two
// bands: 24 GB of DRAM reads per launch instead of ~10, profiles/r02_ncu_i8p.txt), and balances the load exactly.  The last
// cluster to finish resets the counters (sched[1] counts finished clusters), so the two ints only need zeroing once.
// sched == null: static stride (item = cluster + t * clusters).
template <int S, int BITS>
__global__ void __launch_bounds__(I_THREADS, 2)      // "2" caps the kernel at 96 registers per thread: 30 K registers, the other half of the file is the builder's
trmm_sumsq_i8p_kernel(const __grid_constant__ CUtensorMap tmAh, const __grid_constant__ CUtensorMap tmB, int Np, int Mc,
                      const double* __restrict__ inv_scale, const double* __restrict__ amax_ptr, double* __restrict__ qpart,
                      const int probe,        // probe != 0 (GPX_I8_PROBE=nofill, timing experiments only): the ring is filled once, never refilled
                      int* __restrict__ sched, const int band_w) {      // band_w: point tiles per band (I_BAND; GPX_I8_BAND overrides)
    using C = I8Cfg<S, 64>;
    constexpr int BKB = 64;
    extern __shared__ unsigned char smem_raw[];
    unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(base + size_t(C::STAGES) * C::STAGE_BYTES);     // full[3] empty[3] acc_full acc_empty
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * C::STAGES + 2);
    double* sinv = reinterpret_cast<double*>(bars + 2 * C::STAGES + 3);      // [64]
    double* red = sinv + I_N;                                                 // [2][128]
    uint64_t* item_bars = reinterpret_cast<uint64_t*>(red + 2 * I_M);         // [I_RING]
    volatile int* item_ring = reinterpret_cast<volatile int*>(item_bars + I_RING);     // [I_RING]
    const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + C::STAGES);
    const uint32_t acc_full = smem_u32(bars + 2 * C::STAGES), acc_empty = smem_u32(bars + 2 * C::STAGES + 1);
    const uint32_t item_full0 = smem_u32(item_bars), item_ring0 = smem_u32(const_cast<int*>(item_ring));
    const uint32_t tile0 = smem_u32(base);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t crank = cluster_ctarank();
    const int n_mb = Mc / I_M, n_it = Np / (2 * I_N);          // n_it counts PAIRS of row tiles
    const int n_items = n_mb * n_it;
    const int cid = (int)(blockIdx.x >> 1), n_clusters = (int)(gridDim.x >> 1);
    const int per_band = band_w * n_it;
    const bool dynamic = sched != nullptr;

    if (tid == 0) {
        for (int s = 0; s < C::STAGES; ++s) {
            mbar_init(full0 + 8 * s, 1);
            mbar_init(empty0 + 8 * s, 2);
        }
        mbar_init(acc_full, 1);
        mbar_init(acc_empty, 8);
        for (int r = 0; r < I_RING; ++r) mbar_init(item_full0 + 8 * r, 1);
        mbar_fence_init();
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(tmem_slot)), "n"(I_TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();        // the peer's barriers must be initialised before anything is multicast at them
    tc_fence_after();
    const uint32_t tmem_d = *tmem_slot;

    // work item id -> (pair of row tiles itp, point tile mb); both CTAs of the cluster take the same items
    // item of iteration t of this cluster (-1: none left); every role of both CTAs sees the same sequence
    auto item_at = [&](uint32_t t) -> int {
        if (!dynamic) {
            const long id = (long)cid + (long)t * n_clusters;
            return id < n_items ? (int)id : -1;
        }
        const uint32_t slot = t % I_RING;
        mbar_wait_cluster(item_full0 + 8 * slot, (t / I_RING) & 1);
        return item_ring[slot];
    };
    // the scheduler (producer thread of the cluster's first CTA) fetches iteration t's item and publishes it to both CTAs
    auto publish_item = [&](uint32_t t) {
        int id = atomicAdd(sched, 1);
        if (id >= n_items) id = -1;
        const uint32_t slot = t % I_RING;
#pragma unroll
        for (uint32_t r = 0; r < 2; ++r) {
            st_shared_cluster_u32(mapa_shared(item_ring0 + 4 * slot, r), (uint32_t)id);
            mbar_arrive_cluster(mapa_shared(item_full0 + 8 * slot, r));
        }
    };
    auto decode = [&](int id, int& itp, int& mb) {
        const int band = id / per_band;
        const int left = n_mb - band * band_w;
        const int width = left < band_w ? left : band_w;
        const int rem = id - band * per_band;
        itp = n_it - 1 - rem / width;
        mb = band * band_w + rem % width;
    };

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            tma_prefetch_desc(&tmAh); tma_prefetch_desc(&tmB);
            uint32_t git = 0;                                   // k iterations issued so far (over all tiles)
            for (uint32_t t = 0;; ++t) {
                if (dynamic && crank == 0) publish_item(t);
                const int id = item_at(t);
                if (id < 0) break;
                int itp, mb;
                decode(id, itp, mb);
                const int nk = (itp + 1) * 2 * I_N / BKB;       // the pair runs the odd tile's extent (zero slices beyond the diagonal)
                const int it = 2 * itp + (int)crank;
                for (int kt = 0; kt < nk; ++kt, ++git) {
                    const uint32_t s = git % C::STAGES, round = git / C::STAGES;
                    if (probe && round > 0) continue;
                    if (round > 0) mbar_wait(empty0 + 8 * s, (round - 1) & 1);
                    mbar_arrive_expect_tx(full0 + 8 * s, C::STAGE_BYTES);
                    const uint32_t dst = tile0 + s * C::STAGE_BYTES;
                    const int k0 = kt * BKB;
#pragma unroll
                    for (int j = 0; j < S; ++j)     // my 64 points of every K* slice tile, delivered to both CTAs of the pair
                        tma_load_2d_multicast(dst + j * C::A_BYTES + crank * (C::A_BYTES / 2), &tmAh, k0,
                                              j * Mc + mb * I_M + (int)crank * (I_M / 2), full0 + 8 * s, 3);
#pragma unroll
                    for (int j = 0; j < S; ++j)
                        tma_load_2d(dst + S * C::A_BYTES + j * C::B_BYTES, &tmB, k0, j * Np + it * I_N, full0 + 8 * s);
                }
            }
            if (dynamic && crank == 0) {            // the last cluster to run out of items rewinds the counters for the next launch
                __threadfence();
                if (atomicAdd(sched + 1, 1) == n_clusters - 1) {
                    sched[0] = 0;
                    sched[1] = 0;
                    __threadfence();
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            uint32_t git = 0;
            for (uint32_t tile_n = 0;; ++tile_n) {
                const int id = item_at(tile_n);
                if (id < 0) break;
                int itp, mb;
                decode(id, itp, mb);
                const int nk = (itp + 1) * 2 * I_N / BKB;
                if (tile_n > 0) {                               // the epilogue warps have drained the previous tile's accumulators
                    mbar_wait(acc_empty, (tile_n - 1) & 1);
                    tc_fence_after();
                }
                for (int kt = 0; kt < nk; ++kt, ++git) {
                    const uint32_t s = git % C::STAGES, round = git / C::STAGES;
                    if (!probe || round == 0) mbar_wait(full0 + 8 * s, round & 1);
                    tc_fence_after();
                    const uint32_t st = tile0 + s * C::STAGE_BYTES;
#pragma unroll
                    for (int kk = 0; kk < BKB / 32; ++kk) {
                        const uint64_t adv = (uint64_t)((kk * 32) >> 4);     // 32 int8 along K inside the swizzle atom
#pragma unroll
                        for (int a = 0; a < S; ++a) {
                            const uint64_t da = umma_desc_kmajor<BKB>(st + a * C::A_BYTES) + adv;
                            const int ncols = (S - a) * I_N;                  // L^-1 slices 0 .. S-1-a -> diagonals a .. S-1
#pragma unroll
                            for (int c0 = 0; c0 < ncols; c0 += 256) {
                                const int w = (ncols - c0) < 256 ? (ncols - c0) : 256;
                                const uint64_t db = umma_desc_kmajor<BKB>(st + S * C::A_BYTES + (c0 / I_N) * C::B_BYTES) + adv;
                                // the first k step of K* slice 0 touches every accumulator once: it overwrites (no zero fill)
                                umma_i8(tmem_d + (uint32_t)(a * I_N + c0), da, db, i8_idesc(w), (kt | kk | a) != 0 ? 1u : 0u);
                            }
                        }
                    }
                    if (!probe) umma_commit_multicast(empty0 + 8 * s, 3);     // both producers wait for both tensor cores
                }
                umma_commit(acc_full);
            }
        }
    } else {
        // ===== epilogue warps 2..9: TMEM lane quadrant = warp % 4, columns [32 half, 32 half + 32) of every diagonal =====
        const int quad = warp & 3, half = (warp - 2) >> 2;
        const int et = tid - 64;                                 // 0..255 among the epilogue threads
        const uint32_t lane_base = (uint32_t)(quad * 32) << 16;
        const int ek = scale_exponent_b<BITS>(*amax_ptr);
        const double kfac = ldexp(1.0, ek + 1 - 2 * BITS);       // V = (sum_g acc_g 2^{-BITS g}) * 2^{-2 BITS} / (scale_i scale_K): exact
        for (uint32_t tile_n = 0;; ++tile_n) {
            const int id = item_at(tile_n);
            if (id < 0) break;
            int itp, mb;
            decode(id, itp, mb);
            const int it = 2 * itp + (int)crank;
            if (et < I_N) sinv[et] = inv_scale[it * I_N + et] * kfac;
            mbar_wait(acc_full, tile_n & 1);
            tc_fence_after();
            double ssum = 0.0;
#pragma unroll
            for (int cg = 0; cg < 2; ++cg) {
                double t[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) t[j] = 0.0;
#pragma unroll
                for (int g = S - 1; g >= 0; --g) {          // smallest diagonal first
                    uint32_t v[16];
                    tmem_ld_32x32b_x16(tmem_d + lane_base + (uint32_t)(g * I_N + half * 32 + cg * 16), v);
                    tmem_wait_ld();
                    const double wg = 1.0 / (double)(1ull << (BITS * g));
#pragma unroll
                    for (int j = 0; j < 16; ++j)            // (double)(int)v: 2^52 + 2^31 + (v xor 2^31) as a bit pattern, minus the offset
                        t[j] = fma(__hiloint2double(0x43300000, (int)(v[j] ^ 0x80000000u)) - 4503601774854144.0, wg, t[j]);
                }
                if (cg == 1) {                               // every accumulator column of this warp has been read
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(acc_empty);
                }
                if (cg == 0) named_bar_sync(1, 256);         // sinv of this tile is in place
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const double x = t[j] * sinv[half * 32 + cg * 16 + j];
                    ssum = fma(x, x, ssum);
                }
            }
            red[half * I_M + quad * 32 + lane] = ssum;
            named_bar_sync(1, 256);
            if (half == 0) {
                const int m = mb * I_M + quad * 32 + lane;
                qpart[(long)it * Mc + m] = red[quad * 32 + lane] + red[I_M + quad * 32 + lane];
            }
            // (no third barrier: sinv is last read before the barrier above, red before the next tile's first one)
        }
    }
    __syncthreads();
    cluster_sync_all();        // nobody leaves while the peer may still multicast into / arrive on this CTA
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_d), "n"(I_TMEM_COLS) : "memory");
    }
}

// ---- host -------------------------------------------------------------------------------------------------
static int encode_u8_rowmajor(CUtensorMap* tm, const int8_t* base, uint64_t rows, uint64_t cols, uint32_t box_cols,
                              uint32_t box_rows) {
    PFN_encodeTiled enc = get_encode_fn();
    if (!enc) return -900;
    cuuint64_t gdim[2] = {cols, rows};
    cuuint64_t gstride[1] = {cols};
    cuuint32_t box[2] = {box_cols, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<int8_t*>(base), gdim, gstride, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, box_cols == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : -901 - (int)r;
}

int i8_row_tiles(int Np) { return Np / I_N; }

// planes <- slices of src [rows][cols] (ld_src); tri: mirrored inverse with per-row scales -> inv_scale[rows]
int slice_i8_matrix(const double* src, int rows, int cols, long ld_src, int8_t* planes, long plane_rows, int nslices, int bits,
                    int tri, double* inv_scale, const double* amax_ptr, cudaStream_t st) {
    if (cols % TILE || nslices < 5 || nslices > I_SMAX || (bits != 7 && bits != 8)) return -2;
    if (rows <= 0) return 0;
#define GPX_SLICE(S_, B_) slice_i8_kernel<S_, B_><<<rows, 256, 0, st>>>(src, cols, ld_src, planes, plane_rows, tri, inv_scale, amax_ptr)
    if (nslices == 5 && bits == 7) GPX_SLICE(5, 7);
    else if (nslices == 6 && bits == 7) GPX_SLICE(6, 7);
    else if (nslices == 5) GPX_SLICE(5, 8);
    else GPX_SLICE(6, 8);
#undef GPX_SLICE
    GPX_CHECK_LAUNCH();
    return 0;
}

// GPX_I8_IMPL: unset = persistent clusters (trmm_sumsq_i8p_kernel); "tile" = one work item per cluster (the round-1 edition);
// "1cta" = no pairing at all
static int i8_impl() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("GPX_I8_IMPL");
        v = (e && e[0] == '1') ? 1 : (e && e[0] == 't') ? 2 : 0;
    }
    return v;
}

static void cluster2_config(cudaLaunchConfig_t& cfg, cudaLaunchAttribute* at, int grid, size_t smem, cudaStream_t st) {
    cfg = cudaLaunchConfig_t{};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(I_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
}

static int g_i8_clusters = 0;
int i8_clusters() { return g_i8_clusters; }     // clusters of the last configured persistent kernel (diagnostics)

template <int S, int BITS>
static int launch_i8(const int8_t* Ls, const double* inv_scale, int Np, const int8_t* Kp, int Mc, const double* amax_ptr,
                     double* qpart, int* sched, cudaStream_t st) {
    using C = I8Cfg<S, 64>;
    static bool attr[64] = {false};
    static int max_clusters[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) return -30;
    const int impl = i8_impl();
    cudaLaunchConfig_t cfg;
    cudaLaunchAttribute at[1];
    if (!attr[dev]) {
        cudaFuncSetAttribute(trmm_sumsq_i8p_kernel<S, BITS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM_BYTES);
        // the whole 228 KB as shared memory: the SM's carve-out must also hold the K* builder CTA that runs beside this kernel
        cudaFuncSetAttribute(trmm_sumsq_i8p_kernel<S, BITS>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        cudaFuncSetAttribute(trmm_sumsq_i8_kernel<S, 64, true, BITS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM_BYTES);
        cudaFuncSetAttribute(trmm_sumsq_i8_kernel<S, 64, false, BITS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM_BYTES);
        cluster2_config(cfg, at, 2, C::SMEM_BYTES, st);
        int n = 0;
        if (cudaOccupancyMaxActiveClusters(&n, trmm_sumsq_i8p_kernel<S, BITS>, &cfg) != cudaSuccess || n < 1) {
            cudaGetLastError();
            cudaDeviceProp prop;
            cudaGetDeviceProperties(&prop, dev);
            n = prop.multiProcessorCount / 2;
        }
        max_clusters[dev] = n;
        g_i8_clusters = n;
        attr[dev] = true;
    }
    CUtensorMap tmA, tmAh, tmB;
    int rc;
    if ((rc = encode_u8_rowmajor(&tmA, Kp, (uint64_t)S * Mc, Np, 64, I_M))) return rc;
    if ((rc = encode_u8_rowmajor(&tmAh, Kp, (uint64_t)S * Mc, Np, 64, I_M / 2))) return rc;
    if ((rc = encode_u8_rowmajor(&tmB, Ls, (uint64_t)S * Np, Np, 64, I_N))) return rc;
    const int tiles = (Np / I_N) * (Mc / I_M);
    if (impl == 1) {
        trmm_sumsq_i8_kernel<S, 64, false, BITS><<<tiles, I_THREADS, C::SMEM_BYTES, st>>>(tmA, tmB, Np, Mc, tmAh, inv_scale, amax_ptr, qpart);
    } else if (impl == 2) {
        cluster2_config(cfg, at, tiles, C::SMEM_BYTES, st);
        cudaError_t e = cudaLaunchKernelEx(&cfg, trmm_sumsq_i8_kernel<S, 64, true, BITS>, tmA, tmB, Np, Mc, tmAh, inv_scale, amax_ptr, qpart);
        if (e != cudaSuccess) return -1000 - (int)e;
    } else {
        const int items = tiles / 2;
        const int clusters = items < max_clusters[dev] ? items : max_clusters[dev];
        cluster2_config(cfg, at, 2 * clusters, C::SMEM_BYTES, st);
        static int probe = -1;
        if (probe < 0) { const char* pe = getenv("GPX_I8_PROBE"); probe = (pe && pe[0] == 'n') ? 1 : 0; }
        static int sched_off = -1;      // GPX_I8_SCHED=static: fixed stride instead of the ordered atomic hand-out
        if (sched_off < 0) { const char* se = getenv("GPX_I8_SCHED"); sched_off = (se && se[0] == 's') ? 1 : 0; }
        static int band_w = -1;
        if (band_w < 0) { const char* be = getenv("GPX_I8_BAND"); band_w = (be && atoi(be) > 0) ? atoi(be) : I_BAND; }
        cudaError_t e = cudaLaunchKernelEx(&cfg, trmm_sumsq_i8p_kernel<S, BITS>, tmAh, tmB, Np, Mc, inv_scale, amax_ptr, qpart, probe,
                                           sched_off ? (int*)nullptr : sched, band_w);
        if (e != cudaSuccess) return -1000 - (int)e;
    }
    GPX_CHECK_LAUNCH();
    return 0;
}

// largest contraction length whose int32 sums cannot overflow: |digit| <= 64 (7 bit) / 128 (8 bit), S products per diagonal
// Largest padded train size the kernels accept / the largest one whose int32 sums are exact whatever the digits are:
//   |sum| <= (S digit products per accumulator) * 128 * 128 * K < 2^31  <=>  K <= 16384 for 8-bit digits (S <= 6),
//   |digit| <= 64 keeps 7-bit digits exact to K = 32768.
// Between the two, 8-bit digits are exact when the ACTUAL digits of L^-1 allow it: every accumulator is bounded by
//   128 * max_i sum_{planes b} sum_k |digit_b(L^-1[i][k])|      (|K* digit| <= 128),
// which i8_digit_bound() evaluates on the sliced planes (typical: 2.5 x below 2^31 at Np = 26240, five digits).
int i8_max_np(int bits) { return (bits == 7 || bits == 8) ? 32768 : 0; }
int i8_safe_np(int bits) { return bits == 8 ? 16384 : (bits == 7 ? 32768 : 0); }

// out[0] = max over rows of  sum_{plane, k} |digit|   (one block per row of the S planes [Np][Np]); out must be zero on entry.
// Only the columns up to the end of the row's own 128-block count: that is the k range the tensor-core kernel reads for the row
// (gpx_slice_i8 with tri = 1 leaves everything right of it untouched).
__global__ void __launch_bounds__(256) i8_digit_bound_kernel(const int8_t* __restrict__ planes, int S, int Np,
                                                             unsigned long long* __restrict__ out) {
    const int row = blockIdx.x;
    const int kend = (row / TILE + 1) * TILE;
    unsigned int acc = 0;
    for (int j = 0; j < S; ++j) {
        const int4* src = reinterpret_cast<const int4*>(planes + ((long)j * Np + row) * Np);
        for (int c = threadIdx.x; c < kend / 16; c += 256) {
            const int4 v = src[c];
            const unsigned int w[4] = {(unsigned int)v.x, (unsigned int)v.y, (unsigned int)v.z, (unsigned int)v.w};
#pragma unroll
            for (int q = 0; q < 4; ++q) acc = __dp4a(__vabs4(w[q]), 0x01010101u, acc);      // |-128| -> 128 (unsigned bytes)
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    __shared__ unsigned int red[8];
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long t = 0;
        for (int w = 0; w < 8; ++w) t += red[w];
        atomicMax(out, t);
    }
}

int i8_digit_bound(const int8_t* planes, int nslices, int Np, unsigned long long* out, cudaStream_t st) {
    if (Np % TILE || nslices < 1 || nslices > I_SMAX) return -4;
    cudaMemsetAsync(out, 0, sizeof(unsigned long long), st);
    i8_digit_bound_kernel<<<Np, 256, 0, st>>>(planes, nslices, Np, out);
    GPX_CHECK_LAUNCH();
    return 0;
}

// qpart[it][m], it < Np/64.  Ls: S planes [Np][Np] of the sliced inverse, Kp: S planes [Mc][Np] of the sliced K* chunk.
// sched: two device ints, zero before the first launch that uses them (the kernel rewinds them itself), not shared by launches that
// can run at the same time; null = static work distribution.
int trmm_sumsq_i8(const int8_t* Ls, const double* inv_scale, int Np, const int8_t* Kp, int Mc, int nslices, int bits,
                  const double* amax_ptr, double* qpart, int* sched, cudaStream_t st) {
    if (Mc % I_M || Np % TILE) return -6;
    if (bits != 7 && bits != 8) return -7;
    if (Np > i8_max_np(bits)) return -7;              // int32 accumulators
    if (nslices == 5) return bits == 7 ? launch_i8<5, 7>(Ls, inv_scale, Np, Kp, Mc, amax_ptr, qpart, sched, st)
                                       : launch_i8<5, 8>(Ls, inv_scale, Np, Kp, Mc, amax_ptr, qpart, sched, st);
    if (nslices == 6) return bits == 7 ? launch_i8<6, 7>(Ls, inv_scale, Np, Kp, Mc, amax_ptr, qpart, sched, st)
                                       : launch_i8<6, 8>(Ls, inv_scale, Np, Kp, Mc, amax_ptr, qpart, sched, st);
    return -8;
}

// ---- general sliced-integer NT product with a store epilogue (the merge steps of the recursive triangular inverse) ----
//   C[c_row0 + i][c_col0 + j] = sign * sum_{k in range(i)} A[a_row0 + i][k] * B[b_row0 + j][k]      (+ the same value at C2[j][i])
// A (MMA M side, 128-row tiles mi) and B (N side, 64-row tiles nj) are int8 slice planes [S][Np][Np] with per-row scales;
// the k range of tile row mi is [k_base + 128 mi, k_end) (k_from = 1: operand rows that start at their own diagonal block)
// or [k_base, k_base + 128 (mi + 1)) (k_from = 0: rows that end there).  Same pipeline as trmm_sumsq_i8_kernel<S, 64, true>:
// clusters of two CTAs on neighbouring N tiles share the M-side tile through TMA multicast.
struct I8Gemm {
    int n_nj;                   // N tiles (even)
    int a_row0, b_row0;
    int k_base, k_end, k_from;
    int c_row0, c_col0;
    int m_row0, m_col0;         // mirror C2[m_row0 + j][m_col0 + i]; m_row0 < 0: none
    int tri;                    // 1: only the N tiles of the first mi + 1 pairs exist (tiles on or below the block diagonal)
    double sign;
    int accumulate;             // 1: C += sign * product (the Cholesky trailing update), 0: C = sign * product
    int n_mi;                   // M tiles (set by launch_gemm_i8)
    int ldc;                    // leading dimension of C (0: Np); the mirror always has Np
};
constexpr int I8G_BAND = 12;    // M tiles per band of the tile order: the ~74 resident clusters then cover 12 M tiles x 12 N tiles
                                // (24 operand tiles per k step instead of 1 + 148), so each k step's operands are read from
                                // DRAM once and served to the other CTAs from L2

template <int S, int BITS>
__global__ void __launch_bounds__(I_THREADS, 1)
gemm_i8_kernel(const __grid_constant__ CUtensorMap tmAh, const __grid_constant__ CUtensorMap tmB, int Np,
               const double* __restrict__ sa, const double* __restrict__ sb, double* __restrict__ Cout,
               double* __restrict__ Cmir, const I8Gemm g) {
    using C = I8Cfg<S, 64>;
    constexpr int BKB = 64;
    extern __shared__ unsigned char smem_raw[];
    unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(base + size_t(C::STAGES) * C::STAGE_BYTES);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * C::STAGES + 1);
    double* sinv = reinterpret_cast<double*>(bars + 2 * C::STAGES + 2);      // [64]
    const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + C::STAGES), acc_full = smem_u32(bars + 2 * C::STAGES);
    const uint32_t tile0 = smem_u32(base);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t crank = cluster_ctarank();
    const int pairs = g.n_nj / 2;
    const int cid = (int)(blockIdx.x >> 1);
    // band-major tile order: inside a band of I8G_BAND M tiles the M index runs fastest
    const int band = cid / (I8G_BAND * pairs), within = cid - band * (I8G_BAND * pairs);
    const int gm = (g.n_mi - band * I8G_BAND) < I8G_BAND ? (g.n_mi - band * I8G_BAND) : I8G_BAND;
    const int pr = within / gm;
    const int mi = band * I8G_BAND + (within - pr * gm), nj = 2 * pr + (int)crank;
    if (g.tri && (nj >> 1) > mi) return;          // both CTAs of the pair leave together, before any barrier exists
    const int kstart = g.k_from == 1 ? g.k_base + mi * I_M : g.k_base;            // k_from = 2: the same range [k_base, k_end) for every tile
    const int kend = g.k_from ? g.k_end : g.k_base + (mi + 1) * I_M;
    const int nk = (kend - kstart) / BKB;
    const int arow = g.a_row0 + mi * I_M, brow = g.b_row0 + nj * I_N;

    if (tid == 0) {
        for (int s = 0; s < C::STAGES; ++s) {
            mbar_init(full0 + 8 * s, 1);
            mbar_init(empty0 + 8 * s, 2);
        }
        mbar_init(acc_full, 1);
        mbar_fence_init();
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(tmem_slot)), "n"(I_TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    if (tid < I_N) sinv[tid] = sb[brow + tid] * ldexp(1.0, -2 * BITS);
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_d = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            tma_prefetch_desc(&tmAh); tma_prefetch_desc(&tmB);
            for (int kt = 0; kt < nk; ++kt) {
                const int s = kt % C::STAGES, round = kt / C::STAGES;
                if (round > 0) mbar_wait(empty0 + 8 * s, (round - 1) & 1);
                mbar_arrive_expect_tx(full0 + 8 * s, C::STAGE_BYTES);
                const uint32_t dst = tile0 + s * C::STAGE_BYTES;
                const int k0 = kstart + kt * BKB;
#pragma unroll
                for (int j = 0; j < S; ++j)
                    tma_load_2d_multicast(dst + j * C::A_BYTES + crank * (C::A_BYTES / 2), &tmAh, k0,
                                          j * Np + arow + (int)crank * (I_M / 2), full0 + 8 * s, 3);
#pragma unroll
                for (int j = 0; j < S; ++j)
                    tma_load_2d(dst + S * C::A_BYTES + j * C::B_BYTES, &tmB, k0, j * Np + brow, full0 + 8 * s);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            for (int kt = 0; kt < nk; ++kt) {
                const int s = kt % C::STAGES, round = kt / C::STAGES;
                mbar_wait(full0 + 8 * s, round & 1);
                tc_fence_after();
                const uint32_t st = tile0 + s * C::STAGE_BYTES;
#pragma unroll
                for (int kk = 0; kk < BKB / 32; ++kk) {
                    const uint64_t adv = (uint64_t)((kk * 32) >> 4);
#pragma unroll
                    for (int a = 0; a < S; ++a) {
                        const uint64_t da = umma_desc_kmajor<BKB>(st + a * C::A_BYTES) + adv;
                        const int ncols = (S - a) * I_N;
#pragma unroll
                        for (int c0 = 0; c0 < ncols; c0 += 256) {
                            const int w = (ncols - c0) < 256 ? (ncols - c0) : 256;
                            const uint64_t db = umma_desc_kmajor<BKB>(st + S * C::A_BYTES + (c0 / I_N) * C::B_BYTES) + adv;
                            umma_i8(tmem_d + (uint32_t)(a * I_N + c0), da, db, i8_idesc(w), (kt | kk | a) != 0 ? 1u : 0u);
                        }
                    }
                }
                umma_commit_multicast(empty0 + 8 * s, 3);
            }
            umma_commit(acc_full);
        }
    } else {
        const int quad = warp & 3, half = (warp - 2) >> 2;
        const uint32_t lane_base = (uint32_t)(quad * 32) << 16;
        const int r = quad * 32 + lane;
        double* dst = Cout + (long)(g.c_row0 + mi * I_M + r) * (g.ldc ? g.ldc : Np) + g.c_col0 + nj * I_N + half * 32;
        double c_old[32];               // accumulate mode: the old values of C are fetched while the MMAs run
        if (g.accumulate) {
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
                const double2 c = *reinterpret_cast<const double2*>(dst + j);
                c_old[j] = c.x;
                c_old[j + 1] = c.y;
            }
        }
        mbar_wait(acc_full, 0);
        tc_fence_after();
        double t[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) t[j] = 0.0;
#pragma unroll
        for (int gi = S - 1; gi >= 0; --gi) {
            uint32_t v[32];
            tmem_ld_32x32b_x32(tmem_d + lane_base + (uint32_t)(gi * I_N + half * 32), v);
            tmem_wait_ld();
            const double wg = 1.0 / (double)(1ull << (BITS * gi));
#pragma unroll
            for (int j = 0; j < 32; ++j)        // (double)(int)v by the 2^52 trick: one LOP + one DADD instead of I2F.F64 on the XU pipe
                t[j] = fma(__hiloint2double(0x43300000, (int)(v[j] ^ 0x80000000u)) - 4503601774854144.0, wg, t[j]);
        }
        const double si = sa[arow + r] * g.sign;
        if (g.accumulate) {
#pragma unroll
            for (int j = 0; j < 32; ++j) t[j] = fma(t[j], sinv[half * 32 + j] * si, c_old[j]);
        } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) t[j] *= sinv[half * 32 + j] * si;
        }
#pragma unroll
        for (int j = 0; j < 32; j += 2) *reinterpret_cast<double2*>(dst + j) = make_double2(t[j], t[j + 1]);
        if (g.m_row0 >= 0) {       // transposed copy: consecutive lanes write consecutive doubles of one mirror row
            double* mir = Cmir + (long)(g.m_row0 + nj * I_N + half * 32) * Np + g.m_col0 + mi * I_M + r;
#pragma unroll
            for (int j = 0; j < 32; ++j) mir[(long)j * Np] = t[j];
        }
        tc_fence_before();
    }
    __syncthreads();
    cluster_sync_all();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_d), "n"(I_TMEM_COLS) : "memory");
    }
}

// Rows [r0, r0 + gridDim.x) of a row-major FP64 matrix -> int8 slices over the live columns [lo(r), hi(r)):
//   lo(r) = lo_row ? r : c0,  hi(r) = hi_row ? r + 1 : c1;  the written range is the enclosing 128-aligned one, zero outside
//   the live part.  DT != null: inside the row's own 128-block the value comes from DT[block][r % 128][c % 128].
template <int S, int BITS>
__global__ void __launch_bounds__(256) slice_i8_rows_kernel(const double* __restrict__ src, long ld, const double* __restrict__ DT,
                                                            int r0, int c0, int c1, int lo_row, int hi_row, int Np,
                                                            int8_t* __restrict__ planes, double* __restrict__ inv_scale) {
    const int row = r0 + blockIdx.x, b = row / TILE;
    const double* xs = src + (long)row * ld;
    const double* xd = DT ? DT + (long)b * TILE * TILE + (long)(row - b * TILE) * TILE - (long)b * TILE : nullptr;
    const int lo = lo_row ? row : c0, hi = hi_row ? row + 1 : c1;
    const int wlo = lo_row ? b * TILE : c0, whi = hi_row ? (b + 1) * TILE : c1;
    const int blo = b * TILE, bhi = (b + 1) * TILE;
    __shared__ double red[8];
    __shared__ int s_e;
    double a = 0.0;
    for (int c = lo + threadIdx.x; c < hi; c += 256) a = fmax(a, fabs((xd && c >= blo && c < bhi) ? xd[c] : xs[c]));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a = fmax(a, __shfl_xor_sync(0xffffffffu, a, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = a;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) a = fmax(a, red[w]);
        s_e = scale_exponent_b<BITS>(a);
        inv_scale[row] = ldexp(1.0, s_e + 1);
    }
    __syncthreads();
    const double scale = ldexp(1.0, -(s_e + 1) + (BITS == 7 ? I_BITS : 0));
    const long pitch = (long)Np * Np;
    int8_t* out = planes + (long)row * Np;
    for (int cc = wlo + threadIdx.x * 8; cc < whi; cc += 256 * 8) {
        double r[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int c = cc + u;
            r[u] = (c >= lo && c < hi) ? ((xd && c >= blo && c < bhi) ? xd[c] : xs[c]) * scale : 0.0;
        }
        unsigned long long dg[8];
        if (BITS == 8) {
#pragma unroll
            for (int u = 0; u < 8; ++u) dg[u] = balanced_digits<S>(r[u]);
        }
#pragma unroll
        for (int j = 0; j < S; ++j) {
            uint32_t l32 = 0, h32 = 0;
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                uint32_t bb;
                if (BITS == 7) {
                    const double q = rint(r[u]);
                    r[u] = (r[u] - q) * 128.0;
                    bb = (uint32_t)(__double2int_rn(q) & 0xff);
                } else {
                    bb = (uint32_t)(dg[u] >> (8 * (S - 1 - j))) & 0xffu;
                }
                if (u < 4) l32 |= bb << (8 * u); else h32 |= bb << (8 * (u - 4));
            }
            *reinterpret_cast<uint2*>(out + j * pitch + cc) = make_uint2(l32, h32);
        }
    }
}

size_t trtri_i8_workspace_bytes(int Np, int nslices) { return 2 * ((size_t)nslices * Np * Np + (size_t)Np * 8); }

template <int S, int BITS>
static int launch_gemm_i8(const int8_t* PA, const int8_t* PB, int Np, const double* sa, const double* sb, double* Cout, double* Cmir,
                          const I8Gemm& g, int n_mi, cudaStream_t st) {
    using C = I8Cfg<S, 64>;
    static bool attr[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!(dev >= 0 && dev < 64 && attr[dev])) {
        cudaFuncSetAttribute(gemm_i8_kernel<S, BITS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM_BYTES);
        if (dev >= 0 && dev < 64) attr[dev] = true;
    }
    CUtensorMap tmAh, tmB;
    int rc;
    if ((rc = encode_u8_rowmajor(&tmAh, PA, (uint64_t)S * Np, Np, 64, I_M / 2))) return rc;
    if ((rc = encode_u8_rowmajor(&tmB, PB, (uint64_t)S * Np, Np, 64, I_N))) return rc;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(n_mi * g.n_nj);
    cfg.blockDim = dim3(I_THREADS);
    cfg.dynamicSmemBytes = C::SMEM_BYTES;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    I8Gemm gg = g;
    gg.n_mi = n_mi;
    cudaError_t e = cudaLaunchKernelEx(&cfg, gemm_i8_kernel<S, BITS>, tmAh, tmB, Np, sa, sb, Cout, Cmir, gg);
    return e == cudaSuccess ? 0 : -1000 - (int)e;
}

// ---- K_hat^-1 = L^-T L^-1 on the integer tensor cores (lauum for the gradient) ----------------------------------------
// Row i of the operand = column i of L^-1 over k >= i: the transposed diagonal block DT[b] inside the row's own 128-block,
// the mirrored L^-T blocks of S right of it, zero left of the diagonal (slice_i8_rows_kernel with lo = row, DT source).
// Both operands of the product are tiles of that one sliced matrix; the k range of tile row mi starts at its 128-block and
// only the tiles on or below the block diagonal are formed (what gpx_mll_grad reads).
size_t lauum_i8_workspace_bytes(int Np, int nslices) { return (size_t)nslices * Np * Np + (size_t)Np * 8; }

int lauum_i8(const double* Smat, const double* DT, double* W, int Np, int nslices, void* workspace, size_t workspace_bytes,
             cudaStream_t st) {
    if (Np % TILE || Np > 32768) return -4;
    if (nslices < 5 || nslices > I_SMAX) return -5;
    if (workspace_bytes < lauum_i8_workspace_bytes(Np, nslices)) return -7;
    double* inv_scale = reinterpret_cast<double*>(workspace);
    int8_t* planes = reinterpret_cast<int8_t*>(inv_scale + Np);
    const int nbt = Np / I_M;
    const I8Gemm g = {2 * nbt, 0, 0, 0, Np, 1, 0, 0, -1, 0, 1, 1.0, 0, 0, 0};
    if (nslices == 5) {
        slice_i8_rows_kernel<5, 7><<<Np, 256, 0, st>>>(Smat, Np, DT, 0, 0, Np, 1, 0, Np, planes, inv_scale);
        return launch_gemm_i8<5, 7>(planes, planes, Np, inv_scale, inv_scale, W, nullptr, g, nbt, st);
    }
    slice_i8_rows_kernel<6, 7><<<Np, 256, 0, st>>>(Smat, Np, DT, 0, 0, Np, 1, 0, Np, planes, inv_scale);
    return launch_gemm_i8<6, 7>(planes, planes, Np, inv_scale, inv_scale, W, nullptr, g, nbt, st);
}

// One merge level of the recursive inverse (gpx_chol.cu: trtri) for the pair of segments [left, left+h) | [right, right+hr) (in
// 128-blocks) on the integer tensor cores:  T = L[right,left] A^-1 (stored transposed in W), X21 = -C^-1 T -> S (+ mirror).
template <int BITS>
static int trtri_i8_merge_t(const double* L, double* Smat, const double* DT, double* W, int Np, int left, int h, int hr, void* workspace,
                            cudaStream_t st) {
    constexpr int S = 6;
    const size_t plane_bytes = (size_t)S * Np * Np;
    double* sA = reinterpret_cast<double*>(workspace);
    double* sB = sA + Np;
    int8_t* PA = reinterpret_cast<int8_t*>(sB + Np);
    int8_t* PB = PA + plane_bytes;
    const int right = left + h;
    const int l0 = left * TILE, r0 = right * TILE, r1 = (right + hr) * TILE;
    // step 1: M side = columns of A^-1 as rows (left segment, k from the row to the end of the segment), N side = L[right, left]
    slice_i8_rows_kernel<S, BITS><<<h * TILE, 256, 0, st>>>(Smat, Np, DT, l0, 0, r0, 1, 0, Np, PA, sA);
    slice_i8_rows_kernel<S, BITS><<<hr * TILE, 256, 0, st>>>(L, Np, nullptr, r0, l0, r0, 0, 0, Np, PB, sB);
    I8Gemm g1 = {2 * hr, l0, r0, l0, r0, 1, l0, r0, -1, 0, 0, 1.0, 0, 0, 0};
    int rc = launch_gemm_i8<S, BITS>(PA, PB, Np, sA, sB, W, nullptr, g1, h, st);
    if (rc) return rc;
    // step 2: M side = rows of C^-1 (right segment, k from the segment start to the row), N side = T^T = W[left, right]
    slice_i8_rows_kernel<S, BITS><<<hr * TILE, 256, 0, st>>>(Smat, Np, nullptr, r0, r0, 0, 0, 1, Np, PA, sA);
    slice_i8_rows_kernel<S, BITS><<<h * TILE, 256, 0, st>>>(W, Np, nullptr, l0, r0, r1, 0, 0, Np, PB, sB);
    I8Gemm g2 = {2 * h, r0, l0, r0, r1, 0, r0, l0, l0, r0, 0, -1.0, 0, 0, 0};
    rc = launch_gemm_i8<S, BITS>(PA, PB, Np, sA, sB, Smat, Smat, g2, hr, st);
    if (rc) return rc;
    GPX_CHECK_LAUNCH();
    return 0;
}

// ---- Cholesky bulk trailing update on the integer tensor cores (training mode; gpx_chol.cu: potrf) ----------------------
//   K[i][j] -= sum_{k in [c0, c1)} K[i][k] K[j][k]      for the 128-tiles on or below the block diagonal of rows/columns >= r0
// The rows of the finished panels [c0, c1) are sliced into six 8-bit digits (48 bits, one power-of-two scale per row; the sums
// run over c1 - c0 <= 1024 entries, far inside the int32 range at any Np) and multiplied with themselves.
// Two plane buffers: the critical-path stream slices the panels of outer block q + 1 while the helper stream's bulk update of
// outer block q still reads the other buffer.
size_t potrf_i8_workspace_bytes(int Np) { return 2 * ((size_t)6 * Np * Np + (size_t)Np * 8); }

static inline double* potrf_i8_scales(void* workspace, int Np, int buf) {
    return reinterpret_cast<double*>(reinterpret_cast<unsigned char*>(workspace) + (size_t)buf * ((size_t)6 * Np * Np + (size_t)Np * 8));
}

// digits of K[r0.., c0..c1) (the finished panels of one outer block, every row below them) into plane buffer `buf`
int potrf_i8_slice(const double* K, int Np, int c0, int c1, int r0, void* workspace, size_t workspace_bytes, int buf, cudaStream_t st) {
    constexpr int S = 6, BITS = 8;
    if (Np % TILE || r0 % TILE || c0 % TILE || c1 % TILE || c1 > r0 || r0 >= Np) return -4;
    if (workspace_bytes < potrf_i8_workspace_bytes(Np) || (buf & ~1)) return -7;
    double* sA = potrf_i8_scales(workspace, Np, buf);
    int8_t* PA = reinterpret_cast<int8_t*>(sA + Np);
    slice_i8_rows_kernel<S, BITS><<<Np - r0, 256, 0, st>>>(K, Np, nullptr, r0, c0, c1, 0, 0, Np, PA, sA);
    GPX_CHECK_LAUNCH();
    return 0;
}

// K[i][j] -= sum_k ...  for the 128-tiles (i, j), i >= j, of block rows >= r0 / 128 and block columns [r0 / 128, r0 / 128 + ncol_blocks)
// (ncol_blocks < 0: every block column down to the end -- the whole trailing triangle), from plane buffer `buf`.
int potrf_i8_update(double* K, int Np, int c0, int c1, int r0, int ncol_blocks, void* workspace, int buf, cudaStream_t st) {
    constexpr int S = 6, BITS = 8;
    double* sA = potrf_i8_scales(workspace, Np, buf);
    int8_t* PA = reinterpret_cast<int8_t*>(sA + Np);
    const int nbt = (Np - r0) / I_M;
    const int ncb = (ncol_blocks < 0 || ncol_blocks > nbt) ? nbt : ncol_blocks;
    const I8Gemm g = {2 * ncb, r0, r0, c0, c1, 2, r0, r0, -1, 0, 1, -1.0, 1, 0, 0};
    const int rc = launch_gemm_i8<S, BITS>(PA, PA, Np, sA, sA, K, nullptr, g, nbt, st);
    if (rc) return rc;
    GPX_CHECK_LAUNCH();
    return 0;
}

// ---- dense predictive covariance on the integer tensor cores (training loops: the validation loss of every epoch) ----------
//   Vt[m][i] = sum_{k <= i} Ks[m][k] L^-1[i][k]          (Mp x Np, the layout gpx_trmm_store writes)
//   C[m][m'] -= sum_i Vt[m][i] Vt[m'][i]                 (C = K** on entry; only the 128-tiles on or below the diagonal)
// Step 1 puts the rows of L^-1 on the M side (k range of tile row mi ends with its diagonal block) and the test points on the
// N side; V itself goes to the scratch matrix W, its transpose -- what step 2 slices -- to Vt through the mirror store.
template <int BITS>
static int predict_covar_i8_t(const double* Smat, int Np, const double* Ks, int Mp, double* Vt, double* C, double* W, void* workspace,
                              cudaStream_t st) {
    constexpr int S = 6;
    const size_t plane_bytes = (size_t)S * Np * Np;
    double* sA = reinterpret_cast<double*>(workspace);
    double* sB = sA + Np;
    int8_t* PA = reinterpret_cast<int8_t*>(sB + Np);
    int8_t* PB = PA + plane_bytes;
    slice_i8_rows_kernel<S, BITS><<<Np, 256, 0, st>>>(Smat, Np, nullptr, 0, 0, 0, 0, 1, Np, PA, sA);
    slice_i8_rows_kernel<S, BITS><<<Mp, 256, 0, st>>>(Ks, Np, nullptr, 0, 0, Np, 0, 0, Np, PB, sB);
    const I8Gemm g1 = {2 * (Mp / I_M), 0, 0, 0, Np, 0, 0, 0, 0, 0, 0, 1.0, 0, 0, 0};
    int rc = launch_gemm_i8<S, BITS>(PA, PB, Np, sA, sB, W, Vt, g1, Np / I_M, st);
    if (rc) return rc;
    slice_i8_rows_kernel<S, BITS><<<Mp, 256, 0, st>>>(Vt, Np, nullptr, 0, 0, Np, 0, 0, Np, PA, sA);
    const I8Gemm g2 = {2 * (Mp / I_M), 0, 0, 0, Np, 2, 0, 0, -1, 0, 1, -1.0, 1, 0, Mp};
    rc = launch_gemm_i8<S, BITS>(PA, PA, Np, sA, sA, C, nullptr, g2, Mp / I_M, st);
    if (rc) return rc;
    GPX_CHECK_LAUNCH();
    return 0;
}

int predict_covar_i8(const double* Smat, int Np, const double* Ks, int Mp, double* Vt, double* C, double* W, int bits, void* workspace,
                     size_t workspace_bytes, cudaStream_t st) {
    if (Np % TILE || Mp % TILE || Mp > Np || Mp <= 0) return -4;
    if (workspace_bytes < trtri_i8_workspace_bytes(Np, 6)) return -7;
    if (bits == 8 && Np > 16384) return -5;
    if (Np > 32768) return -5;
    return bits == 8 ? predict_covar_i8_t<8>(Smat, Np, Ks, Mp, Vt, C, W, workspace, st)
                     : predict_covar_i8_t<7>(Smat, Np, Ks, Mp, Vt, C, W, workspace, st);
}

// bits = 8: balanced radix-256 digits (48 bits per operand instead of 42; Np <= 16384 keeps the int32 sums exact)
int trtri_i8_merge(const double* L, double* Smat, const double* DT, double* W, int Np, int left, int h, int hr, void* workspace,
                   int bits, cudaStream_t st) {
    return bits == 8 ? trtri_i8_merge_t<8>(L, Smat, DT, W, Np, left, h, hr, workspace, st)
                     : trtri_i8_merge_t<7>(L, Smat, DT, W, Np, left, h, hr, workspace, st);
}

// ---- roofline probe: the rate at which one SM retires kind::i8 MMAs from shared memory that is already resident ----
// (no TMA, no epilogue): M = 128, N = 256, K = 32 per instruction into one 256-column accumulator.
__global__ void __launch_bounds__(64, 1) peak_i8mma_kernel(int iters) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bar = reinterpret_cast<uint64_t*>(base + 24576);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
    for (int i = threadIdx.x; i < 24576 / 4; i += blockDim.x) {       // pseudo-random 7-bit slices: realistic operand toggling
        uint32_t h = (uint32_t)i * 2654435761u + blockIdx.x * 40503u;
        h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
        reinterpret_cast<uint32_t*>(base)[i] = h & 0x3f3f3f3fu;
    }
    if (threadIdx.x == 0) { mbar_init(smem_u32(bar), 1); mbar_fence_init(); }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(tmem_slot)), "n"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");     // generic-proxy stores above -> visible to the tensor core
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_d = *tmem_slot;
    if (threadIdx.x == 0) {
        const uint64_t da = umma_desc_kmajor<64>(smem_u32(base)), db = umma_desc_kmajor<64>(smem_u32(base) + 8192);
        const uint32_t idesc = i8_idesc(256);
        for (int it = 0; it < iters; ++it) {
            umma_i8(tmem_d, da, db, idesc, it != 0);
            umma_i8(tmem_d, da + 2, db + 2, idesc, 1u);
        }
        umma_commit(smem_u32(bar));
        mbar_wait(smem_u32(bar), 0);
        tc_fence_after();
    }
    __syncthreads();
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_d), "n"(512) : "memory");
    }
}

// returns the int8 operations issued (2 per multiply-add)
double peak_i8mma(int blocks, int iters, cudaStream_t st) {
    cudaFuncSetAttribute(peak_i8mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768);
    peak_i8mma_kernel<<<blocks, 64, 32768, st>>>(iters);
    return (double)blocks * iters * 2.0 * (128.0 * 256.0 * 32.0 * 2.0);
}

}  // namespace gpx
