// NT tile GEMM engine, TMA edition: acc(128x128) = sum_k A[m][k] * B[n][k] on FP64 DMMA with both operand tiles
// delivered by cp.async.bulk.tensor (128 x 16 boxes, 128-byte swizzle) through a 6-stage full/empty mbarrier
// ring.  Same warp layout / fragment mapping as gemm_nt_mainloop (gpx_common.cuh); thread 0 is the producer and
// refills the ring two k-tiles behind the consumers, so there is no CTA-wide barrier in the k loop.
#pragma once
#include "gpx_common.cuh"
#include "gpx_tma.cuh"

namespace gpx {

constexpr int GT_STAGES = 6;
constexpr int GT_TILE_BYTES = TILE * G_BK * 8;          // 16 KB
constexpr int GT_STAGE_BYTES = 2 * GT_TILE_BYTES;
constexpr int GT_THREADS = 256;
constexpr size_t GT_SMEM_BYTES = 1024 + size_t(GT_STAGES) * GT_STAGE_BYTES + 2 * GT_STAGES * 8;

// Operand tile rows [row, row+128) of a row-major matrix described by `map`; element (r, k) is at tensor
// coordinates (col0 + k, row + r).  For k in [alt_k0, alt_k1) the tile comes from `alt` instead, at
// (k - alt_k0, alt_row + r) -- used to splice the transposed diagonal blocks DT into the mirrored inverse.
struct TOperand {
    const CUtensorMap* map;
    int row, col0;
    const CUtensorMap* alt;
    int alt_row, alt_k0, alt_k1;
};
__device__ __forceinline__ TOperand t_operand(const CUtensorMap* map, int row, int col0) {
    TOperand o;
    o.map = map; o.row = row; o.col0 = col0; o.alt = nullptr; o.alt_row = 0; o.alt_k0 = 0; o.alt_k1 = 0;
    return o;
}
__device__ __forceinline__ TOperand t_operand_alt(const CUtensorMap* map, int row, int col0, const CUtensorMap* alt,
                                                  int alt_row, int k0, int k1) {
    TOperand o;
    o.map = map; o.row = row; o.col0 = col0; o.alt = alt; o.alt_row = alt_row; o.alt_k0 = k0; o.alt_k1 = k1;
    return o;
}
__device__ __forceinline__ void t_issue(const TOperand& o, int k, uint32_t dst, uint32_t bar) {
    if (o.alt != nullptr && k >= o.alt_k0 && k < o.alt_k1) tma_load_2d(dst, o.alt, k - o.alt_k0, o.alt_row, bar);
    else tma_load_2d(dst, o.map, o.col0 + k, o.row, bar);
}

// `smem_raw`: dynamic shared memory of GT_SMEM_BYTES.  Must be called by all GT_THREADS threads, once per CTA.
// INIT_ZERO = false keeps the caller's accumulators (e.g. a preloaded C tile); NEG_A accumulates -A B^T (the
// negation is an operand modifier of the DMMA, not an instruction).
template <bool INIT_ZERO = true, bool NEG_A = false>
__device__ __forceinline__ void gemm_nt_tma_mainloop(double (&acc)[8][4][2], const TOperand& A, const TOperand& B,
                                                     int k_begin, int k_end, unsigned char* smem_raw) {
    unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(base + size_t(GT_STAGES) * GT_STAGE_BYTES);
    const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + GT_STAGES);
    const uint32_t tile0 = smem_u32(base);
    const int tid = threadIdx.x;
    const int nk = (k_end - k_begin) / G_BK;
    if (tid == 0) {
        for (int s = 0; s < GT_STAGES; ++s) {
            mbar_init(full0 + 8 * s, 1);
            mbar_init(empty0 + 8 * s, GT_THREADS / 32);
        }
        mbar_fence_init();
    }
    __syncthreads();
    if (tid == 0) {
        const int npro = nk < GT_STAGES ? nk : GT_STAGES;
        for (int kt = 0; kt < npro; ++kt) {
            mbar_arrive_expect_tx(full0 + 8 * kt, GT_STAGE_BYTES);
            const uint32_t dst = tile0 + kt * GT_STAGE_BYTES;
            t_issue(A, k_begin + kt * G_BK, dst, full0 + 8 * kt);
            t_issue(B, k_begin + kt * G_BK, dst + GT_TILE_BYTES, full0 + 8 * kt);
        }
    }
    const int lane = tid & 31, warp = tid >> 5;
    const int wm = warp >> 2, wn = warp & 3;
    const int lr = lane >> 2, lc = lane & 3;
    const int off0 = ((2 * lc) ^ lr) * 16, off1 = off0 ^ 16;      // 128B-swizzled 16-byte chunk positions
    if (INIT_ZERO) {
#pragma unroll
        for (int s = 0; s < 8; ++s)
#pragma unroll
            for (int n = 0; n < 4; ++n) { acc[s][n][0] = 0.0; acc[s][n][1] = 0.0; }
    }

    for (int kt = 0; kt < nk; ++kt) {
        const int s_ = kt % GT_STAGES, round = kt / GT_STAGES;
        if (tid == 0) {
            const int kl = kt + GT_STAGES - 2;
            if (kt >= 2 && kl < nk) {
                const int sl = kl % GT_STAGES;
                mbar_wait(empty0 + 8 * sl, ((kl / GT_STAGES) - 1) & 1);
                mbar_arrive_expect_tx(full0 + 8 * sl, GT_STAGE_BYTES);
                const uint32_t dst = tile0 + sl * GT_STAGE_BYTES;
                t_issue(A, k_begin + kl * G_BK, dst, full0 + 8 * sl);
                t_issue(B, k_begin + kl * G_BK, dst + GT_TILE_BYTES, full0 + 8 * sl);
            }
        }
        __syncwarp();
        mbar_wait(full0 + 8 * s_, round & 1);
        const unsigned char* at = base + s_ * GT_STAGE_BYTES + (wm * 64 + lr) * 128;
        const unsigned char* bt = base + s_ * GT_STAGE_BYTES + GT_TILE_BYTES + (wn * 32 + lr) * 128;
        double bf[4][4];
#pragma unroll
        for (int n = 0; n < 4; ++n) {
            double2 v0 = *reinterpret_cast<const double2*>(bt + n * 1024 + off0);
            double2 v1 = *reinterpret_cast<const double2*>(bt + n * 1024 + off1);
            bf[n][0] = v0.x; bf[n][1] = v0.y; bf[n][2] = v1.x; bf[n][3] = v1.y;
        }
#pragma unroll
        for (int s = 0; s < 8; ++s) {
            double2 a0 = *reinterpret_cast<const double2*>(at + s * 1024 + off0);
            double2 a1 = *reinterpret_cast<const double2*>(at + s * 1024 + off1);
            double af[4] = {a0.x, a0.y, a1.x, a1.y};
            if (NEG_A) { af[0] = -af[0]; af[1] = -af[1]; af[2] = -af[2]; af[3] = -af[3]; }
#pragma unroll
            for (int st = 0; st < 4; ++st)
#pragma unroll
                for (int n = 0; n < 4; ++n) dmma884(acc[s][n][0], acc[s][n][1], af[st], bf[n][st]);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(empty0 + 8 * s_);
    }
}

// ------------------------------------------------------------------------------------------------------------------
// Same engine for narrower output tiles: (WM * S * 8) rows x 128 columns, 8 warps as WM x (8 / WM), each warp S x NB
// 8x8 blocks.  Used where only a few tiles exist and the kernel sits on a critical path (potrf's panel and look-ahead
// column): 64- or 32-row tiles spread the same work over 2-4x more SMs.  The A operand needs a tensor map whose box has
// WM*S*8 rows; B keeps the 128-row box.
//   <2, 4, 4>:  64 x 128        <1, 4, 2>:  32 x 128
// ------------------------------------------------------------------------------------------------------------------
template <int WM, int S, int NB, bool INIT_ZERO = true, bool NEG_A = false>
__device__ __forceinline__ void gemm_nt_tma_mainloop_t(double (&acc)[S][NB][2], const TOperand& A, const TOperand& B,
                                                       int k_begin, int k_end, unsigned char* smem_raw) {
    constexpr int WN = 8 / WM;
    static_assert(WN * NB * 8 == TILE, "B tile is 128 rows");
    constexpr int A_BYTES = WM * S * 8 * G_BK * 8;
    constexpr int STAGE_BYTES = A_BYTES + GT_TILE_BYTES;
    static_assert(A_BYTES % 1024 == 0, "swizzle atoms");
    unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(base + size_t(GT_STAGES) * STAGE_BYTES);
    const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + GT_STAGES);
    const uint32_t tile0 = smem_u32(base);
    const int tid = threadIdx.x;
    const int nk = (k_end - k_begin) / G_BK;
    if (tid == 0) {
        for (int s = 0; s < GT_STAGES; ++s) {
            mbar_init(full0 + 8 * s, 1);
            mbar_init(empty0 + 8 * s, GT_THREADS / 32);
        }
        mbar_fence_init();
    }
    __syncthreads();
    if (tid == 0) {
        const int npro = nk < GT_STAGES ? nk : GT_STAGES;
        for (int kt = 0; kt < npro; ++kt) {
            mbar_arrive_expect_tx(full0 + 8 * kt, STAGE_BYTES);
            const uint32_t dst = tile0 + kt * STAGE_BYTES;
            t_issue(A, k_begin + kt * G_BK, dst, full0 + 8 * kt);
            t_issue(B, k_begin + kt * G_BK, dst + A_BYTES, full0 + 8 * kt);
        }
    }
    const int lane = tid & 31, warp = tid >> 5;
    const int wm = warp / WN, wn = warp % WN;
    const int lr = lane >> 2, lc = lane & 3;
    const int off0 = ((2 * lc) ^ lr) * 16, off1 = off0 ^ 16;
    if (INIT_ZERO) {
#pragma unroll
        for (int s = 0; s < S; ++s)
#pragma unroll
            for (int n = 0; n < NB; ++n) { acc[s][n][0] = 0.0; acc[s][n][1] = 0.0; }
    }
    for (int kt = 0; kt < nk; ++kt) {
        const int s_ = kt % GT_STAGES, round = kt / GT_STAGES;
        if (tid == 0) {
            const int kl = kt + GT_STAGES - 2;
            if (kt >= 2 && kl < nk) {
                const int sl = kl % GT_STAGES;
                mbar_wait(empty0 + 8 * sl, ((kl / GT_STAGES) - 1) & 1);
                mbar_arrive_expect_tx(full0 + 8 * sl, STAGE_BYTES);
                const uint32_t dst = tile0 + sl * STAGE_BYTES;
                t_issue(A, k_begin + kl * G_BK, dst, full0 + 8 * sl);
                t_issue(B, k_begin + kl * G_BK, dst + A_BYTES, full0 + 8 * sl);
            }
        }
        __syncwarp();
        mbar_wait(full0 + 8 * s_, round & 1);
        const unsigned char* at = base + s_ * STAGE_BYTES + (wm * S * 8 + lr) * 128;
        const unsigned char* bt = base + s_ * STAGE_BYTES + A_BYTES + (wn * NB * 8 + lr) * 128;
        double bf[NB][4];
#pragma unroll
        for (int n = 0; n < NB; ++n) {
            double2 v0 = *reinterpret_cast<const double2*>(bt + n * 1024 + off0);
            double2 v1 = *reinterpret_cast<const double2*>(bt + n * 1024 + off1);
            bf[n][0] = v0.x; bf[n][1] = v0.y; bf[n][2] = v1.x; bf[n][3] = v1.y;
        }
#pragma unroll
        for (int s = 0; s < S; ++s) {
            double2 a0 = *reinterpret_cast<const double2*>(at + s * 1024 + off0);
            double2 a1 = *reinterpret_cast<const double2*>(at + s * 1024 + off1);
            double af[4] = {a0.x, a0.y, a1.x, a1.y};
            if (NEG_A) { af[0] = -af[0]; af[1] = -af[1]; af[2] = -af[2]; af[3] = -af[3]; }
#pragma unroll
            for (int st = 0; st < 4; ++st)
#pragma unroll
                for (int n = 0; n < NB; ++n) dmma884(acc[s][n][0], acc[s][n][1], af[st], bf[n][st]);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(empty0 + 8 * s_);
    }
}

// accumulator element -> (row, col) inside the (WM*S*8) x 128 tile; f(row, col, v0&, v1&), col even
template <int WM, int S, int NB, typename F>
__device__ __forceinline__ void gemm_foreach_pair_t(double (&acc)[S][NB][2], F f) {
    constexpr int WN = 8 / WM;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int wm = warp / WN, wn = warp % WN;
    const int lr = lane >> 2, lc = lane & 3;
#pragma unroll
    for (int s = 0; s < S; ++s)
#pragma unroll
        for (int n = 0; n < NB; ++n) f(wm * S * 8 + s * 8 + lr, wn * NB * 8 + n * 8 + 2 * lc, acc[s][n][0], acc[s][n][1]);
}

}  // namespace gpx
