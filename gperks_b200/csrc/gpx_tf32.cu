// "TF32 mode" of the predictive variance (north_star: rel 1e-3 vs the float64 reference): the contraction
// V = K* L^-T runs on the 5th-generation tensor cores (tcgen05.mma kind::tf32, FP32 accumulators in TMEM) as a
// 3xTF32 split product
//       A B ~= A_hi B_hi + A_hi B_lo + A_lo B_hi,      x_hi = tf32(x), x_lo = tf32(x - x_hi)
// which carries ~22 mantissa bits (single-pass TF32 fails the tolerance: var = s - sum V^2 cancels).  Only the
// variance contraction is reduced precision: the kernel matrix, the posterior mean, the Cholesky factor and
// L^-1 stay FP64; sum V^2 is accumulated per 256-row tile in FP32 and across tiles in FP64.
//
// Tile: 128 test points (MMA M, TMEM lanes) x 256 rows of L^-1 (MMA N, TMEM columns), K = training index in
// 32-wide slices (128-byte rows, SWIZZLE_128B).  Warp roles: warp 0 = TMA producer, warp 1 = TMEM allocator +
// MMA issuer (one elected thread), warps 2..5 = epilogue (one TMEM lane = one test point per thread, so the
// sum over the 256 columns needs no cross-thread reduction).
#include "gpx_common.cuh"
#include "gpx_tma.cuh"
#include "gpx_tc.cuh"
#include <stdlib.h>

namespace gpx {

constexpr int F_M = 128;                 // test points per tile
constexpr int F_N = 256;                 // L^-1 rows per tile
constexpr int F_BK = 32;                 // floats per k slice (128 B)
constexpr int F_STAGES = 2;
constexpr int F_A_BYTES = F_M * 128;     // 16 KB
constexpr int F_B_BYTES = F_N * 128;     // 32 KB
constexpr int F_STAGE_BYTES = 2 * F_A_BYTES + 2 * F_B_BYTES;   // 96 KB: A_hi, A_lo, B_hi, B_lo
constexpr int F_THREADS = 320;             // warp 0 TMA, warp 1 MMA issuer, warps 2..9 epilogue (two column halves)
constexpr int F_BAND = 4;
constexpr size_t F_SMEM_BYTES = 1024 + size_t(F_STAGES) * F_STAGE_BYTES + 128 + 2 * F_M * 4;
constexpr int F_TMEM_COLS = 512;         // two ping-pong chunk accumulators of 256 columns

__device__ __forceinline__ float tf32_round(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;\n" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}
__device__ __forceinline__ void split_tf32(double x, float& hi, float& lo) {
    hi = tf32_round((float)x);
    lo = tf32_round((float)(x - (double)hi));
}

// ---- FP64 -> (hi, lo) split (L^-1 once per factorisation) ------------------------------------------------
// ld > 0: `src` is the mirrored inverse S[ld][ld]; entries in 128-blocks above the diagonal (the L^-T mirror) are
// written as zero, because a 256-row MMA tile reads k up to the end of its second 128-block for all its rows.
__global__ void split_tf32_kernel(const double* __restrict__ src, long n, float* __restrict__ hi, float* __restrict__ lo,
                                  int ld) {
    long i = ((long)blockIdx.x * blockDim.x + threadIdx.x) * 2;
    const long stride = (long)gridDim.x * blockDim.x * 2;
    for (; i + 1 < n; i += stride) {
        double2 v = *reinterpret_cast<const double2*>(src + i);
        if (ld > 0) {
            const long row = i / ld, col = i - row * ld;        // col even, ld even: both entries in one block
            if ((col >> 7) > (row >> 7)) { v.x = 0.0; v.y = 0.0; }
        }
        float2 h, l;
        split_tf32(v.x, h.x, l.x);
        split_tf32(v.y, h.y, l.y);
        *reinterpret_cast<float2*>(hi + i) = h;
        *reinterpret_cast<float2*>(lo + i) = l;
    }
}

int split_tf32_matrix(const double* src, long n, float* hi, float* lo, int ld, cudaStream_t st) {
    if (n % 2 || (ld > 0 && ld % 2)) return -2;
    split_tf32_kernel<<<148 * 8, 256, 0, st>>>(src, n, hi, lo, ld);
    GPX_CHECK_LAUNCH();
    return 0;
}

// ---- tcgen05 helpers ----------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t umma_desc_kmajor_sw128(uint32_t smem_addr) {
    // cute::UMMA::SmemDescriptor, K-major, SWIZZLE_128B: 8-row groups 1024 B apart (SBO), LBO unused (=1),
    // version 1 (Blackwell), layout_type 2
    return (uint64_t)((smem_addr & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) |
           ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
// cute::UMMA::InstrDescriptor: D = F32, A = B = TF32, both K-major, N >> 3 at bit 17, M >> 4 at bit 24
constexpr uint32_t F_IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(F_N >> 3) << 17) | ((uint32_t)(F_M >> 4) << 24);

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(F_IDESC), "r"(accumulate)
        : "memory");
}
// ---- the kernel --------------------------------------------------------------------------------------
// grid.x = n_it * n_mb tiles; qpart[it][m] (ld = Mc) = sum over the 256 rows of tile `it` of V[m][i]^2
//
// MC = true (default): CTAs are launched as clusters of two that work on the SAME 256 rows of L^-1 and on two
// neighbouring 128-point tiles.  Each CTA fetches only its half (128 rows) of the L^-1 tile and TMA-multicasts it
// into both CTAs' shared memory, so the L2 serves every B byte once per pair: operand traffic per CTA and stage
// drops from 96 KB to 64 KB of L2 reads, which is what the single-CTA version was bound by (ncu: tensor pipe 68 %).
// A stage may only be refilled when BOTH CTAs' tensor cores are done with it => tcgen05.commit multicasts its
// arrival to both CTAs' `empty` barriers (count 2).
template <bool MC>
__global__ void __launch_bounds__(F_THREADS, 1)
trmm_sumsq_tf32_kernel(const __grid_constant__ CUtensorMap tmKhi, const __grid_constant__ CUtensorMap tmKlo,
                       const __grid_constant__ CUtensorMap tmShi, const __grid_constant__ CUtensorMap tmSlo, int Np,
                       int Mc, double* __restrict__ qpart, int kc_slices) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    // barriers: full[S], empty[S], chunk_full[2], chunk_free[2]
    uint64_t* bars = reinterpret_cast<uint64_t*>(base + size_t(F_STAGES) * F_STAGE_BYTES);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * F_STAGES + 4);
    float* red = reinterpret_cast<float*>(bars + 2 * F_STAGES + 6);        // [2][128] column-half partial sums
    const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + F_STAGES);
    const uint32_t chunk_full = smem_u32(bars + 2 * F_STAGES), chunk_free = smem_u32(bars + 2 * F_STAGES + 2);
    const uint32_t tile0 = smem_u32(base);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    const uint32_t crank = MC ? cluster_ctarank() : 0u;
    const int n_mb = MC ? Mc / (2 * F_M) : Mc / F_M, n_it = (Np + F_N - 1) / F_N;    // MC: n_mb counts PAIRS of m tiles
    int it, mb;
    {
        const int per_full_band = F_BAND * n_mb;
        const int id = MC ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
        const int band = id / per_full_band;
        const int hi = n_it - band * F_BAND;
        const int width = hi < F_BAND ? hi : F_BAND;
        const int rem = id - band * per_full_band;
        mb = rem / width;
        it = hi - 1 - (rem - mb * width);
        if (MC) mb = 2 * mb + (int)crank;
    }
    int kend = (it + 1) * F_N;
    if (kend > Np) kend = Np;
    const int nk = kend / F_BK;                 // Np is a multiple of 128
    // The tensor core accumulates in FP32 with truncation, so a long K chain drifts (measured: 3072 accumulate
    // steps at K=8192 shift sum V^2 by ~2e-4 relative).  K is therefore cut into chunks of kc_slices*32 that
    // ping-pong between two TMEM accumulators (columns [0,256) and [256,512)); while the tensor core fills one,
    // the 8 epilogue warps drain the other into a running total held in registers (round-to-nearest FP32 adds:
    // 128 columns per thread), so the flush costs the tensor core nothing.
    const int nchunks = (nk + kc_slices - 1) / kc_slices;

    if (tid == 0) {
        for (int s = 0; s < F_STAGES; ++s) {
            mbar_init(full0 + 8 * s, 1);
            mbar_init(empty0 + 8 * s, MC ? 2 : 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(chunk_full + 8 * b, 1);
            mbar_init(chunk_free + 8 * b, 8);
        }
        mbar_fence_init();
    }
    if (warp == 1) {    // TMEM: 256 columns x 128 lanes of FP32 accumulators
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(tmem_slot)), "n"(F_TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    if (MC) cluster_sync_all();        // the peer's barriers must be initialised before anything is multicast at them
    tc_fence_after();
    const uint32_t tmem_d = *tmem_slot;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            tma_prefetch_desc(&tmKhi); tma_prefetch_desc(&tmKlo); tma_prefetch_desc(&tmShi); tma_prefetch_desc(&tmSlo);
            for (int kt = 0; kt < nk; ++kt) {
                const int s = kt % F_STAGES, round = kt / F_STAGES;
                if (round > 0) mbar_wait(empty0 + 8 * s, (round - 1) & 1);
                mbar_arrive_expect_tx(full0 + 8 * s, F_STAGE_BYTES);
                const uint32_t dst = tile0 + s * F_STAGE_BYTES;
                const int k0 = kt * F_BK;
                tma_load_2d(dst, &tmKhi, k0, mb * F_M, full0 + 8 * s);
                tma_load_2d(dst + F_A_BYTES, &tmKlo, k0, mb * F_M, full0 + 8 * s);
                if (MC) {   // my half of the L^-1 tile, delivered to both CTAs of the pair
                    const uint32_t half = crank * (F_B_BYTES / 2);
                    const int row = it * F_N + (int)crank * (F_N / 2);
                    tma_load_2d_multicast(dst + 2 * F_A_BYTES + half, &tmShi, k0, row, full0 + 8 * s, 3);
                    tma_load_2d_multicast(dst + 2 * F_A_BYTES + F_B_BYTES + half, &tmSlo, k0, row, full0 + 8 * s, 3);
                } else {
                    tma_load_2d(dst + 2 * F_A_BYTES, &tmShi, k0, it * F_N, full0 + 8 * s);
                    tma_load_2d(dst + 2 * F_A_BYTES + F_B_BYTES, &tmSlo, k0, it * F_N, full0 + 8 * s);
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: one thread drives the tensor core =====
        if (lane == 0) {
            for (int kt = 0; kt < nk; ++kt) {
                const int s = kt % F_STAGES, round = kt / F_STAGES;
                const int chunk = kt / kc_slices, in_chunk = kt - chunk * kc_slices;
                const int buf = chunk & 1;
                if (in_chunk == 0 && chunk >= 2) {      // the epilogue must have drained this accumulator (chunk - 2)
                    mbar_wait(chunk_free + 8 * buf, ((chunk >> 1) - 1) & 1);
                    tc_fence_after();
                }
                mbar_wait(full0 + 8 * s, round & 1);
                tc_fence_after();
                const uint32_t st = tile0 + s * F_STAGE_BYTES;
                const uint32_t acc_addr = tmem_d + (uint32_t)(buf * F_N);
                const uint64_t a_hi = umma_desc_kmajor_sw128(st);
                const uint64_t a_lo = umma_desc_kmajor_sw128(st + F_A_BYTES);
                const uint64_t b_hi = umma_desc_kmajor_sw128(st + 2 * F_A_BYTES);
                const uint64_t b_lo = umma_desc_kmajor_sw128(st + 2 * F_A_BYTES + F_B_BYTES);
#pragma unroll
                for (int k = 0; k < F_BK / 8; ++k) {
                    const uint64_t adv = (uint64_t)((k * 32) >> 4);      // 8 tf32 = 32 B along K inside the swizzle atom
                    umma_tf32(acc_addr, a_lo + adv, b_hi + adv, (in_chunk | k) != 0);   // small terms first
                    umma_tf32(acc_addr, a_hi + adv, b_lo + adv, 1u);
                    umma_tf32(acc_addr, a_hi + adv, b_hi + adv, 1u);
                }
                if (MC) umma_commit_multicast(empty0 + 8 * s, 3);     // both CTAs' producers wait for both tensor cores
                else umma_commit(empty0 + 8 * s);                     // smem slot reusable once these MMAs retire
                if (in_chunk == kc_slices - 1 || kt == nk - 1) umma_commit(chunk_full + 8 * buf);
            }
        }
    } else {
        // ===== epilogue warps 2..9: TMEM lane quadrant = warp % 4, column half = (warp - 2) / 4 =====
        const int quad = warp & 3, half = (warp - 2) >> 2;
        const uint32_t lane_base = (uint32_t)(quad * 32) << 16;
        float tot[F_N / 2];
#pragma unroll
        for (int j = 0; j < F_N / 2; ++j) tot[j] = 0.f;
        for (int chunk = 0; chunk < nchunks; ++chunk) {
            const int buf = chunk & 1;
            mbar_wait(chunk_full + 8 * buf, (chunk >> 1) & 1);
            tc_fence_after();
#pragma unroll
            for (int c = 0; c < F_N / 2; c += 32) {
                uint32_t v[32];
                tmem_ld_32x32b_x32(tmem_d + lane_base + (uint32_t)(buf * F_N + half * (F_N / 2) + c), v);
                tmem_wait_ld();
#pragma unroll
                for (int j = 0; j < 32; ++j) tot[c + j] += __uint_as_float(v[j]);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(chunk_free + 8 * buf);
        }
        float ssum = 0.f;
#pragma unroll
        for (int j = 0; j < F_N / 2; ++j) ssum = fmaf(tot[j], tot[j], ssum);
        red[half * F_M + quad * 32 + lane] = ssum;
        named_bar_sync(1, 256);
        if (half == 0) {
            const int m = mb * F_M + quad * 32 + lane;
            qpart[(long)it * Mc + m] = (double)red[quad * 32 + lane] + (double)red[F_M + quad * 32 + lane];
        }
        tc_fence_before();
    }
    __syncthreads();
    if (MC) cluster_sync_all();        // nobody leaves while the peer may still multicast into / arrive on this CTA
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_d), "n"(F_TMEM_COLS) : "memory");
    }
}

// ---- host ------------------------------------------------------------------------------------------------
static int encode_f32_rowmajor(CUtensorMap* tm, const float* base, uint64_t rows, uint64_t cols, uint64_t ld,
                               uint32_t box_rows) {
    PFN_encodeTiled enc = get_encode_fn();
    if (!enc) return -900;
    cuuint64_t gdim[2] = {cols, rows};
    cuuint64_t gstride[1] = {ld * 4};
    cuuint32_t box[2] = {F_BK, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), gdim, gstride, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : -901 - (int)r;
}

static bool g_tf32_attr[64] = {false};

int tf32_row_tiles(int Np) { return (Np + F_N - 1) / F_N; }

static int use_tf32_multicast() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("GPX_TF32_IMPL"); v = (e && e[0] == '1') ? 0 : 1; }   // GPX_TF32_IMPL=1cta
    return v;
}

// qpart[it][m], it < tf32_row_tiles(Np).  Mc must be a multiple of 256 (CTA pairs share an L^-1 tile).
int trmm_sumsq_tf32(const float* Shi, const float* Slo, int Np, const float* Khi, const float* Klo, int Mc,
                    double* qpart, int kchunk, cudaStream_t st) {
    if (Mc % (2 * F_M)) return -6;
    int dev = 0;
    cudaGetDevice(&dev);
    if (!(dev >= 0 && dev < 64 && g_tf32_attr[dev])) {
        cudaFuncSetAttribute(trmm_sumsq_tf32_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)F_SMEM_BYTES);
        cudaFuncSetAttribute(trmm_sumsq_tf32_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)F_SMEM_BYTES);
        if (dev >= 0 && dev < 64) g_tf32_attr[dev] = true;
    }
    const bool mc = use_tf32_multicast();
    CUtensorMap tKh, tKl, tSh, tSl;
    int rc;
    if ((rc = encode_f32_rowmajor(&tKh, Khi, Mc, Np, Np, F_M))) return rc;
    if ((rc = encode_f32_rowmajor(&tKl, Klo, Mc, Np, Np, F_M))) return rc;
    if ((rc = encode_f32_rowmajor(&tSh, Shi, Np, Np, Np, mc ? F_N / 2 : F_N))) return rc;
    if ((rc = encode_f32_rowmajor(&tSl, Slo, Np, Np, Np, mc ? F_N / 2 : F_N))) return rc;
    const int tiles = tf32_row_tiles(Np) * (Mc / F_M);
    // K elements accumulated in TMEM before a flush (multiple of 32; <= 0 selects the default 32)
    if (kchunk <= 0) kchunk = 32;
    const int kc_slices = kchunk < F_BK ? 1 : kchunk / F_BK;
    if (!mc) {
        trmm_sumsq_tf32_kernel<false><<<tiles, F_THREADS, F_SMEM_BYTES, st>>>(tKh, tKl, tSh, tSl, Np, Mc, qpart, kc_slices);
    } else {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(tiles);
        cfg.blockDim = dim3(F_THREADS);
        cfg.dynamicSmemBytes = F_SMEM_BYTES;
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        cudaError_t e = cudaLaunchKernelEx(&cfg, trmm_sumsq_tf32_kernel<true>, tKh, tKl, tSh, tSl, Np, Mc, qpart, kc_slices);
        if (e != cudaSuccess) return -1000 - (int)e;
    }
    GPX_CHECK_LAUNCH();
    return 0;
}

}  // namespace gpx
