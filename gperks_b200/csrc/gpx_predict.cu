// Predictive mean + (noisy) variance over large point sets -- the north-star path.
// Replaces GPEmulator.predict's  likelihood(model(X_new))  (GPErks/train/emulator.py:348-363):
//   mu*  = m(x*) + K* alpha
//   var* = s - || L^-1 K*^T ||^2 + noise                     (then un-standardised: sigma_y, mu_y)
//
// Per chunk of Mc test points:
//   1. kmat_cross            K*[m][j]  (+ fused mean partial sums)            gpx_kmat.cu
//   2. trmm_sumsq_kernel     V = L^-1 K*^T as an NT triangular GEMM on FP64 DMMA; V is never
//                            written: each 128x128 tile is squared and column-reduced in the
//                            epilogue into qpart[i_blk][m]
//   3. predict_finalize      deterministic sums over the partials, prior mean, clamp, sqrt, unscale
#include "gpx_common.cuh"
#include "gpx_tma.cuh"
#include "gpx_gemm_tma.cuh"
#include <stdlib.h>

namespace gpx {

// grid.x = nb * (Mc/128); heavy (long-k) row blocks are scheduled first.
__global__ void __launch_bounds__(G_THREADS, 1)
trmm_sumsq_kernel(const double* __restrict__ S, int Np, const double* __restrict__ Ks, int Mc,
                  double* __restrict__ qpart) {
    extern __shared__ double smem[];
    const int n_mb = Mc / TILE;
    const int nb = Np / TILE;
    const int ib = nb - 1 - (int)(blockIdx.x / n_mb);
    const int mb = blockIdx.x % n_mb;
    Operand A = make_operand(S + ((long)ib * TILE) * Np, Np);        // rows i of L^-1 (diag block masked)
    Operand B = make_operand(Ks + ((long)mb * TILE) * Np, Np);       // rows m of K*
    double acc[8][4][2];
    gemm_nt_mainloop(acc, A, B, 0, (ib + 1) * TILE, smem);
    // column sums of squares: this thread holds rows {wm*64 + s*8 + lr}, cols wn*32 + n*8 + 2*lc + e
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int wm = warp >> 2, wn = warp & 3;
    const int lc = lane & 3;
    double* red = smem;   // [2][128]
#pragma unroll
    for (int n = 0; n < 4; ++n) {
        double c0 = 0.0, c1 = 0.0;
#pragma unroll
        for (int s = 0; s < 8; ++s) {
            c0 = fma(acc[s][n][0], acc[s][n][0], c0);
            c1 = fma(acc[s][n][1], acc[s][n][1], c1);
        }
        // reduce over the 8 row-lanes (lane bits 2..4)
#pragma unroll
        for (int o = 4; o < 32; o <<= 1) {
            c0 += __shfl_xor_sync(0xffffffffu, c0, o);
            c1 += __shfl_xor_sync(0xffffffffu, c1, o);
        }
        if ((lane >> 2) == 0) {
            red[wm * TILE + wn * 32 + n * 8 + 2 * lc] = c0;
            red[wm * TILE + wn * 32 + n * 8 + 2 * lc + 1] = c1;
        }
    }
    __syncthreads();
    if (threadIdx.x < TILE)
        qpart[(long)ib * Mc + (long)mb * TILE + threadIdx.x] = red[threadIdx.x] + red[TILE + threadIdx.x];
}


// ------------------------------------------------------------------------------------------------
// TMA version of the variance contraction (default).  Same tile / warp layout and DMMA inner loop as
// trmm_sumsq_kernel, but operand tiles arrive by cp.async.bulk.tensor (one elected producer thread, 128-byte
// swizzled 128 x 16 boxes) through a 6-stage full/empty mbarrier ring, so the 8 MMA warps issue nothing but
// LDS.128 + DMMA and never meet at a CTA-wide barrier inside the k loop.
//
// Rasterisation: row blocks are walked in bands of T_BAND (heaviest band first); inside a band consecutive
// CTAs share the K* panel (same mb) and differ in ib, so a resident wave touches ~T_BAND L^-1 panels and
// ~148/T_BAND K* panels instead of 1 and 148: each K* panel is fetched from HBM once per band, not once
// per row block.
// ------------------------------------------------------------------------------------------------
constexpr int T_STAGES = 6;
constexpr int T_TILE_BYTES = TILE * G_BK * 8;            // 16 KB: 128 rows x 128 B
constexpr int T_STAGE_BYTES = 2 * T_TILE_BYTES;
constexpr int T_CONSUMERS = 256;
constexpr int T_THREADS = T_CONSUMERS;      // no dedicated producer warp: thread 0 refills the ring 2 tiles behind
constexpr int T_BAND = 8;
constexpr size_t T_SMEM_BYTES = 1024 + size_t(T_STAGES) * T_STAGE_BYTES + 2 * TILE * 8 + 2 * T_STAGES * 8;

__global__ void __launch_bounds__(T_THREADS, 1)
trmm_sumsq_tma_kernel(const __grid_constant__ CUtensorMap tmS, const __grid_constant__ CUtensorMap tmK, int Np, int Mc,
                      double* __restrict__ qpart) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    double* red = reinterpret_cast<double*>(base + size_t(T_STAGES) * T_STAGE_BYTES);       // [2][128]
    uint64_t* bars = reinterpret_cast<uint64_t*>(red + 2 * TILE);                           // full[S], empty[S]
    const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + T_STAGES);
    const uint32_t tile0 = smem_u32(base);

    const int n_mb = Mc / TILE, nb = Np / TILE;
    // band-major rasterisation (see above)
    int id = blockIdx.x;
    int ib, mb;
    {
        const int per_full_band = T_BAND * n_mb;
        int band = id / per_full_band;
        int hi = nb - band * T_BAND;                 // row blocks [hi - width, hi)
        int width = hi < T_BAND ? hi : T_BAND;
        int rem = id - band * per_full_band;
        mb = rem / width;
        ib = hi - 1 - (rem - mb * width);
    }
    const int nk = (ib + 1) * (TILE / G_BK);
    const int tid = threadIdx.x;
    if (tid == 0) {
        for (int s = 0; s < T_STAGES; ++s) {
            mbar_init(full0 + 8 * s, 1);
            mbar_init(empty0 + 8 * s, T_CONSUMERS / 32);
        }
        mbar_fence_init();
    }
    __syncthreads();

    // prologue: thread 0 fills every stage of the ring
    if (tid == 0) {
        tma_prefetch_desc(&tmS);
        tma_prefetch_desc(&tmK);
        const int npro = nk < T_STAGES ? nk : T_STAGES;
        for (int kt = 0; kt < npro; ++kt) {
            mbar_arrive_expect_tx(full0 + 8 * kt, T_STAGE_BYTES);
            const uint32_t dst = tile0 + kt * T_STAGE_BYTES;
            tma_load_2d(dst, &tmS, kt * G_BK, ib * TILE, full0 + 8 * kt);
            tma_load_2d(dst + T_TILE_BYTES, &tmK, kt * G_BK, mb * TILE, full0 + 8 * kt);
        }
    }

    // ===== 8 MMA warps =====
    const int lane = tid & 31, warp = tid >> 5;
    const int wm = warp >> 2, wn = warp & 3;
    const int lr = lane >> 2, lc = lane & 3;
    const int off0 = ((2 * lc) ^ lr) * 16, off1 = off0 ^ 16;      // 128B-swizzled 16-byte chunk positions
    double acc[8][4][2];
#pragma unroll
    for (int s = 0; s < 8; ++s)
#pragma unroll
        for (int n = 0; n < 4; ++n) { acc[s][n][0] = 0.0; acc[s][n][1] = 0.0; }

    for (int kt = 0; kt < nk; ++kt) {
        const int s_ = kt % T_STAGES, round = kt / T_STAGES;
        if (tid == 0) {
            // refill the stage that every warp released two iterations ago with k-tile kt + S - 2
            const int kl = kt + T_STAGES - 2;
            if (kt >= 2 && kl < nk) {
                const int sl = kl % T_STAGES;
                mbar_wait(empty0 + 8 * sl, ((kl / T_STAGES) - 1) & 1);
                mbar_arrive_expect_tx(full0 + 8 * sl, T_STAGE_BYTES);
                const uint32_t dst = tile0 + sl * T_STAGE_BYTES;
                tma_load_2d(dst, &tmS, kl * G_BK, ib * TILE, full0 + 8 * sl);
                tma_load_2d(dst + T_TILE_BYTES, &tmK, kl * G_BK, mb * TILE, full0 + 8 * sl);
            }
        }
        __syncwarp();
        mbar_wait(full0 + 8 * s_, round & 1);
        const unsigned char* at = base + s_ * T_STAGE_BYTES + (wm * 64 + lr) * 128;
        const unsigned char* bt = base + s_ * T_STAGE_BYTES + T_TILE_BYTES + (wn * 32 + lr) * 128;
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
#pragma unroll
            for (int st = 0; st < 4; ++st)
#pragma unroll
                for (int n = 0; n < 4; ++n) dmma884(acc[s][n][0], acc[s][n][1], af[st], bf[n][st]);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(empty0 + 8 * s_);
    }

    // epilogue: column sums of squares over the 128 rows of the tile
#pragma unroll
    for (int n = 0; n < 4; ++n) {
        double c0 = 0.0, c1 = 0.0;
#pragma unroll
        for (int s = 0; s < 8; ++s) {
            c0 = fma(acc[s][n][0], acc[s][n][0], c0);
            c1 = fma(acc[s][n][1], acc[s][n][1], c1);
        }
#pragma unroll
        for (int o = 4; o < 32; o <<= 1) {
            c0 += __shfl_xor_sync(0xffffffffu, c0, o);
            c1 += __shfl_xor_sync(0xffffffffu, c1, o);
        }
        if (lr == 0) {
            red[wm * TILE + wn * 32 + n * 8 + 2 * lc] = c0;
            red[wm * TILE + wn * 32 + n * 8 + 2 * lc + 1] = c1;
        }
    }
    named_bar_sync(1, T_CONSUMERS);
    if (tid < TILE) qpart[(long)ib * Mc + (long)mb * TILE + tid] = red[tid] + red[TILE + tid];
}

static int use_tma_impl() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("GPX_TRMM_IMPL");
        v = (e && e[0] == 'c') ? 0 : 1;          // GPX_TRMM_IMPL=cpasync selects the LDGSTS version
    }
    return v;
}

static int launch_trmm(const double* S, int Np, const double* Ks, int Mc, double* qpart, cudaStream_t st) {
    if (!use_tma_impl()) {
        trmm_sumsq_kernel<<<(Np / TILE) * (Mc / TILE), G_THREADS, G_SMEM_BYTES, st>>>(S, Np, Ks, Mc, qpart);
        return 0;
    }
    CUtensorMap tmS, tmK;
    int rc = encode_f64_rowmajor(&tmS, S, Np, Np, Np, TILE);
    if (rc) return rc;
    rc = encode_f64_rowmajor(&tmK, Ks, Mc, Np, Np, TILE);
    if (rc) return rc;
    trmm_sumsq_tma_kernel<<<(Np / TILE) * (Mc / TILE), T_THREADS, T_SMEM_BYTES, st>>>(tmS, tmK, Np, Mc, qpart);
    return 0;
}

// One thread per test point of the chunk.
//   mean = prior(x*) + sum_jb mean_part[jb][m];  var = max(s + noise - sum_ib qpart[ib][m], min_var)
//   out_mean = y_sigma*mean + y_mu ; out_std = y_sigma*sqrt(var)
// prior: either a precomputed array prior_mean[m] or the polynomial mean sum_p w_p prod_q x[idx[p][q]] + b
// evaluated on the unit-cube-scaled inputs (GPErks/gp/mean.py:29-36).
__global__ void predict_finalize_kernel(const double* __restrict__ Xs, long m_valid, int D,
                                        const double* __restrict__ xa, const double* __restrict__ xr,
                                        const double* __restrict__ theta, const double* __restrict__ mean_part,
                                        const double* __restrict__ qpart, int nb, int Mc,
                                        const double* __restrict__ prior_mean, const int* __restrict__ feat_idx,
                                        int n_feat, int degree, const double* __restrict__ weights,
                                        const double* __restrict__ bias, double y_mu, double y_sigma, double min_var,
                                        double* __restrict__ out_mean, double* __restrict__ out_std, int nq) {
    const long m = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= m_valid) return;
    double mean = 0.0;
    if (prior_mean != nullptr) {
        mean = prior_mean[m];
    } else {
        if (weights != nullptr) {
            for (int p = 0; p < n_feat; ++p) {
                double f = 1.0;
                for (int q = 0; q < degree; ++q) {
                    int d = feat_idx[p * degree + q];
                    if (d >= 0) {
                        double x = Xs[m * D + d];
                        if (xa != nullptr) x = (x - xa[d]) / xr[d];
                        f *= x;
                    }
                }
                mean = fma(weights[p], f, mean);
            }
        }
        if (bias != nullptr) mean += bias[0];
    }
    double km = 0.0, q = 0.0;
    for (int b = 0; b < nb; ++b) km += mean_part[(long)b * Mc + m];
    for (int b = 0; b < nq; ++b) q += qpart[(long)b * Mc + m];
    mean += km;
    double var = theta[D] + theta[D + 1] - q;
    var = fmax(var, min_var);
    out_mean[m] = y_sigma * mean + y_mu;
    out_std[m] = y_sigma * sqrt(var);
}

// cudaFuncSetAttribute is per device: remember which devices have been configured
static bool g_pred_attr_dev[64] = {false};
static void set_pred_attrs();

int kmat_cross(const double* Xs, long m_valid, const double* X, int N, int D, const double* theta, const double* xa,
               const double* xr, int kind, const double* alpha, double* Ks, int Np, double* mean_part, int Mc,
               cudaStream_t st);

size_t predict_workspace_bytes(int Np, int Mc) {
    const size_t nb = Np / TILE;
    return (size_t)Mc * Np * 8 + 2 * nb * (size_t)Mc * 8;
}

// The variance TRMM alone (bench / tests): qpart[ib][m] = sum_{i in row block ib} (L^-1 K*^T)[i][m]^2
int trmm_sumsq(const double* S, int Np, const double* Ks, int Mc, double* qpart, cudaStream_t st) {
    set_pred_attrs();
    int rc = launch_trmm(S, Np, Ks, Mc, qpart, st);
    if (rc) return rc;
    GPX_CHECK_LAUNCH();
    return 0;
}

// Device-resident predict over M points, processed in chunks of Mc (multiple of 128).
int predict_meanvar(const double* Xs, long M, const double* X, int N, int D, const double* theta, const double* xa,
                    const double* xr, int kind, const double* S, int Np, const double* alpha,
                    const double* prior_mean, const int* feat_idx, int n_feat, int degree, const double* weights,
                    const double* bias, double y_mu, double y_sigma, double min_var, double* out_mean,
                    double* out_std, void* workspace, size_t workspace_bytes, int Mc, cudaStream_t st) {
    if (Mc <= 0 || Mc % TILE != 0) return -25;
    if (Np % TILE != 0 || Np < N) return -11;
    if (D > MAXD || D < 1) return -5;
    if (workspace_bytes < predict_workspace_bytes(Np, Mc)) return -23;
    set_pred_attrs();
    const int nb = Np / TILE;
    double* Ks = reinterpret_cast<double*>(workspace);
    double* mean_part = Ks + (size_t)Mc * Np;
    double* qpart = mean_part + (size_t)nb * Mc;
    for (long m0 = 0; m0 < M; m0 += Mc) {
        const long mv = (M - m0 < Mc) ? (M - m0) : Mc;
        const int mc = round_up((int)mv, TILE);      // only the tiles that hold valid points
        const double* xs = Xs + m0 * D;
        int rc = kmat_cross(xs, mv, X, N, D, theta, xa, xr, kind, alpha, Ks, Np, mean_part, mc, st);
        if (rc) return rc;
        rc = launch_trmm(S, Np, Ks, mc, qpart, st);
        if (rc) return rc;
        predict_finalize_kernel<<<(int)((mv + 255) / 256), 256, 0, st>>>(
            xs, mv, D, xa, xr, theta, mean_part, qpart, nb, mc, prior_mean ? prior_mean + m0 : nullptr, feat_idx,
            n_feat, degree, weights, bias, y_mu, y_sigma, min_var, out_mean + m0, out_std + m0, nb);
    }
    GPX_CHECK_LAUNCH();
    return 0;
}

// ---- dense predictive covariance (with_covar / sample / validation loss; test-set sized M) ---------------------
// Vt[m][i] = sum_{k <= i} Ks[m][k] * Linv[i][k]   (= (L^-1 K*^T)^T, rows = test points), tiles (mb, ib)
__global__ void __launch_bounds__(GT_THREADS, 1)
trmm_store_tma_kernel(const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmS,
                      double* __restrict__ Vt, int Np) {
    extern __shared__ unsigned char smem_raw[];
    const int ib = blockIdx.x, mb = blockIdx.y;
    TOperand A = t_operand(&tmK, mb * TILE, 0);
    TOperand B = t_operand(&tmS, ib * TILE, 0);
    double acc[8][4][2];
    gemm_nt_tma_mainloop(acc, A, B, 0, (ib + 1) * TILE, smem_raw);
    double* Ct = Vt + ((long)mb * TILE) * Np + (long)ib * TILE;
    gemm_foreach_pair(acc, [&](int row, int col, double v0, double v1) {
        *reinterpret_cast<double2*>(Ct + (long)row * Np + col) = make_double2(v0, v1);
    });
}

// C[m1][m2] -= sum_k Vt[m1][k] * Vt[m2][k]   (C = K** on entry, Mp x Mp), all tiles
__global__ void __launch_bounds__(GT_THREADS, 1)
syrk_sub_tma_kernel(const __grid_constant__ CUtensorMap tmV, double* __restrict__ C, int Mp, int Np) {
    extern __shared__ unsigned char smem_raw[];
    const int t2 = blockIdx.x, t1 = blockIdx.y;
    TOperand A = t_operand(&tmV, t1 * TILE, 0);
    TOperand B = t_operand(&tmV, t2 * TILE, 0);
    double acc[8][4][2];
    gemm_nt_tma_mainloop(acc, A, B, 0, Np, smem_raw);
    double* Ct = C + ((long)t1 * TILE) * Mp + (long)t2 * TILE;
    gemm_foreach_pair(acc, [&](int row, int col, double v0, double v1) {
        double2* ptr = reinterpret_cast<double2*>(Ct + (long)row * Mp + col);
        double2 c = *ptr;
        c.x -= v0; c.y -= v1;
        *ptr = c;
    });
}

static bool g_cov_attr[64] = {false};
static void set_cov_attrs() {
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && g_cov_attr[dev]) return;
    cudaFuncSetAttribute(trmm_store_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GT_SMEM_BYTES);
    cudaFuncSetAttribute(syrk_sub_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GT_SMEM_BYTES);
    if (dev >= 0 && dev < 64) g_cov_attr[dev] = true;
}

int trmm_store(const double* S, int Np, const double* Ks, int Mp, double* Vt, cudaStream_t st) {
    set_cov_attrs();
    CUtensorMap tmK, tmS;
    int rc;
    if ((rc = encode_f64_rowmajor(&tmK, Ks, Mp, Np, Np, TILE))) return rc;
    if ((rc = encode_f64_rowmajor(&tmS, S, Np, Np, Np, TILE))) return rc;
    trmm_store_tma_kernel<<<dim3(Np / TILE, Mp / TILE), GT_THREADS, GT_SMEM_BYTES, st>>>(tmK, tmS, Vt, Np);
    GPX_CHECK_LAUNCH();
    return 0;
}

int syrk_sub(const double* Vt, int Mp, int Np, double* C, cudaStream_t st) {
    set_cov_attrs();
    CUtensorMap tmV;
    int rc;
    if ((rc = encode_f64_rowmajor(&tmV, Vt, Mp, Np, Np, TILE))) return rc;
    syrk_sub_tma_kernel<<<dim3(Mp / TILE, Mp / TILE), GT_THREADS, GT_SMEM_BYTES, st>>>(tmV, C, Mp, Np);
    GPX_CHECK_LAUNCH();
    return 0;
}

// ---- TF32 mode (tcgen05, 3xTF32) ------------------------------------------------------------------------
int kmat_cross_split(const double* Xs, long m_valid, const double* X, int N, int D, const double* theta,
                     const double* xa, const double* xr, int kind, const double* alpha, float* Khi, float* Klo, int Np,
                     double* mean_part, int Mc, int fp32_eval, cudaStream_t st);
int trmm_sumsq_tf32(const float* Shi, const float* Slo, int Np, const float* Khi, const float* Klo, int Mc,
                    double* qpart, int kchunk, cudaStream_t st);
int tf32_row_tiles(int Np);

size_t predict_tf32_workspace_bytes(int Np, int Mc) {
    const size_t nb = Np / TILE;
    return (size_t)Mc * Np * 8 + (nb + tf32_row_tiles(Np)) * (size_t)Mc * 8;
}

int predict_meanvar_tf32(const double* Xs, long M, const double* X, int N, int D, const double* theta, const double* xa,
                         const double* xr, int kind, const float* Shi, const float* Slo, int Np, const double* alpha,
                         const double* prior_mean, const int* feat_idx, int n_feat, int degree, const double* weights,
                         const double* bias, double y_mu, double y_sigma, double min_var, double* out_mean,
                         double* out_std, void* workspace, size_t workspace_bytes, int Mc, int kchunk,
                         cudaStream_t st) {
    if (Mc <= 0 || Mc % (2 * TILE) != 0) return -25;      // CTA pairs of the tcgen05 kernel cover 256 points
    if (Np % TILE != 0 || Np < N) return -11;
    if (D > MAXD || D < 1) return -5;
    if (workspace_bytes < predict_tf32_workspace_bytes(Np, Mc)) return -23;
    const int nb = Np / TILE, nt = tf32_row_tiles(Np);
    // kchunk < 0: kernel values evaluated in FP64 then split (posterior mean bit-identical to the FP64 path);
    // kchunk >= 0: evaluated in FP32 (default of the mode, ~5x cheaper builder)
    const int fp32_eval = kchunk >= 0;
    const int kc = kchunk < 0 ? -kchunk : kchunk;
    float* Khi = reinterpret_cast<float*>(workspace);
    float* Klo = Khi + (size_t)Mc * Np;
    double* mean_part = reinterpret_cast<double*>(Klo + (size_t)Mc * Np);
    double* qpart = mean_part + (size_t)nb * Mc;
    for (long m0 = 0; m0 < M; m0 += Mc) {
        const long mv = (M - m0 < Mc) ? (M - m0) : Mc;
        const int mc = round_up((int)mv, 2 * TILE);
        const double* xs = Xs + m0 * D;
        int rc = kmat_cross_split(xs, mv, X, N, D, theta, xa, xr, kind, alpha, Khi, Klo, Np, mean_part, mc, fp32_eval, st);
        if (rc) return rc;
        rc = trmm_sumsq_tf32(Shi, Slo, Np, Khi, Klo, mc, qpart, kc, st);
        if (rc) return rc;
        // mean partials come in nb row blocks, variance partials in nt 256-row tiles
        predict_finalize_kernel<<<(int)((mv + 255) / 256), 256, 0, st>>>(
            xs, mv, D, xa, xr, theta, mean_part, qpart, nb, mc, prior_mean ? prior_mean + m0 : nullptr, feat_idx,
            n_feat, degree, weights, bias, y_mu, y_sigma, min_var, out_mean + m0, out_std + m0, nt);
    }
    GPX_CHECK_LAUNCH();
    return 0;
}

// ---- sliced-integer mode (tcgen05 kind::i8, gpx_i8.cu) -------------------------------------------------------
int trmm_sumsq_i8(const int8_t* Ls, const double* inv_scale, int Np, const int8_t* Kp, int Mc, int nslices, int bits,
                  const double* amax_ptr, double* qpart, int* sched, cudaStream_t st);
int i8_row_tiles(int Np);
int i8_max_np(int bits);
int kmat_cross_i8(const double* Xs, long m_valid, const double* X, int N, int D, const double* theta, const double* xa,
                  const double* xr, int kind, const double* alpha, int8_t* planes, long plane_rows, int nslices, int bits, int Np,
                  double* mean_part, int Mc, cudaStream_t st);

int implausibility(const double* mean, const double* std_, long M, long ld, int E, const double* z, const double* v, int maxno,
                   double* I, double* PV, cudaStream_t st);

// ---- the chunk pipeline shared by the one-emulator predict and the multi-emulator history-matching wave ----------------
// A work item = (emulator, chunk of <= Mc test points).  With `mma` (a second, higher-priority stream owned by the caller) the
// tensor-core kernel of item i runs on `mma` while the K* builder of item i+1 runs on `st`: the builder is FP64-ALU work in
// <= 128 registers and a few KB of shared memory per CTA, so its CTAs are resident on the same SMs as the tensor-core kernel's
// (one 188 KB CTA per SM for five slices).  Order on `st`: builder(0), builder(1), [wait mma(0)] finalize(0), builder(2),
// [wait mma(1)] finalize(1), ... -- everything the caller sees is ordered on `st`; buffer (i & 1) is rewritten by builder(i+2)
// only after finalize(i).  mma == nullptr (or a single item): the same launches back to back on `st`.
struct I8Emulator {          // device pointers of one factorised emulator (mirrors gpx_i8_emulator of include/gpx.h)
    const double* X; const double* theta; const double* xa; const double* xr; const double* alpha;
    const int8_t* Ls; const double* inv_scale;
    const int* feat_idx; const double* weights; const double* bias;
    double y_mu, y_sigma, min_var;
    int N, Np, kind, nslices, bits, n_feat, degree, reserved;
};
static_assert(sizeof(I8Emulator) == 10 * 8 + 3 * 8 + 8 * 4, "I8Emulator must match gpx_i8_emulator (include/gpx.h)");
struct I8Item {
    const I8Emulator* em;
    const double* xs;        // first test point of the chunk (raw inputs, D columns)
    long mv;                 // valid points
    const double* prior;     // prior mean of those points, or null (polynomial / none)
    double* out_mean;
    double* out_std;
};

static size_t i8_item_bytes(int Np, int Mc, int nslices) {
    const size_t nb = Np / TILE;
    return (size_t)nslices * Mc * Np + (nb + i8_row_tiles(Np)) * (size_t)Mc * 8;
}

template <class ItemAt, class AfterFinalize>
static int run_i8_pipeline(long n_items, ItemAt item_at, AfterFinalize after_finalize, int D, size_t buf_bytes, void* workspace,
                           cudaStream_t st, cudaStream_t mma) {
    // workspace = [256 B: the two scheduler counters of the tensor-core kernel | item buffer 0 | item buffer 1 | caller's tail]
    int* sched = reinterpret_cast<int*>(workspace);
    cudaMemsetAsync(sched, 0, 2 * sizeof(int), st);      // ordered before the first tensor-core launch through its build event
    unsigned char* buf[2] = {reinterpret_cast<unsigned char*>(workspace) + 256, reinterpret_cast<unsigned char*>(workspace) + 256 + buf_bytes};
    const bool overlap = mma != nullptr && mma != st && n_items > 1;
    cudaEvent_t ev_built[2] = {nullptr, nullptr}, ev_mma[2] = {nullptr, nullptr};
    if (overlap) {
        for (int i = 0; i < 2; ++i) {
            if (cudaEventCreateWithFlags(&ev_built[i], cudaEventDisableTiming) != cudaSuccess ||
                cudaEventCreateWithFlags(&ev_mma[i], cudaEventDisableTiming) != cudaSuccess)
                return -31;
        }
    }
    int rc = 0;
    // item buffer = [slice planes of the item | mean_part | qpart], sized by the caller for the largest item
    auto parts = [&](long i, const I8Item& it, int8_t*& Kp, double*& mean_part, double*& qpart, int& mc) {
        const I8Emulator& e = *it.em;
        mc = round_up((int)it.mv, TILE);
        Kp = reinterpret_cast<int8_t*>(buf[i & 1]);
        const size_t plane_bytes = ((size_t)e.nslices * mc * e.Np + 255) / 256 * 256;
        mean_part = reinterpret_cast<double*>(buf[i & 1] + plane_bytes);
        qpart = mean_part + (size_t)(e.Np / TILE) * mc;
    };
    auto build = [&](long i) -> int {
        const I8Item it = item_at(i);
        const I8Emulator& e = *it.em;
        int8_t* Kp; double *mean_part, *qpart; int mc;
        parts(i, it, Kp, mean_part, qpart, mc);
        int r = kmat_cross_i8(it.xs, it.mv, e.X, e.N, D, e.theta, e.xa, e.xr, e.kind, e.alpha, Kp, mc, e.nslices, e.bits, e.Np,
                              mean_part, mc, st);
        if (r == 0 && overlap) {
            cudaEventRecord(ev_built[i & 1], st);
            cudaStreamWaitEvent(mma, ev_built[i & 1], 0);
        }
        return r;
    };
    auto contract = [&](long i) -> int {
        const I8Item it = item_at(i);
        const I8Emulator& e = *it.em;
        int8_t* Kp; double *mean_part, *qpart; int mc;
        parts(i, it, Kp, mean_part, qpart, mc);
        int r = trmm_sumsq_i8(e.Ls, e.inv_scale, e.Np, Kp, mc, e.nslices, e.bits, e.theta + D, qpart, sched, overlap ? mma : st);
        if (r == 0 && overlap) cudaEventRecord(ev_mma[i & 1], mma);
        return r;
    };
    auto finalize = [&](long i) -> int {
        const I8Item it = item_at(i);
        const I8Emulator& e = *it.em;
        int8_t* Kp; double *mean_part, *qpart; int mc;
        parts(i, it, Kp, mean_part, qpart, mc);
        if (overlap) cudaStreamWaitEvent(st, ev_mma[i & 1], 0);
        predict_finalize_kernel<<<(int)((it.mv + 255) / 256), 256, 0, st>>>(
            it.xs, it.mv, D, e.xa, e.xr, e.theta, mean_part, qpart, e.Np / TILE, mc, it.prior, e.feat_idx, e.n_feat, e.degree,
            e.weights, e.bias, e.y_mu, e.y_sigma, e.min_var, it.out_mean, it.out_std, i8_row_tiles(e.Np));
        return after_finalize(i);
    };
    // software pipeline: build(i+1) (and the enqueue of contract(i+1)) precede the wait for contract(i)
    rc = build(0);
    if (rc == 0) rc = contract(0);
    for (long i = 0; i < n_items && rc == 0; ++i) {
        if (i + 1 < n_items) {
            rc = build(i + 1);
            if (rc == 0) rc = contract(i + 1);      // on `mma` it queues behind contract(i); serial mode: after build(i+1) on `st`
            if (rc) break;
        }
        rc = finalize(i);
    }
    if (overlap) {
        if (rc) {                                   // never leave `mma` work un-joined
            cudaEvent_t j = nullptr;
            if (cudaEventCreateWithFlags(&j, cudaEventDisableTiming) == cudaSuccess) {
                cudaEventRecord(j, mma);
                cudaStreamWaitEvent(st, j, 0);
                cudaEventDestroy(j);
            }
        }
        for (int i = 0; i < 2; ++i) { cudaEventDestroy(ev_built[i]); cudaEventDestroy(ev_mma[i]); }
    }
    if (rc) return rc;
    GPX_CHECK_LAUNCH();
    return 0;
}

// workspace: two item buffers, each [K* slice planes [S][mc][Np] int8 | mean_part [nb][mc] | qpart [nt][mc]]
size_t predict_i8_workspace_bytes(int Np, int Mc, int nslices) { return 256 + 2 * (i8_item_bytes(Np, Mc, nslices) + 256); }
size_t predict_i8_planes_offset(int Np, int Mc, int nslices, int buffer) { return 256 + (size_t)buffer * (i8_item_bytes(Np, Mc, nslices) + 256); }

// As predict_meanvar; the variance contraction runs on the integer tensor cores from `nslices` slices of `bits` bits of L^-1
// (Ls, inv_scale: gpx_slice_i8 of S, once per factorisation) and of each K* chunk (written directly as slices by the builder).
// K*, the posterior mean and the final reduction are evaluated exactly like the FP64 path's, so the mean is bit-identical
// to predict_meanvar.
int predict_meanvar_i8(const double* Xs, long M, const double* X, int N, int D, const double* theta, const double* xa,
                       const double* xr, int kind, const int8_t* Ls, const double* inv_scale, int nslices, int bits, int Np,
                       const double* alpha, const double* prior_mean, const int* feat_idx, int n_feat, int degree,
                       const double* weights, const double* bias, double y_mu, double y_sigma, double min_var,
                       double* out_mean, double* out_std, void* workspace, size_t workspace_bytes, int Mc, cudaStream_t st,
                       cudaStream_t mma) {
    if (Mc <= 0 || Mc % TILE != 0) return -25;
    if (Np % TILE != 0 || Np < N) return -11;
    if (D > MAXD || D < 1) return -5;
    if (Np > i8_max_np(bits)) return -13;
    if (workspace_bytes < predict_i8_workspace_bytes(Np, Mc, nslices)) return -23;
    const I8Emulator em = {X, theta, xa, xr, alpha, Ls, inv_scale, feat_idx, weights, bias, y_mu, y_sigma, min_var,
                           N, Np, kind, nslices, bits, n_feat, degree, 0};
    const long n_chunks = (M + Mc - 1) / Mc;
    auto item_at = [&](long c) {
        const long m0 = c * Mc;
        I8Item it = {&em, Xs + m0 * D, (M - m0 < Mc) ? (M - m0) : (long)Mc, prior_mean ? prior_mean + m0 : nullptr, out_mean + m0,
                     out_std + m0};
        return it;
    };
    return run_i8_pipeline(n_chunks, item_at, [](long) { return 0; }, D, i8_item_bytes(Np, Mc, nslices) + 256, workspace, st, mma);
}

// ---- multi-emulator history-matching wave -------------------------------------------------------
// Replaces the body of Wave.compute_impl (GPErks/perks/history_matching.py:43-64): for every chunk of test points all E emulators
// predict from the same device-resident X* rows -- the (chunk, emulator) items run through ONE chunk pipeline, so the K* builder of
// emulator e+1 overlaps the tensor-core kernel of emulator e -- into chunk-local [E][Mc] buffers, and the implausibility epilogue
// (maxno-th largest over the emulators) turns them into I / PV: per-emulator predictions never leave the chip unless asked for.
size_t wave_i8_workspace_bytes(const I8Emulator* em, int E, int Mc) {
    size_t item = 0;
    for (int e = 0; e < E; ++e) {
        const size_t b = i8_item_bytes(em[e].Np, Mc, em[e].nslices) + 256;
        item = b > item ? b : item;
    }
    return 256 + 2 * item + (size_t)2 * E * Mc * 8;
}

int wave_i8(const I8Emulator* em, int E, int D, const double* Xs, long M, const double* z, const double* v, int maxno,
            double* I, double* PV, double* mean_out, double* std_out, void* workspace, size_t workspace_bytes, int Mc,
            cudaStream_t st, cudaStream_t mma) {
    if (E < 1 || E > 64) return -2;
    if (Mc <= 0 || Mc % TILE != 0) return -15;
    if (D > MAXD || D < 1) return -3;
    for (int e = 0; e < E; ++e) {
        if (em[e].Np % TILE || em[e].Np < em[e].N || em[e].Np > i8_max_np(em[e].bits)) return -1;
        if (em[e].nslices < 5 || em[e].nslices > 6) return -1;
    }
    if (workspace_bytes < wave_i8_workspace_bytes(em, E, Mc)) return -13;
    if (M == 0) return 0;
    size_t item = 0;
    for (int e = 0; e < E; ++e) {
        const size_t b = i8_item_bytes(em[e].Np, Mc, em[e].nslices) + 256;
        item = b > item ? b : item;
    }
    double* cmean = reinterpret_cast<double*>(reinterpret_cast<unsigned char*>(workspace) + 256 + 2 * item);      // [E][Mc] chunk-local
    double* cstd = cmean + (size_t)E * Mc;
    const long n_chunks = (M + Mc - 1) / Mc;
    const bool full = mean_out != nullptr && std_out != nullptr;
    auto item_at = [&](long i) {
        const long c = i / E;
        const int e = (int)(i - c * E);
        const long m0 = c * Mc;
        I8Item it = {em + e, Xs + m0 * D, (M - m0 < Mc) ? (M - m0) : (long)Mc, nullptr,
                     full ? mean_out + (size_t)e * M + m0 : cmean + (size_t)e * Mc, full ? std_out + (size_t)e * M + m0 : cstd + (size_t)e * Mc};
        return it;
    };
    auto after = [&](long i) -> int {
        const long c = i / E;
        if ((int)(i - c * E) != E - 1) return 0;
        const long m0 = c * Mc;
        const long mv = (M - m0 < Mc) ? (M - m0) : (long)Mc;
        return full ? implausibility(mean_out + m0, std_out + m0, mv, M, E, z, v, maxno, I + m0, PV + m0, st)
                    : implausibility(cmean, cstd, mv, Mc, E, z, v, maxno, I + m0, PV + m0, st);
    };
    return run_i8_pipeline(n_chunks * E, item_at, after, D, item, workspace, st, mma);
}

static void set_pred_attrs() {
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && g_pred_attr_dev[dev]) return;
    cudaFuncSetAttribute(trmm_sumsq_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G_SMEM_BYTES);
    cudaFuncSetAttribute(trmm_sumsq_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)T_SMEM_BYTES);
    if (dev >= 0 && dev < 64) g_pred_attr_dev[dev] = true;
}

}  // namespace gpx
