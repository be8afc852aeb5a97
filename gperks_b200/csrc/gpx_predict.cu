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
                                        double* __restrict__ out_mean, double* __restrict__ out_std) {
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
    for (int b = 0; b < nb; ++b) {
        km += mean_part[(long)b * Mc + m];
        q += qpart[(long)b * Mc + m];
    }
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
    trmm_sumsq_kernel<<<(Np / TILE) * (Mc / TILE), G_THREADS, G_SMEM_BYTES, st>>>(S, Np, Ks, Mc, qpart);
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
        trmm_sumsq_kernel<<<nb * (mc / TILE), G_THREADS, G_SMEM_BYTES, st>>>(S, Np, Ks, mc, qpart);
        predict_finalize_kernel<<<(int)((mv + 255) / 256), 256, 0, st>>>(
            xs, mv, D, xa, xr, theta, mean_part, qpart, nb, mc, prior_mean ? prior_mean + m0 : nullptr, feat_idx,
            n_feat, degree, weights, bias, y_mu, y_sigma, min_var, out_mean + m0, out_std + m0);
    }
    GPX_CHECK_LAUNCH();
    return 0;
}

static void set_pred_attrs() {
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && g_pred_attr_dev[dev]) return;
    cudaFuncSetAttribute(trmm_sumsq_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G_SMEM_BYTES);
    if (dev >= 0 && dev < 64) g_pred_attr_dev[dev] = true;
}

}  // namespace gpx
