// extern "C" surface of libgpx.so -- see include/gpx.h for the contract of every entry point.
#include "gpx_common.cuh"
#include "../../include/gpx.h"

namespace gpx {
int kmat_sym(const double*, int, int, const double*, int, double*, int, cudaStream_t);
int kmat_cross(const double*, long, const double*, int, int, const double*, const double*, const double*, int,
               const double*, double*, int, double*, int, cudaStream_t);
int mll_grad(const double*, int, int, const double*, int, const double*, int, const double*, double*, double*,
             cudaStream_t);
int potrf(double*, int, int, double*, double*, double*, int*, cudaStream_t, void*, size_t);
size_t potrf_i8_workspace_bytes(int);
int predict_covar_i8(const double*, int, const double*, int, double*, double*, double*, int, void*, size_t, cudaStream_t);
int trtri(const double*, double*, const double*, double*, int, cudaStream_t);
int lauum(const double*, const double*, double*, int, cudaStream_t);
int solve_alpha(const double*, const double*, int, const double*, double*, double*, cudaStream_t);
int mll_value(const double*, const double*, int, int, const double*, double*, cudaStream_t);
size_t predict_workspace_bytes(int, int);
int trmm_sumsq(const double*, int, const double*, int, double*, cudaStream_t);
int trmm_store(const double*, int, const double*, int, double*, cudaStream_t);
int syrk_sub(const double*, int, int, double*, cudaStream_t);
int split_tf32_matrix(const double*, long, float*, float*, int, cudaStream_t);
int trmm_sumsq_tf32(const float*, const float*, int, const float*, const float*, int, double*, int, cudaStream_t);
int tf32_row_tiles(int);
size_t predict_tf32_workspace_bytes(int, int);
int slice_i8_matrix(const double*, int, int, long, int8_t*, long, int, int, int, double*, const double*, cudaStream_t);
double peak_i8mma(int, int, cudaStream_t);
size_t trtri_i8_workspace_bytes(int, int);
int trtri_i8(const double*, double*, const double*, double*, int, int, void*, size_t, cudaStream_t);
size_t lauum_i8_workspace_bytes(int, int);
int lauum_i8(const double*, const double*, double*, int, int, void*, size_t, cudaStream_t);
int kmat_cross_i8(const double*, long, const double*, int, int, const double*, const double*, const double*, int, const double*,
                  int8_t*, long, int, int, int, double*, int, cudaStream_t);
int trmm_sumsq_i8(const int8_t*, const double*, int, const int8_t*, int, int, int, const double*, double*, int*, cudaStream_t);
int i8_row_tiles(int);
int i8_max_np(int);
int i8_safe_np(int);
int i8_digit_bound(const int8_t*, int, int, unsigned long long*, cudaStream_t);
int i8_clusters();
size_t predict_i8_workspace_bytes(int, int, int);
size_t predict_i8_planes_offset(int, int, int, int);
int predict_meanvar_i8(const double*, long, const double*, int, int, const double*, const double*, const double*, int,
                       const int8_t*, const double*, int, int, int, const double*, const double*, const int*, int, int,
                       const double*, const double*, double, double, double, double*, double*, void*, size_t, int,
                       cudaStream_t, cudaStream_t);
int predict_meanvar_tf32(const double*, long, const double*, int, int, const double*, const double*, const double*, int,
                         const float*, const float*, int, const double*, const double*, const int*, int, int,
                         const double*, const double*, double, double, double, double*, double*, void*, size_t, int,
                         int, cudaStream_t);
int implausibility(const double*, const double*, long, long, int, const double*, const double*, int, double*, double*,
                   cudaStream_t);
struct I8Emulator;
size_t wave_i8_workspace_bytes(const I8Emulator*, int, int);
int wave_i8(const I8Emulator*, int, int, const double*, long, const double*, const double*, int, double*, double*, double*, double*,
            void*, size_t, int, cudaStream_t, cudaStream_t);
int predict_mean(const double*, long, const double*, int, int, const double*, const double*, const double*, int, const double*, int,
                 const double*, const int*, int, int, const double*, const double*, double, double, double*, cudaStream_t);
int saltelli_chunk_rows();
int saltelli_partial(const double*, const double*, const double*, const double*, const double*, const double*, long, int, long, long,
                     int, unsigned long long, double*, cudaStream_t);
int saltelli_reduce(const double*, long, int, int, double*, cudaStream_t);
size_t compact_below_workspace_bytes(long);
int compact_below(const double*, long, double, long long*, long long*, void*, size_t, cudaStream_t);
int predict_meanvar(const double*, long, const double*, int, int, const double*, const double*, const double*, int,
                    const double*, int, const double*, const double*, const int*, int, int, const double*,
                    const double*, double, double, double, double*, double*, void*, size_t, int, cudaStream_t);

// ---- peak microbenchmarks -------------------------------------------------------------------
__global__ void __launch_bounds__(256) peak_dmma_kernel(int iters, double* sink) {
    double c[16][2];
#pragma unroll
    for (int i = 0; i < 16; ++i) { c[i][0] = 0.0; c[i][1] = 0.0; }
    double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) dmma884(c[i][0], c[i][1], a, b);
    }
    double t = 0.0;
#pragma unroll
    for (int i = 0; i < 16; ++i) t += c[i][0] + c[i][1];
    if (t == 12345.678) sink[0] = t;
}
__global__ void __launch_bounds__(256) peak_dfma_kernel(int iters, double* sink) {
    double c[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) c[i] = i;
    double a = 1.0 + threadIdx.x * 1e-9, b = 1e-9 * threadIdx.x;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) c[i] = fma(c[i], a, b);
    }
    double t = 0.0;
#pragma unroll
    for (int i = 0; i < 16; ++i) t += c[i];
    if (t == 12345.678) sink[0] = t;
}
}  // namespace gpx

using namespace gpx;
#define ST(s) reinterpret_cast<cudaStream_t>(s)

extern "C" {

int gpx_version(void) { return 203; }
int gpx_padded(int n) { return round_up(n < 1 ? 1 : n, TILE); }
int gpx_max_dim(void) { return MAXD; }

static int check_np(int N, int Np) { return (Np % TILE == 0 && Np >= N && N >= 1) ? 0 : -1; }

int gpx_kmat_sym(const double* X, int N, int D, const double* theta, int kind, double* K, int Np, void* stream) {
    if (!X || !theta || !K) return -1;
    if (check_np(N, Np)) return -7;
    if (D < 1 || D > MAXD) return -3;
    return kmat_sym(X, N, D, theta, kind, K, Np, ST(stream));
}

int gpx_kmat_cross(const double* Xs, long m_valid, const double* X, int N, int D, const double* theta,
                   const double* xa, const double* xr, int kind, const double* alpha, double* Ks, int Np,
                   double* mean_part, int Mc, void* stream) {
    if (!Xs || !X || !theta || !alpha || !Ks || !mean_part) return -1;
    if (check_np(N, Np)) return -12;
    if (D < 1 || D > MAXD) return -5;
    if (Mc <= 0 || Mc % TILE != 0 || m_valid > Mc) return -14;
    if ((xa == nullptr) != (xr == nullptr)) return -7;
    return kmat_cross(Xs, m_valid, X, N, D, theta, xa, xr, kind, alpha, Ks, Np, mean_part, Mc, ST(stream));
}

int gpx_potrf(double* K, int Np, int N, double* S, double* DT, double* logdet_blk, int* info, void* stream) {
    if (!K || !S || !DT || !logdet_blk || !info) return -1;
    if (check_np(N, Np)) return -2;
    return potrf(K, Np, N, S, DT, logdet_blk, info, ST(stream), nullptr, 0);
}

size_t gpx_potrf_i8_workspace_bytes(int Np) { return potrf_i8_workspace_bytes(Np); }

int gpx_potrf_i8(double* K, int Np, int N, double* S, double* DT, double* logdet_blk, int* info, void* workspace,
                 size_t workspace_bytes, void* stream) {
    if (!K || !S || !DT || !logdet_blk || !info || !workspace) return -1;
    if (check_np(N, Np)) return -2;
    if (workspace_bytes < potrf_i8_workspace_bytes(Np)) return -7;
    return potrf(K, Np, N, S, DT, logdet_blk, info, ST(stream), workspace, workspace_bytes);
}

int gpx_trtri(const double* L, double* S, const double* DT, double* W, int Np, void* stream) {
    if (!L || !S || !DT || !W) return -1;
    if (Np % TILE) return -5;
    return trtri(L, S, DT, W, Np, ST(stream));
}

int gpx_lauum(const double* S, const double* DT, double* Kinv, int Np, void* stream) {
    if (!S || !DT || !Kinv) return -1;
    if (Np % TILE) return -4;
    return lauum(S, DT, Kinv, Np, ST(stream));
}

int gpx_solve_alpha(const double* S, const double* DT, int Np, const double* r, double* u, double* alpha,
                    void* stream) {
    if (!S || !DT || !r || !u || !alpha) return -1;
    if (Np % TILE) return -3;
    return solve_alpha(S, DT, Np, r, u, alpha, ST(stream));
}

int gpx_mll_value(const double* r, const double* alpha, int Np, int N, const double* logdet_blk, double* out,
                  void* stream) {
    if (!r || !alpha || !logdet_blk || !out) return -1;
    if (check_np(N, Np)) return -3;
    return mll_value(r, alpha, Np, N, logdet_blk, out, ST(stream));
}

int gpx_mll_grad(const double* X, int N, int D, const double* theta, int kind, const double* Kinv, int Np,
                 const double* alpha, double* partial, double* grad, void* stream) {
    if (!X || !theta || !Kinv || !alpha || !partial || !grad) return -1;
    if (check_np(N, Np)) return -7;
    if (D < 1 || D > MAXD) return -3;
    return mll_grad(X, N, D, theta, kind, Kinv, Np, alpha, partial, grad, ST(stream));
}

int gpx_trmm_sumsq(const double* S, int Np, const double* Ks, int Mc, double* qpart, void* stream) {
    if (!S || !Ks || !qpart) return -1;
    if (Np % TILE || Mc % TILE || Np <= 0 || Mc <= 0) return -2;
    return trmm_sumsq(S, Np, Ks, Mc, qpart, ST(stream));
}

size_t gpx_predict_workspace_bytes(int Np, int Mc) { return predict_workspace_bytes(Np, Mc); }

int gpx_predict_meanvar(const double* Xs, long M, const double* X, int N, int D, const double* theta,
                        const double* xa, const double* xr, int kind, const double* S, int Np,
                        const double* alpha, const double* prior_mean, const int* feat_idx, int n_feat,
                        int degree, const double* weights, const double* bias, double y_mu, double y_sigma,
                        double min_var, double* out_mean, double* out_std, void* workspace,
                        size_t workspace_bytes, int Mc, void* stream) {
    if (!Xs || !X || !theta || !S || !alpha || !out_mean || !out_std || !workspace) return -1;
    if (M < 0) return -2;
    if (M == 0) return 0;
    if ((xa == nullptr) != (xr == nullptr)) return -7;
    if (!prior_mean && weights && n_feat > 0 && (!feat_idx || degree < 1)) return -14;
    return predict_meanvar(Xs, M, X, N, D, theta, xa, xr, kind, S, Np, alpha, prior_mean, feat_idx, n_feat, degree,
                           weights, bias, y_mu, y_sigma, min_var, out_mean, out_std, workspace, workspace_bytes, Mc,
                           ST(stream));
}

int gpx_trmm_store(const double* S, int Np, const double* Ks, int Mp, double* Vt, void* stream) {
    if (!S || !Ks || !Vt) return -1;
    if (Np % TILE || Mp % TILE || Np <= 0 || Mp <= 0) return -2;
    return trmm_store(S, Np, Ks, Mp, Vt, ST(stream));
}

int gpx_syrk_sub(const double* Vt, int Mp, int Np, double* C, void* stream) {
    if (!Vt || !C) return -1;
    if (Np % TILE || Mp % TILE || Np <= 0 || Mp <= 0) return -2;
    return syrk_sub(Vt, Mp, Np, C, ST(stream));
}

int gpx_split_tf32(const double* src, long n, float* hi, float* lo, int ld, void* stream) {
    if (!src || !hi || !lo) return -1;
    if (n < 0 || ld < 0) return -2;
    return split_tf32_matrix(src, n, hi, lo, ld, ST(stream));
}

int gpx_tf32_row_tiles(int Np) { return tf32_row_tiles(Np); }

int gpx_trmm_sumsq_tf32(const float* Shi, const float* Slo, int Np, const float* Khi, const float* Klo, int Mc,
                        double* qpart, int kchunk, void* stream) {
    if (!Shi || !Slo || !Khi || !Klo || !qpart) return -1;
    if (Np % TILE || Mc % (2 * TILE) || Np <= 0 || Mc <= 0) return -3;
    if (kchunk > 0 && kchunk % 32) return -8;
    return trmm_sumsq_tf32(Shi, Slo, Np, Khi, Klo, Mc, qpart, kchunk, ST(stream));
}

size_t gpx_predict_tf32_workspace_bytes(int Np, int Mc) { return predict_tf32_workspace_bytes(Np, Mc); }

int gpx_predict_meanvar_tf32(const double* Xs, long M, const double* X, int N, int D, const double* theta,
                             const double* xa, const double* xr, int kind, const float* Shi, const float* Slo, int Np,
                             const double* alpha, const double* prior_mean, const int* feat_idx, int n_feat,
                             int degree, const double* weights, const double* bias, double y_mu, double y_sigma,
                             double min_var, double* out_mean, double* out_std, void* workspace,
                             size_t workspace_bytes, int Mc, int kchunk, void* stream) {
    if (kchunk % 32) return -27;
    if (!Xs || !X || !theta || !Shi || !Slo || !alpha || !out_mean || !out_std || !workspace) return -1;
    if (M < 0) return -2;
    if (M == 0) return 0;
    if ((xa == nullptr) != (xr == nullptr)) return -7;
    if (!prior_mean && weights && n_feat > 0 && (!feat_idx || degree < 1)) return -14;
    return predict_meanvar_tf32(Xs, M, X, N, D, theta, xa, xr, kind, Shi, Slo, Np, alpha, prior_mean, feat_idx, n_feat,
                                degree, weights, bias, y_mu, y_sigma, min_var, out_mean, out_std, workspace,
                                workspace_bytes, Mc, kchunk, ST(stream));
}

int gpx_slice_i8(const double* src, int rows, int cols, long ld_src, signed char* planes, long plane_rows, int nslices,
                 int bits, int tri, double* inv_scale, const double* amax_ptr, void* stream) {
    if (!src || !planes) return -1;
    if (rows < 0 || cols <= 0 || cols % TILE || ld_src < cols || plane_rows < rows) return -2;
    if (nslices < 5 || nslices > 6) return -7;
    if (bits != 7 && bits != 8) return -8;
    if (tri ? !inv_scale : !amax_ptr) return -10;
    return slice_i8_matrix(src, rows, cols, ld_src, reinterpret_cast<int8_t*>(planes), plane_rows, nslices, bits, tri, inv_scale,
                           amax_ptr, ST(stream));
}

int gpx_i8_max_n(int bits) { return (bits == 7 || bits == 8) ? i8_max_np(bits) : 0; }
int gpx_i8_safe_n(int bits) { return i8_safe_np(bits); }
int gpx_i8_digit_bound(const signed char* planes, int nslices, int Np, unsigned long long* out, void* stream) {
    if (!planes || !out) return -1;
    return i8_digit_bound(reinterpret_cast<const int8_t*>(planes), nslices, Np, out, ST(stream));
}
int gpx_i8_clusters(void) { return i8_clusters(); }

int gpx_i8_row_tiles(int Np) { return i8_row_tiles(Np); }

int gpx_kmat_cross_i8(const double* Xs, long m_valid, const double* X, int N, int D, const double* theta, const double* xa,
                      const double* xr, int kind, const double* alpha, signed char* planes, long plane_rows, int nslices,
                      int bits, int Np, double* mean_part, int Mc, void* stream) {
    if (!Xs || !X || !theta || !alpha || !planes || !mean_part) return -1;
    if (D < 1 || D > MAXD) return -5;
    if (Np % TILE || Mc % TILE || Np < N || Mc <= 0 || plane_rows < Mc) return -3;
    if (nslices < 5 || nslices > 6) return -13;
    if (bits != 7 && bits != 8) return -14;
    if ((xa == nullptr) != (xr == nullptr)) return -7;
    return kmat_cross_i8(Xs, m_valid, X, N, D, theta, xa, xr, kind, alpha, reinterpret_cast<int8_t*>(planes), plane_rows,
                         nslices, bits, Np, mean_part, Mc, ST(stream));
}

int gpx_trmm_sumsq_i8(const signed char* Ls, const double* inv_scale, int Np, const signed char* Kp, int Mc, int nslices,
                      int bits, const double* amax_ptr, double* qpart, int* sched, void* stream) {
    if (!Ls || !inv_scale || !Kp || !amax_ptr || !qpart) return -1;
    if (Np % TILE || Mc % TILE || Np <= 0 || Mc <= 0) return -3;
    if (nslices < 5 || nslices > 6) return -6;
    if (bits != 7 && bits != 8) return -7;
    return trmm_sumsq_i8(reinterpret_cast<const int8_t*>(Ls), inv_scale, Np, reinterpret_cast<const int8_t*>(Kp), Mc,
                         nslices, bits, amax_ptr, qpart, sched, ST(stream));
}

size_t gpx_trtri_i8_workspace_bytes(int Np) { return trtri_i8_workspace_bytes(Np, 6); }

int gpx_trtri_i8(const double* L, double* S, const double* DT, double* W, int Np, int bits, void* workspace, size_t workspace_bytes,
                 void* stream) {
    if (!L || !S || !DT || !W || !workspace) return -1;
    if (Np <= 0 || Np % TILE) return -4;
    return trtri_i8(L, S, DT, W, Np, bits, workspace, workspace_bytes, ST(stream));
}

int gpx_predict_covar_i8(const double* S, int Np, const double* Ks, int Mp, double* Vt, double* C, double* W, int bits, void* workspace,
                         size_t workspace_bytes, void* stream) {
    if (!S || !Ks || !Vt || !C || !W || !workspace) return -1;
    return predict_covar_i8(S, Np, Ks, Mp, Vt, C, W, bits, workspace, workspace_bytes, ST(stream));
}

size_t gpx_lauum_i8_workspace_bytes(int Np, int nslices) { return lauum_i8_workspace_bytes(Np, nslices); }

int gpx_lauum_i8(const double* S, const double* DT, double* Kinv, int Np, int nslices, void* workspace, size_t workspace_bytes,
                 void* stream) {
    if (!S || !DT || !Kinv || !workspace) return -1;
    if (Np <= 0 || Np % TILE) return -4;
    return lauum_i8(S, DT, Kinv, Np, nslices, workspace, workspace_bytes, ST(stream));
}

size_t gpx_predict_i8_workspace_bytes(int Np, int Mc, int nslices) { return predict_i8_workspace_bytes(Np, Mc, nslices); }
size_t gpx_predict_i8_planes_offset(int Np, int Mc, int nslices, int buffer) {
    return predict_i8_planes_offset(Np, Mc, nslices, buffer & 1);
}

int gpx_predict_meanvar_i8(const double* Xs, long M, const double* X, int N, int D, const double* theta, const double* xa,
                           const double* xr, int kind, const signed char* Ls, const double* inv_scale, int nslices, int bits,
                           int Np, const double* alpha, const double* prior_mean, const int* feat_idx, int n_feat, int degree,
                           const double* weights, const double* bias, double y_mu, double y_sigma, double min_var,
                           double* out_mean, double* out_std, void* workspace, size_t workspace_bytes, int Mc,
                           void* stream, void* mma_stream) {
    if (!Xs || !X || !theta || !Ls || !inv_scale || !alpha || !out_mean || !out_std || !workspace) return -1;
    if (M < 0) return -2;
    if (nslices < 5 || nslices > 6) return -12;
    if (bits != 7 && bits != 8) return -13;
    if (Np > i8_max_np(bits)) return -14;
    if (M == 0) return 0;
    if ((xa == nullptr) != (xr == nullptr)) return -7;
    if (!prior_mean && weights && n_feat > 0 && (!feat_idx || degree < 1)) return -14;
    return predict_meanvar_i8(Xs, M, X, N, D, theta, xa, xr, kind, reinterpret_cast<const int8_t*>(Ls), inv_scale, nslices,
                              bits, Np, alpha, prior_mean, feat_idx, n_feat, degree, weights, bias, y_mu, y_sigma, min_var,
                              out_mean, out_std, workspace, workspace_bytes, Mc, ST(stream), ST(mma_stream));
}

int gpx_implausibility(const double* mean, const double* std_, long M, int E, const double* z, const double* v,
                       int maxno, double* I, double* PV, void* stream) {
    if (!mean || !std_ || !z || !v || !I || !PV) return -1;
    if (M < 0 || E < 1) return -3;
    return implausibility(mean, std_, M, M, E, z, v, maxno, I, PV, ST(stream));
}

size_t gpx_wave_i8_workspace_bytes(const gpx_i8_emulator* em, int E, int Mc) {
    if (!em || E < 1 || Mc <= 0) return 0;
    return wave_i8_workspace_bytes(reinterpret_cast<const I8Emulator*>(em), E, Mc);
}

int gpx_wave_i8(const gpx_i8_emulator* em, int E, int D, const double* Xs, long M, const double* z, const double* v, int maxno,
                double* I, double* PV, double* mean_out, double* std_out, void* workspace, size_t workspace_bytes, int Mc,
                void* stream, void* mma_stream) {
    if (!em || !Xs || !z || !v || !I || !PV || !workspace) return -1;
    if (M < 0) return -5;
    if ((mean_out == nullptr) != (std_out == nullptr)) return -11;
    for (int e = 0; e < E; ++e) {
        const gpx_i8_emulator& m = em[e];
        if (!m.X || !m.theta || !m.alpha || !m.Ls || !m.inv_scale) return -1;
        if ((m.xa == nullptr) != (m.xr == nullptr)) return -1;
        if (m.weights && m.n_feat > 0 && (!m.feat_idx || m.degree < 1)) return -1;
    }
    return wave_i8(reinterpret_cast<const I8Emulator*>(em), E, D, Xs, M, z, v, maxno, I, PV, mean_out, std_out, workspace,
                   workspace_bytes, Mc, ST(stream), ST(mma_stream));
}

int gpx_predict_mean(const double* Xs, long M, const double* X, int N, int D, const double* theta, const double* xa, const double* xr,
                     int kind, const double* alpha, int Np, const double* prior_mean, const int* feat_idx, int n_feat, int degree,
                     const double* weights, const double* bias, double y_mu, double y_sigma, double* out_mean, void* stream) {
    if (!Xs || !X || !theta || !alpha || !out_mean) return -1;
    if (M < 0) return -2;
    if (check_np(N, Np)) return -11;
    if (D < 1 || D > MAXD) return -5;
    if ((xa == nullptr) != (xr == nullptr)) return -7;
    if (!prior_mean && weights && n_feat > 0 && (!feat_idx || degree < 1)) return -14;
    return predict_mean(Xs, M, X, N, D, theta, xa, xr, kind, alpha, Np, prior_mean, feat_idx, n_feat, degree, weights, bias, y_mu,
                        y_sigma, out_mean, ST(stream));
}

int gpx_saltelli_chunk_rows(void) { return saltelli_chunk_rows(); }

int gpx_saltelli_partial(const double* mean_A, const double* mean_B, const double* mean_AB, const double* std_A, const double* std_B,
                         const double* std_AB, long n, int d, long j_begin, long j_end, int n_draws, unsigned long long seed,
                         double* partial, void* stream) {
    if (!mean_A || !mean_B || !mean_AB || !partial) return -1;
    if (n_draws > 0 && (!std_A || !std_B || !std_AB)) return -4;
    if (n_draws < 0) return -11;
    return saltelli_partial(mean_A, mean_B, mean_AB, std_A, std_B, std_AB, n, d, j_begin, j_end, n_draws, seed, partial, ST(stream));
}

int gpx_saltelli_reduce(const double* partial, long n_chunks, int d, int n_draws, double* sums, void* stream) {
    if (!partial || !sums) return -1;
    if (n_chunks < 1 || d < 1 || n_draws < 0) return -2;
    return saltelli_reduce(partial, n_chunks, d, n_draws, sums, ST(stream));
}

size_t gpx_compact_below_workspace_bytes(long M) { return compact_below_workspace_bytes(M < 0 ? 0 : M); }

int gpx_compact_below(const double* values, long M, double cutoff, long long* idx, long long* count, void* workspace,
                      size_t workspace_bytes, void* stream) {
    if (!values || !idx || !count || !workspace) return -1;
    return compact_below(values, M, cutoff, idx, count, workspace, workspace_bytes, ST(stream));
}

double gpx_peak_dmma(int blocks, int iters, double* sink, void* stream) {
    peak_dmma_kernel<<<blocks, 256, 0, ST(stream)>>>(iters, sink);
    // per warp per iter: 16 DMMA x (8*8*4 FMA) x 2 FLOP
    return (double)blocks * 8.0 * iters * 16.0 * 512.0;
}
double gpx_peak_i8mma(int blocks, int iters, void* stream) { return peak_i8mma(blocks, iters, ST(stream)); }
double gpx_peak_dfma(int blocks, int iters, double* sink, void* stream) {
    peak_dfma_kernel<<<blocks, 256, 0, ST(stream)>>>(iters, sink);
    return (double)blocks * 256.0 * iters * 16.0 * 2.0;
}

}  // extern "C"
