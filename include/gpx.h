/* gpx.h -- C ABI of libgpx.so, the sm_100a exact-GP kernel library behind gperks_b200.
 *
 * The reference (stelong/GPErks) has no FFI: its exact-GP arithmetic is delegated to
 * gpytorch/linear_operator Python calls.  Each entry point below names the reference call site whose
 * arithmetic it replaces (paths relative to the GPErks repository root).
 *
 * Conventions
 *   - plain C, raw DEVICE pointers (e.g. torch.Tensor.data_ptr()), FP64, row-major, no torch types;
 *   - `stream` is a cudaStream_t passed as void*; every call only enqueues work on it: no internal
 *     allocation, no host synchronisation, no static device state => re-entrant across streams;
 *   - factor matrices use the padded leading dimension Np = gpx_padded(N) (multiple of 128); the padding
 *     block of K_hat is the identity;
 *   - theta (device, D+2 doubles): [0,D) = 1/lengthscale_d, [D] = outputscale, [D+1] = noise;
 *   - return 0 = enqueued, <0 = invalid argument / launch error.  Numerical status (LAPACK-style info:
 *     1-based index of the first non-positive pivot) is written to a device int, never thrown.
 */
#ifndef GPX_H
#define GPX_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

enum gpx_kind { GPX_RBF = 0, GPX_MATERN12 = 1, GPX_MATERN32 = 2, GPX_MATERN52 = 3 };

int gpx_version(void);
int gpx_padded(int n);
int gpx_max_dim(void);

/* K_hat = s*k(X,X) + noise*I, lower 128x128 tiles of K[Np][Np] (identity padding).
 * replaces: ScaleKernel(RBFKernel|MaternKernel).forward + GaussianLikelihood noise add,
 *           GPErks/gp/model.py:18-21 and GPErks/train/emulator.py:326 (model(X_train) in train mode). */
int gpx_kmat_sym(const double* X, int N, int D, const double* theta, int kind, double* K, int Np, void* stream);

/* Ks[m][j] = s*k(x*_m, X_j) for a chunk of Mc (multiple of 128) test rows, ld = Np, plus
 * mean_part[jb][m] = sum_{j in 128-tile jb} Ks[m][j]*alpha[j].  xa/xr (nullable): unit-cube scaler
 * min and range applied to raw inputs ((x-a)/(b-a), GPErks/gp/data/data_scaler.py:43-45).
 * replaces: covar_module(X_new, X_train) + K* @ mean_cache, GPErks/train/emulator.py:352-357. */
int gpx_kmat_cross(const double* Xs, long m_valid, const double* X, int N, int D, const double* theta,
                   const double* xa, const double* xr, int kind, const double* alpha, double* Ks, int Np,
                   double* mean_part, int Mc, void* stream);

/* In-place blocked Cholesky of K (lower) -> L; writes the inverted diagonal blocks into S and DT,
 * per-block sum(log L_ii) into logdet_blk[Np/128], and info (DT is also used as scratch while the call runs; LAPACK convention: 0, or the 1-based order of
 * the first leading minor that is not positive definite; the factor is then undefined).
 * Stream semantics: all work is ordered after `stream` and joined back into it before the call returns, but it
 * runs on two per-device helper streams (a high-priority one for the panel chain, a low-priority one for the
 * bulk trailing updates) created on first use; concurrent callers are serialised while they enqueue.
 * No host synchronisation.
 * replaces: psd_safe_cholesky / torch.linalg.cholesky_ex inside ExactMarginalLogLikelihood and the
 *           prediction strategy, GPErks/train/emulator.py:326 and :356. */
int gpx_potrf(double* K, int Np, int N, double* S, double* DT, double* logdet_blk, int* info, void* stream);

/* Completes S to L^-1 (lower) with L^-T mirrored in the upper off-diagonal blocks. W: Np*Np scratch. */
int gpx_trtri(const double* L, double* S, const double* DT, double* W, int Np, void* stream);

/* Kinv = L^-T L^-1 (full symmetric).   replaces: autograd through cholesky (cholesky_backward). */
int gpx_lauum(const double* S, const double* DT, double* Kinv, int Np, void* stream);

/* alpha = K_hat^-1 r = L^-T (L^-1 r); r zero-padded to Np; u: scratch[Np].
 * replaces: cholesky_solve for the mean cache / the MLL quadratic form. */
int gpx_solve_alpha(const double* S, const double* DT, int Np, const double* r, double* u, double* alpha,
                    void* stream);

/* out[0] = -mll = [r.alpha/2 + sum log L_ii + N/2 log 2pi]/N, out[1] = r.alpha, out[2] = sum log L_ii.
 * replaces: ExactMarginalLogLikelihood.forward (divided by N), GPErks/train/emulator.py:326,332. */
int gpx_mll_value(const double* r, const double* alpha, int Np, int N, const double* logdet_blk, double* out,
                  void* stream);

/* grad[0..D) = d(-mll)/d lengthscale_d, grad[D] = d/d outputscale, grad[D+1] = d/d noise
 * (w.r.t. the constrained values).  partial: scratch (Np/128)^2*(D+2) doubles.
 * replaces: loss.backward() through the kernel + Cholesky graph, GPErks/train/emulator.py:327. */
int gpx_mll_grad(const double* X, int N, int D, const double* theta, int kind, const double* Kinv, int Np,
                 const double* alpha, double* partial, double* grad, void* stream);

/* Variance contraction alone: qpart[ib][m] = sum_{i in 128-row block ib} (L^-1 Ks^T)[i][m]^2 for a chunk
 * Ks[Mc][Np] (the dominant kernel of gpx_predict_meanvar; exposed for timing and parity tests).
 * replaces: the K* @ R root product of fast_pred_var, GPErks/train/emulator.py:354-358. */
int gpx_trmm_sumsq(const double* S, int Np, const double* Ks, int Mc, double* qpart, void* stream);

/* Posterior mean and noisy std for M device-resident raw points, un-standardised with (y_mu, y_sigma).
 * prior mean: prior_mean[M] if non-null, else sum_p weights[p]*prod_q x[feat_idx[p*degree+q]] + bias
 * on the unit-cube inputs.  workspace >= gpx_predict_workspace_bytes(Np, Mc).
 * replaces: GPEmulator.predict, GPErks/train/emulator.py:348-363. */
size_t gpx_predict_workspace_bytes(int Np, int Mc);
int gpx_predict_meanvar(const double* Xs, long M, const double* X, int N, int D, const double* theta,
                        const double* xa, const double* xr, int kind, const double* S, int Np,
                        const double* alpha, const double* prior_mean, const int* feat_idx, int n_feat,
                        int degree, const double* weights, const double* bias, double y_mu, double y_sigma,
                        double min_var, double* out_mean, double* out_std, void* workspace,
                        size_t workspace_bytes, int Mc, void* stream);

/* Dense predictive covariance for test-set sized M (with_covar, sample, validation loss):
 * gpx_trmm_store: Vt[m][i] = (K* L^-T)[m][i] for Ks[Mp][Np] -> Vt[Mp][Np];
 * gpx_syrk_sub:   C[Mp][Mp] -= Vt Vt^T  (C holds K** on entry, the predictive covariance on exit).
 * replaces: predictions.covariance_matrix / .sample / the eval-mode mll, GPErks/train/emulator.py:331-333,359-360,394. */
int gpx_trmm_store(const double* S, int Np, const double* Ks, int Mp, double* Vt, void* stream);
int gpx_syrk_sub(const double* Vt, int Mp, int Np, double* C, void* stream);

/* ---- TF32 mode of the predictive variance (north_star tolerance: rel 1e-3): 3xTF32 on tcgen05 / TMEM ----
 * gpx_split_tf32: (hi, lo) float split of n FP64 values, x ~= hi + lo, both exactly representable in TF32

 * (run once per factorisation on S = L^-1; ld = Np makes it zero the mirrored L^-T blocks above the diagonal, ld = 0
 * splits a flat array).  gpx_trmm_sumsq_tf32: qpart[it][m] = sum over the 256 rows of tile
 * `it` of (K* L^-T)[m][i]^2, it < gpx_tf32_row_tiles(Np); Mc must be a multiple of 256 (CTA pairs share an L^-1 tile
 * through TMA multicast).  gpx_predict_meanvar_tf32: as gpx_predict_meanvar with
 * the variance contraction on the tensor cores; kernel matrix and posterior mean remain FP64.
 * kchunk: K elements accumulated in one TMEM accumulator before it is drained (round-to-nearest) into the running
 * total (multiple of 32, <= 0 -> 32): the tensor core's FP32 accumulation truncates, so the error of std grows
 * ~linearly with kchunk (7.5e-5 at 32, 2.6e-4 at 128, 5.5e-3 unchunked; N=8192).  Two accumulators ping-pong, so the
 * drain overlaps the next chunk's MMAs and short chunks cost nothing.
 * gpx_predict_meanvar_tf32 only: kchunk >= 0 evaluates K* in FP32 (entries ~3e-7 relative); kchunk < 0 evaluates it in
 * FP64 before the split (posterior mean bit-identical to gpx_predict_meanvar) and uses |kchunk|.
 * replaces: the same call sites as gpx_predict_meanvar (GPErks/train/emulator.py:348-363). */
int gpx_split_tf32(const double* src, long n, float* hi, float* lo, int ld, void* stream);
int gpx_tf32_row_tiles(int Np);
int gpx_trmm_sumsq_tf32(const float* Shi, const float* Slo, int Np, const float* Khi, const float* Klo, int Mc,
                        double* qpart, int kchunk, void* stream);
size_t gpx_predict_tf32_workspace_bytes(int Np, int Mc);
int gpx_predict_meanvar_tf32(const double* Xs, long M, const double* X, int N, int D, const double* theta,
                             const double* xa, const double* xr, int kind, const float* Shi, const float* Slo, int Np,
                             const double* alpha, const double* prior_mean, const int* feat_idx, int n_feat,
                             int degree, const double* weights, const double* bias, double y_mu, double y_sigma,
                             double min_var, double* out_mean, double* out_std, void* workspace,
                             size_t workspace_bytes, int Mc, int kchunk, void* stream);

/* ---- sliced-integer mode of the predictive variance: FP64-grade results from the integer tensor cores -------------
 * (tcgen05 kind::i8, exact int32 accumulation in TMEM; gperks_b200/csrc/gpx_i8.cu).  Operands are scaled by powers of
 * two and cut into `nslices` (5 or 6) int8 digits of `bits` bits; the products of slices a+b < nslices are integer GEMMs,
 * combined in FP64.
 *   bits = 7: signed 7-bit slices (|digit| <= 64, repeated round-to-nearest); N <= 32768.  std agrees with gpx_predict_meanvar to
 *             ~5e-8 relative with 5 slices (5e-7 at noise 1e-4) and ~5e-9 with 6.
 *   bits = 8: balanced radix-256 digits in [-128, 127] of the operand rounded ONCE to 8*nslices bits.  The int32 sums are exact
 *             for ANY digits up to N_pad = 16384 (= gpx_i8_safe_n(8)); from there to 32768 (= gpx_i8_max_n(8)) they are exact when
 *             128 * gpx_i8_digit_bound(planes of L^-1) < 2^31 -- the caller checks that once per factorisation (the Python engine
 *             does, and falls back to 7-bit digits otherwise); the kernels do not.
 *             Five slices (40 bits, 15 slice products) stand in for six 7-bit ones (42 bits, 21 products); six carry 48 bits.
 * The posterior mean is bit-identical to gpx_predict_meanvar in every case (K* and the mean are evaluated in FP64).
 * Plane layout: planes[j][row][p(col)] (int8, row pitch = cols, plane pitch = plane_rows*cols) with p permuting the columns
 * inside every 128-block, p(c) = (c & ~127) | (c % 16) * 8 | (c % 128) / 16 -- identical on both operands, so every inner
 * product is unchanged; it lets the K* builder store 8 consecutive bytes per slice straight from registers.
 * gpx_slice_i8: planes <- slices of src[rows][cols] (leading dimension ld_src, cols % 128 == 0).  tri != 0: src is the mirrored
 *   inverse S of gpx_trtri; entries right of the diagonal are sliced as zero (columns beyond the row's 128-block are left
 *   untouched: never read) and every row gets its own scale, inv_scale[row] = 1/scale.  tri == 0: one scale for the whole matrix
 *   from the device scalar *amax_ptr >= max|src| (K* <= outputscale).  Run once per factorisation on S.
 * gpx_trmm_sumsq_i8: qpart[it][m] = sum over the 64 rows of tile `it` of (K* L^-T)[m][i]^2, it < gpx_i8_row_tiles(Np);
 *   Mc multiple of 128, Np <= gpx_i8_max_n(bits) (int32 accumulators).  sched: two device ints, zero before the first launch that
 *   uses them, from which the persistent kernel's clusters draw their work items in order (it rewinds them when it finishes; do not
 *   share them between launches that may overlap); NULL = static work distribution.
 * gpx_predict_meanvar_i8: as gpx_predict_meanvar with the variance contraction on the integer tensor cores.  Chunks are
 *   double-buffered in the workspace; with a non-null `mma_stream` (a second stream of the caller, ideally of higher priority)
 *   the tensor-core kernel of chunk c runs there while the K* builder of chunk c+1 runs on `stream` -- the two kernels are
 *   co-resident on every SM.  All results are ordered on `stream`; `mma_stream` is joined before the call's last kernel.
 *   The call creates and destroys four cudaEvents (no other state).
 * replaces: the same call sites as gpx_predict_meanvar (GPErks/train/emulator.py:348-363). */
int gpx_slice_i8(const double* src, int rows, int cols, long ld_src, signed char* planes, long plane_rows, int nslices,
                 int bits, int tri, double* inv_scale, const double* amax_ptr, void* stream);
int gpx_i8_row_tiles(int Np);
int gpx_i8_max_n(int bits);
int gpx_i8_safe_n(int bits);
/* out[0] (device) = max over the rows i of L^-1 of sum_{planes} sum_k |digit|: bounds every int32 accumulator by 128 * out[0] */
int gpx_i8_digit_bound(const signed char* planes, int nslices, int Np, unsigned long long* out, void* stream);
int gpx_i8_clusters(void);   /* diagnostics: resident 2-CTA clusters of the persistent tensor-core kernel configured last (0 = none yet) */
/* gpx_kmat_cross_i8: gpx_kmat_cross and gpx_slice_i8 (tri = 0, amax = outputscale) in one pass: the K* chunk is evaluated in
 * FP64 and written only as its int8 slices (same bytes, same mean_part as the two calls). */
int gpx_kmat_cross_i8(const double* Xs, long m_valid, const double* X, int N, int D, const double* theta, const double* xa,
                      const double* xr, int kind, const double* alpha, signed char* planes, long plane_rows, int nslices,
                      int bits, int Np, double* mean_part, int Mc, void* stream);
int gpx_trmm_sumsq_i8(const signed char* Ls, const double* inv_scale, int Np, const signed char* Kp, int Mc, int nslices,
                      int bits, const double* amax_ptr, double* qpart, int* sched, void* stream);
/* gpx_lauum_i8: gpx_lauum on the integer tensor cores.  The columns of L^-1 (rows of L^-T: DT inside the diagonal blocks, the
 * mirrored blocks of S beside them) are sliced with one power-of-two scale per column into the workspace, and
 * Kinv = L^-T L^-1 is formed by the same kind::i8 kernel (store epilogue) for the 128-tiles on and below the diagonal -- the
 * part gpx_mll_grad reads; tiles above the diagonal are NOT written.  Entries agree with gpx_lauum to ~1e-10 of the scale
 * of their row and column (6 slices).  Np <= 32768.
 * replaces: the K_hat^-1 of loss.backward(), GPErks/train/emulator.py:327. */
/* gpx_trtri_i8: gpx_trtri with the merge levels whose segments have >= 1024 rows (the top log2(Np/1024) levels, ~85 % of the
 * FLOPs at Np = 8192) on the integer tensor cores: per level and segment pair the four operands (L[right,left], the columns of
 * A^-1, the rows of C^-1, T^T) are sliced with per-row scales into the workspace and the two products run on a kind::i8 NT GEMM
 * with a store epilogue (negate, transposed mirror).  Entries of L^-1 agree with gpx_trtri to ~1e-11 of their row/column scale.
 * bits = 7: Np <= 32768; bits = 8 (balanced radix-256 digits, 48 bits per operand, ~64 x finer): Np <= 16384.
 * replaces: the same call sites as gpx_trtri. */
/* gpx_potrf_i8: gpx_potrf for training loops.  Same schedule, same outputs; the trailing updates of the two-level schedule (once
 * per outer block of 4 / 6 / 8 panels: the look-ahead columns on the critical-path stream, everything right of them on the helper
 * stream -- ~70 % of the FLOPs at Np = 8192, ~90 % at 26240) run on the integer tensor cores: the finished panel rows are sliced
 * into six 8-bit digits with one power-of-two scale per row (alternating plane buffers in the workspace) and
 * K[i][j] -= sum_k L[i][k] L[j][k] is formed by the kind::i8 NT GEMM kernel with an accumulate epilogue; int32 sums are exact
 * (at most 1024 terms), so every update carries ~2^-46 of sqrt(K_ii K_jj), against ~2^-52 sqrt(k) for FP64 DMMA.  Panels, diagonal
 * blocks, the K = 128 updates inside an outer block and the chain-bound tail (last 24 block columns) are the FP64 kernels of
 * gpx_potrf.  L within ~4e-12 of gpx_potrf's at Np = 8192.  Np > 3072 (below, it is gpx_potrf).
 * replaces: the same call sites as gpx_potrf, inside GPEmulator.train / train_auto loops only. */
size_t gpx_potrf_i8_workspace_bytes(int Np);
int gpx_potrf_i8(double* K, int Np, int N, double* S, double* DT, double* logdet_blk, int* info, void* workspace,
                 size_t workspace_bytes, void* stream);
size_t gpx_trtri_i8_workspace_bytes(int Np);
int gpx_trtri_i8(const double* L, double* S, const double* DT, double* W, int Np, int bits, void* workspace, size_t workspace_bytes,
                 void* stream);
/* gpx_predict_covar_i8: gpx_trmm_store + gpx_syrk_sub on the integer tensor cores, for the dense predictive covariance that the
 * validation loss of every training epoch needs (six digits per operand: 8 bit for Np <= 16384, 7 bit up to 32768).  Ks[Mp][Np]
 * (gpx_kmat_cross) and C[Mp][Mp] = K** on entry; on exit Vt[Mp][Np] = K* L^-T and the 128-tiles of C on or below the diagonal
 * hold K** - Vt Vt^T (tiles above it are untouched).  W is an Np x Np scratch matrix (receives V), workspace as for gpx_trtri_i8.
 * Mp <= Np.  Entries of Vt within ~2^-46 (8 bit) / 2^-40 (7 bit) of their row and column scales.
 * replaces: the same call sites as gpx_trmm_store / gpx_syrk_sub, inside GPEmulator.train loops only. */
int gpx_predict_covar_i8(const double* S, int Np, const double* Ks, int Mp, double* Vt, double* C, double* W, int bits, void* workspace,
                         size_t workspace_bytes, void* stream);
size_t gpx_lauum_i8_workspace_bytes(int Np, int nslices);
int gpx_lauum_i8(const double* S, const double* DT, double* Kinv, int Np, int nslices, void* workspace, size_t workspace_bytes,
                 void* stream);
size_t gpx_predict_i8_workspace_bytes(int Np, int Mc, int nslices);
/* byte offset inside the workspace of the K* slice planes of chunk buffer 0 / 1 (what the last chunks left there; timing tools) */
size_t gpx_predict_i8_planes_offset(int Np, int Mc, int nslices, int buffer);
int gpx_predict_meanvar_i8(const double* Xs, long M, const double* X, int N, int D, const double* theta, const double* xa,
                           const double* xr, int kind, const signed char* Ls, const double* inv_scale, int nslices, int bits,
                           int Np, const double* alpha, const double* prior_mean, const int* feat_idx, int n_feat, int degree,
                           const double* weights, const double* bias, double y_mu, double y_sigma, double min_var,
                           double* out_mean, double* out_std, void* workspace, size_t workspace_bytes, int Mc,
                           void* stream, void* mma_stream);

/* History-matching epilogue over E emulators: I[m] = maxno-th largest_e sqrt((mean[e][m]-z[e])^2/(std[e][m]^2+v[e])),
 * PV[m] = maxno-th largest_e std[e][m]^2/v[e]; inputs emulator-major [E][M]; 1 <= maxno <= min(E, 8).
 * replaces: the per-point loop with two np.sort in Wave.compute_impl, GPErks/perks/history_matching.py:57-62. */
int gpx_implausibility(const double* mean, const double* std_, long M, int E, const double* z, const double* v,
                       int maxno, double* I, double* PV, void* stream);

/* ---- multi-emulator history-matching wave ---------------------------------------------------------
 * gpx_i8_emulator: the device pointers of one factorised emulator on a sliced-integer mode (a HOST struct; every pointer inside is
 *   a device pointer): X[N][D] scaled training inputs, theta[D+2], xa/xr[D] unit-cube scaler of the raw test inputs (nullable
 *   pair), alpha[Np], Ls / inv_scale: gpx_slice_i8(tri = 1) of the inverse factor, polynomial prior mean (feat_idx[n_feat][degree],
 *   weights[n_feat], bias[1]; nullable), output un-standardisation (y_mu, y_sigma), variance floor.
 * gpx_wave_i8: for every chunk of Mc rows of Xs[M][D] all E emulators predict from the same device-resident rows -- the (chunk,
 *   emulator) items run through one two-stream chunk pipeline (see gpx_predict_meanvar_i8) -- and the implausibility epilogue
 *   writes I[M], PV[M].  mean_out / std_out (both or neither; emulator-major [E][M]) additionally keep every emulator's
 *   predictions; with NULL they only exist chunk by chunk inside the workspace.
 * gpx_compact_below: idx[0 .. *count) = the ascending indices m with values[m] < cutoff (deterministic, no atomics); idx must hold
 *   M entries.  replaces: np.where(I < cutoff) of Wave.find_regions, GPErks/perks/history_matching.py:66-77.
 * replaces: Wave.compute_impl, GPErks/perks/history_matching.py:43-64. */
typedef struct gpx_i8_emulator {
    const double* X; const double* theta; const double* xa; const double* xr; const double* alpha;
    const signed char* Ls; const double* inv_scale;
    const int* feat_idx; const double* weights; const double* bias;
    double y_mu, y_sigma, min_var;
    int N, Np, kind, nslices, bits, n_feat, degree, reserved;
} gpx_i8_emulator;
size_t gpx_wave_i8_workspace_bytes(const gpx_i8_emulator* em, int E, int Mc);
int gpx_wave_i8(const gpx_i8_emulator* em, int E, int D, const double* Xs, long M, const double* z, const double* v, int maxno,
                double* I, double* PV, double* mean_out, double* std_out, void* workspace, size_t workspace_bytes, int Mc,
                void* stream, void* mma_stream);
size_t gpx_compact_below_workspace_bytes(long M);
int gpx_compact_below(const double* values, long M, double cutoff, long long* idx, long long* count, void* workspace,
                      size_t workspace_bytes, void* stream);

/* ---- Saltelli / Sobol' GSA on the device ----------------------------------------------------------
 * gpx_predict_mean: posterior mean only (K* never materialised; bit-identical to the mean of gpx_predict_meanvar), for the
 *   indices of the emulator's posterior mean over Saltelli designs of 1e7 points.
 * gpx_saltelli_partial / gpx_saltelli_reduce: with mean (and std) of every design point resident -- layout A[n], B[n],
 *   AB[n][d] (row j = A_j with coordinate p taken from B_j) -- the sums scipy.stats.sobol_indices needs, per draw k:
 *   sums[k][0..4) = sum f_A, sum f_B, sum f_A^2, sum f_B^2;  sums[k][4 + 3p ..) = sum f_B D_p, sum D_p, sum D_p^2, D_p = f_AB^(p) - f_A,
 *   where f = mean + std * z, z ~ N(0,1) independent per point and draw (Philox4x32-10 keyed by seed; n_draws = 0: the noise-free
 *   sums of the mean, one row).  Rows are processed in chunks of gpx_saltelli_chunk_rows(); `partial` holds one [draws][4 + 3d]
 *   block per chunk of [j_begin, j_end) (j_begin chunk-aligned) and gpx_saltelli_reduce adds the blocks in chunk order, so the
 *   result does not depend on how chunks are dealt over GPUs.  The arrays are indexed by the GLOBAL design row (it is also the
 *   Philox counter): a rank that holds only rows [j_begin, j_end) passes base pointers shifted back by j_begin rows.
 * replaces: emulator.sample + the per-draw scipy estimators of SobolGSA.estimate_Sobol_indices_with_emulator,
 *   GPErks/perks/gsa.py:66-88. */
int gpx_predict_mean(const double* Xs, long M, const double* X, int N, int D, const double* theta, const double* xa, const double* xr,
                     int kind, const double* alpha, int Np, const double* prior_mean, const int* feat_idx, int n_feat, int degree,
                     const double* weights, const double* bias, double y_mu, double y_sigma, double* out_mean, void* stream);
int gpx_saltelli_chunk_rows(void);
int gpx_saltelli_partial(const double* mean_A, const double* mean_B, const double* mean_AB, const double* std_A, const double* std_B,
                         const double* std_AB, long n, int d, long j_begin, long j_end, int n_draws, unsigned long long seed,
                         double* partial, void* stream);
int gpx_saltelli_reduce(const double* partial, long n_chunks, int d, int n_draws, double* sums, void* stream);

/* Microbenchmarks for the roofline denominators: enqueue `iters` dependent-chain-free DMMA.8x8x4
 * (or DFMA) per warp on `blocks` CTAs of 256 threads; returns the FLOP count of the launch. */
double gpx_peak_dmma(int blocks, int iters, double* sink, void* stream);
/* kind::i8 MMA issue rate from resident shared memory (M=128, N=256, K=32, one CTA per SM when blocks = #SMs): returns the int8
 * operations issued; divide by the measured duration. */
double gpx_peak_i8mma(int blocks, int iters, void* stream);
double gpx_peak_dfma(int blocks, int iters, double* sink, void* stream);

#ifdef __cplusplus
}
#endif
#endif
