// Blocked Cholesky (with log-determinant and LAPACK-style info), recursive triangular inverse,
// K_hat^-1 = L^-T L^-1, and the triangular mat-vecs that give alpha = K_hat^-1 r.
//
// Replaces what gpytorch's exact path does with torch.linalg.cholesky_ex / cholesky_solve /
// autograd-through-cholesky (psd_safe_cholesky; call sites GPErks/train/emulator.py:326-327 for
// the MLL forward/backward and :356 for the prediction caches).
//
// Every dense product is arranged as an NT GEMM (both operands k-contiguous) on the shared FP64
// DMMA tile engine in gpx_common.cuh:
//   S  (Np x Np): lower part = L^-1, strictly-upper part of each diagonal 128-block = 0,
//                 upper off-diagonal blocks = mirror (L^-T), so rows of S serve as "columns" of L^-1.
//   DT (nb x 128 x 128): transposed diagonal blocks of L^-1 (upper triangular, zero below).
#include "gpx_common.cuh"

namespace gpx {

// ------------------------------------------------------------------------------------------------
// Diagonal block: Cholesky of a 128 x 128 SPD block + its triangular inverse, one CTA, 512 threads.
// Thread t: row/column r = t/4, k-partition h = t%4 (dot products split 4 ways, shuffle-combined).
// ------------------------------------------------------------------------------------------------
constexpr int DG_THREADS = 512;
constexpr int DG_LD = 132;   // 8 (mod 32) words per row: conflict-free 8-row x 4-k half-warp reads

__global__ void __launch_bounds__(DG_THREADS, 1)
potrf_diag_kernel(double* __restrict__ K, int Np, int blk, int N, double* __restrict__ S, double* __restrict__ DT,
                  double* __restrict__ logdet_blk, int* __restrict__ info) {
    extern __shared__ double a[];            // [128][DG_LD]  lower: L ; strictly upper: X^T (inverse)
    double* rinv = a + TILE * DG_LD;         // [128] 1/L_jj
    double* piv = rinv + TILE;               // [1]
    double* lsum = piv + 2;                  // [128] log L_jj
    const int tid = threadIdx.x;
    const long off = (long)blk * TILE;
    double* Ab = K + off * Np + off;
    for (int e = tid; e < TILE * TILE; e += DG_THREADS) {
        int r = e >> 7, c = e & 127;
        a[r * DG_LD + c] = Ab[(long)r * Np + c];
    }
    __syncthreads();
    const int r = tid >> 2, h = tid & 3;
    // ---- left-looking Cholesky, column by column
    for (int j = 0; j < TILE; ++j) {
        double s = 0.0;
        if (r >= j) {
            const double* ar = a + r * DG_LD;
            const double* aj = a + j * DG_LD;
            for (int k = h; k < j; k += 4) s = fma(ar[k], aj[k], s);
        }
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        double v = 0.0;
        if (r >= j) v = a[r * DG_LD + j] - s;
        if (r == j && h == 0) piv[0] = v;
        __syncthreads();
        const double d = piv[0];
        if (tid == 0) {
            if (!(d > 0.0) && *info == 0) *info = (int)off + j + 1;   // LAPACK: first non-positive pivot
        }
        const double sd = sqrt(d);
        if (h == 0 && r >= j) {
            if (r == j) { a[r * DG_LD + j] = sd; rinv[j] = 1.0 / sd; lsum[j] = log(sd); }
            else a[r * DG_LD + j] = v / sd;
        }
        __syncthreads();
    }
    // ---- triangular inverse X = L^-1 by forward substitution, one column c per 4-thread group.
    // X[i][c] (i > c) is stored transposed in the strictly-upper part: a[c][i]; X[c][c] = rinv[c].
    {
        const int c = r;
        double* xc = a + c * DG_LD;
        const double rc = rinv[c];
        const int cmin = (tid >> 5) * 8;   // smallest column of this warp: warp-uniform trip count
        for (int i = cmin + 1; i < TILE; ++i) {
            const bool active = i > c;
            const double* li = a + i * DG_LD;
            double s = 0.0;
            if (active) {
                if (h == 0) s = li[c] * rc;
                for (int k = c + 1 + h; k < i; k += 4) s = fma(li[k], xc[k], s);
            }
            s += __shfl_xor_sync(0xffffffffu, s, 1);
            s += __shfl_xor_sync(0xffffffffu, s, 2);
            if (active && h == 0) xc[i] = -s * rinv[i];
            __syncwarp();
        }
    }
    __syncthreads();
    // ---- write back: L (lower, zero above), D = L^-1 into S (zero above), D^T into DT
    double* Sb = S + off * Np + off;
    double* DTb = DT + (long)blk * TILE * TILE;
    for (int e = tid; e < TILE * TILE; e += DG_THREADS) {
        int i = e >> 7, c = e & 127;
        double l = (c <= i) ? a[i * DG_LD + c] : 0.0;
        double x = (c < i) ? a[c * DG_LD + i] : ((c == i) ? rinv[i] : 0.0);
        Ab[(long)i * Np + c] = l;
        Sb[(long)i * Np + c] = x;
    }
    for (int e = tid; e < TILE * TILE; e += DG_THREADS) {
        int c = e >> 7, i = e & 127;   // DT[c][i] = D[i][c]
        double x = (c < i) ? a[c * DG_LD + i] : ((c == i) ? rinv[i] : 0.0);
        DTb[c * TILE + i] = x;
    }
    if (tid < 32) {
        double t = 0.0;
        for (int j = tid; j < TILE; j += 32)
            if (off + j < N) t += lsum[j];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if (tid == 0) logdet_blk[blk] = t;
    }
}

// ------------------------------------------------------------------------------------------------
// Panel: L[i][p] = A[i][p] * D_p^T   for row tiles i > p.   grid.x = nb - p - 1
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(G_THREADS, 1) potrf_panel_kernel(double* __restrict__ K, int Np, int p,
                                                                   const double* __restrict__ S) {
    extern __shared__ double smem[];
    const int ti = p + 1 + blockIdx.x;
    double* Ct = K + (long)ti * TILE * Np + (long)p * TILE;
    Operand A = make_operand(Ct, Np);
    Operand B = make_operand(S + (long)p * TILE * Np + (long)p * TILE, Np);
    double acc[8][4][2];
    gemm_nt_mainloop(acc, A, B, 0, TILE, smem);
    gemm_foreach_pair(acc, [&](int row, int col, double v0, double v1) {
        *reinterpret_cast<double2*>(Ct + (long)row * Np + col) = make_double2(v0, v1);
    });
}

// Trailing update: A[i][j] -= L[i][p] L[j][p]^T for p < j <= i.  grid = (n, n), n = nb - p - 1.
__global__ void __launch_bounds__(G_THREADS, 1) potrf_trailing_kernel(double* __restrict__ K, int Np, int p) {
    if (blockIdx.x > blockIdx.y) return;
    extern __shared__ double smem[];
    const int ti = p + 1 + blockIdx.y, tj = p + 1 + blockIdx.x;
    Operand A = make_operand(K + (long)ti * TILE * Np + (long)p * TILE, Np);
    Operand B = make_operand(K + (long)tj * TILE * Np + (long)p * TILE, Np);
    double acc[8][4][2];
    gemm_nt_mainloop(acc, A, B, 0, TILE, smem);
    double* Ct = K + (long)ti * TILE * Np + (long)tj * TILE;
    gemm_foreach_pair(acc, [&](int row, int col, double v0, double v1) {
        double2* ptr = reinterpret_cast<double2*>(Ct + (long)row * Np + col);
        double2 c = *ptr;
        c.x -= v0; c.y -= v1;
        *ptr = c;
    });
}

// ------------------------------------------------------------------------------------------------
// Recursive inverse, merge step for segments of h blocks (pair q covers blocks [2hq, 2hq+2h)):
//   X21 = -C^-1 * B * A^-1,   A^-1 = S[left,left], C^-1 = S[right,right], B = L[right,left]
// step 1:  T = B * A^-1, stored transposed in W:  W[left j][right i] = T[i][j]
// step 2:  X21 = -(C^-1 * T) -> S[right i][left j], mirrored into S[left j][right i]
// grid = (h tiles j, <=h tiles i, pairs)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(G_THREADS, 1)
trtri_step1_kernel(const double* __restrict__ L, const double* __restrict__ S, const double* __restrict__ DT,
                   double* __restrict__ W, int Np, int nb, int h) {
    const int tj = blockIdx.x, ti = blockIdx.y, q = blockIdx.z;
    const int left = 2 * h * q, right = left + h;
    if (right + ti >= nb) return;
    extern __shared__ double smem[];
    const long p0 = (long)left * TILE;
    // A rows i (right segment), k over the left segment: L[right i][left k]
    Operand A = make_operand(L + ((long)(right + ti) * TILE) * Np + p0, Np);
    // B rows j (left segment): A^-1[k][j] = mirror S[left j][left k] for k-blocks > tj, DT for k-block == tj
    Operand B = make_operand_alt(S + ((long)(left + tj) * TILE) * Np + p0, Np, DT + (long)(left + tj) * TILE * TILE,
                                 TILE, tj * TILE, (tj + 1) * TILE);
    double acc[8][4][2];
    gemm_nt_mainloop(acc, A, B, tj * TILE, h * TILE, smem);
    double* Wt = W + ((long)(left + tj) * TILE) * Np + (long)(right + ti) * TILE;   // W[j][i]
    gemm_foreach_pair(acc, [&](int row, int col, double v0, double v1) {
        Wt[(long)col * Np + row] = v0;
        Wt[(long)(col + 1) * Np + row] = v1;
    });
}

__global__ void __launch_bounds__(G_THREADS, 1)
trtri_step2_kernel(double* __restrict__ S, const double* __restrict__ W, int Np, int nb, int h) {
    const int tj = blockIdx.x, ti = blockIdx.y, q = blockIdx.z;
    const int left = 2 * h * q, right = left + h;
    if (right + ti >= nb) return;
    extern __shared__ double smem[];
    const long r0 = (long)right * TILE;
    // A rows i: C^-1[i][k], k <= i  (diag block of S is already masked lower-triangular)
    Operand A = make_operand(S + ((long)(right + ti) * TILE) * Np + r0, Np);
    // B rows j: T^T[j][k] = W[left j][right k]
    Operand B = make_operand(W + ((long)(left + tj) * TILE) * Np + r0, Np);
    double acc[8][4][2];
    gemm_nt_mainloop(acc, A, B, 0, (ti + 1) * TILE, smem);
    double* Xt = S + ((long)(right + ti) * TILE) * Np + (long)(left + tj) * TILE;    // S[i][j]
    double* Xm = S + ((long)(left + tj) * TILE) * Np + (long)(right + ti) * TILE;    // S[j][i] mirror
    gemm_foreach_pair(acc, [&](int row, int col, double v0, double v1) {
        *reinterpret_cast<double2*>(Xt + (long)row * Np + col) = make_double2(-v0, -v1);
        Xm[(long)col * Np + row] = -v0;
        Xm[(long)(col + 1) * Np + row] = -v1;
    });
}

// ------------------------------------------------------------------------------------------------
// K_hat^-1 = L^-T L^-1:  C[i][j] = sum_{k >= blockstart(i)} Linv[k][i] Linv[k][j]   (tiles j <= i),
// mirrored to the upper triangle.  grid = (nb, nb)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(G_THREADS, 1)
lauum_kernel(const double* __restrict__ S, const double* __restrict__ DT, double* __restrict__ Kinv, int Np) {
    const int tj = blockIdx.x, ti = blockIdx.y;
    if (tj > ti) return;
    extern __shared__ double smem[];
    Operand A = make_operand_alt(S + ((long)ti * TILE) * Np, Np, DT + (long)ti * TILE * TILE, TILE, ti * TILE,
                                 (ti + 1) * TILE);
    Operand B = (tj == ti) ? A : make_operand(S + ((long)tj * TILE) * Np, Np);
    double acc[8][4][2];
    gemm_nt_mainloop(acc, A, B, ti * TILE, Np, smem);
    double* Ct = Kinv + ((long)ti * TILE) * Np + (long)tj * TILE;
    double* Cm = Kinv + ((long)tj * TILE) * Np + (long)ti * TILE;
    const bool diag = (ti == tj);
    gemm_foreach_pair(acc, [&](int row, int col, double v0, double v1) {
        *reinterpret_cast<double2*>(Ct + (long)row * Np + col) = make_double2(v0, v1);
        if (!diag) {
            Cm[(long)col * Np + row] = v0;
            Cm[(long)(col + 1) * Np + row] = v1;
        }
    });
}

// ------------------------------------------------------------------------------------------------
// Triangular mat-vecs on S (one warp per row, 8 rows per CTA):
//   mode 0:  u[i]     = sum_{k <= i} Linv[i][k] r[k]                      (rows of the lower part)
//   mode 1:  alpha[i] = sum_{k >= i} Linv[k][i] u[k] = DT-block part + mirror part of row i
// ------------------------------------------------------------------------------------------------
__global__ void trmv_kernel(const double* __restrict__ S, const double* __restrict__ DT, int Np,
                            const double* __restrict__ x, double* __restrict__ y, int mode) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int i = blockIdx.x * 8 + warp;
    if (i >= Np) return;
    const int b = i / TILE;
    const double* row = S + (long)i * Np;
    double s = 0.0;
    if (mode == 0) {
        const int kend = (b + 1) * TILE;    // masked zeros above the diagonal inside the block
        for (int k = lane; k < kend; k += 32) s = fma(row[k], x[k], s);
    } else {
        const double* drow = DT + (long)b * TILE * TILE + (long)(i - b * TILE) * TILE;
        for (int k = lane; k < TILE; k += 32) s = fma(drow[k], x[b * TILE + k], s);
        for (int k = (b + 1) * TILE + lane; k < Np; k += 32) s = fma(row[k], x[k], s);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) y[i] = s;
}

// loss = [ 1/2 r^T alpha + sum_blk logdet_blk + N/2 log(2 pi) ] / N      (= -mll)
// out[0] = loss, out[1] = quad = r^T alpha, out[2] = sum log L_ii
__global__ void mll_value_kernel(const double* __restrict__ r, const double* __restrict__ alpha, int Np, int N,
                                 const double* __restrict__ logdet_blk, int nb, double* __restrict__ out) {
    __shared__ double red[256];
    double q = 0.0;
    for (int i = threadIdx.x; i < Np; i += blockDim.x) q = fma(r[i], alpha[i], q);
    red[threadIdx.x] = q;
    __syncthreads();
    for (int o = blockDim.x / 2; o > 0; o >>= 1) {
        if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        double ld = 0.0;
        for (int b = 0; b < nb; ++b) ld += logdet_blk[b];
        const double quad = red[0];
        out[0] = (0.5 * quad + ld + 0.5 * N * 1.8378770664093453) / N;
        out[1] = quad;
        out[2] = ld;
    }
}

// ------------------------------------------------------------------------------------------------
// host drivers
// ------------------------------------------------------------------------------------------------
static bool g_attr_set[64] = {false};   // cudaFuncSetAttribute is per device
static void set_gemm_attrs() {
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && g_attr_set[dev]) return;
    cudaFuncSetAttribute(potrf_panel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G_SMEM_BYTES);
    cudaFuncSetAttribute(potrf_trailing_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G_SMEM_BYTES);
    cudaFuncSetAttribute(trtri_step1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G_SMEM_BYTES);
    cudaFuncSetAttribute(trtri_step2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G_SMEM_BYTES);
    cudaFuncSetAttribute(lauum_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G_SMEM_BYTES);
    const int dg = (TILE * DG_LD + 2 * TILE + 2) * 8;
    cudaFuncSetAttribute(potrf_diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, dg);
    if (dev >= 0 && dev < 64) g_attr_set[dev] = true;
}

// K (lower tiles valid) -> L in place; diagonal-block inverses into S / DT; per-block log-dets; info.
int potrf(double* K, int Np, int N, double* S, double* DT, double* logdet_blk, int* info, cudaStream_t st) {
    set_gemm_attrs();
    const int nb = Np / TILE;
    const int dg = (TILE * DG_LD + 2 * TILE + 2) * 8;
    cudaMemsetAsync(info, 0, sizeof(int), st);
    for (int p = 0; p < nb; ++p) {
        potrf_diag_kernel<<<1, DG_THREADS, dg, st>>>(K, Np, p, N, S, DT, logdet_blk, info);
        const int n = nb - p - 1;
        if (n > 0) {
            potrf_panel_kernel<<<n, G_THREADS, G_SMEM_BYTES, st>>>(K, Np, p, S);
            potrf_trailing_kernel<<<dim3(n, n), G_THREADS, G_SMEM_BYTES, st>>>(K, Np, p);
        }
    }
    GPX_CHECK_LAUNCH();
    return 0;
}

// Completes S (diagonal blocks already hold D_b) to the full mirrored inverse. W = Np x Np scratch.
int trtri(const double* L, double* S, const double* DT, double* W, int Np, cudaStream_t st) {
    set_gemm_attrs();
    const int nb = Np / TILE;
    for (int h = 1; h < nb; h *= 2) {
        const int pairs = (nb + 2 * h - 1) / (2 * h);
        dim3 grid(h, h, pairs);
        trtri_step1_kernel<<<grid, G_THREADS, G_SMEM_BYTES, st>>>(L, S, DT, W, Np, nb, h);
        trtri_step2_kernel<<<grid, G_THREADS, G_SMEM_BYTES, st>>>(S, W, Np, nb, h);
    }
    GPX_CHECK_LAUNCH();
    return 0;
}

int lauum(const double* S, const double* DT, double* Kinv, int Np, cudaStream_t st) {
    set_gemm_attrs();
    const int nb = Np / TILE;
    lauum_kernel<<<dim3(nb, nb), G_THREADS, G_SMEM_BYTES, st>>>(S, DT, Kinv, Np);
    GPX_CHECK_LAUNCH();
    return 0;
}

// alpha = Linv^T (Linv r);  r must be zero-padded to Np; u is scratch [Np].
int solve_alpha(const double* S, const double* DT, int Np, const double* r, double* u, double* alpha,
                cudaStream_t st) {
    const int grid = (Np + 7) / 8;
    trmv_kernel<<<grid, 256, 0, st>>>(S, DT, Np, r, u, 0);
    trmv_kernel<<<grid, 256, 0, st>>>(S, DT, Np, u, alpha, 1);
    GPX_CHECK_LAUNCH();
    return 0;
}

int mll_value(const double* r, const double* alpha, int Np, int N, const double* logdet_blk, double* out,
              cudaStream_t st) {
    mll_value_kernel<<<1, 256, 0, st>>>(r, alpha, Np, N, logdet_blk, Np / TILE, out);
    GPX_CHECK_LAUNCH();
    return 0;
}

}  // namespace gpx
