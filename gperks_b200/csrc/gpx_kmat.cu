// Pairwise-distance kernels: fused ARD RBF / Matern(1/2,3/2,5/2) + ScaleKernel + noise builders.
//
//   kmat_sym   : K_hat = s*k(X,X) + noise*I   (lower 128x128 tiles, identity in the padding)
//                replaces gpytorch ScaleKernel(RBF|Matern).forward + GaussianLikelihood noise add
//                (call sites GPErks/gp/model.py:18-21, GPErks/train/emulator.py:326).
//   kmat_cross : K*[m][j] = s*k(x*_m, X_j) written k(=j)-contiguous for the NT TRMM, with the
//                posterior-mean partial sums  sum_j K*[m][j]*alpha[j]  fused in
//                (GPErks/train/emulator.py:352-357).
//   mll_grad   : (1/2N) sum_ij (K_hat^-1 - alpha alpha^T)_ij dK_hat_ij/dtheta, kernel recomputed in
//                registers, two-stage deterministic reduction (no N^2*D tensor ever materialised).
//
// All share one tile body: a 128 x 128 tile per CTA, 256 threads, each thread an 8 x 8 micro-tile strided by 16 so that a warp's
// stores cover two full 128-byte row segments, walked two rows per pass (the distance loop is bound by shared-memory
// wavefronts: rows_q2).  Coordinates are staged (transposed, zero-padded to an even number of dimensions) pre-multiplied by
// coord_scale<KIND>() / lengthscale, and the squared-distance accumulators start at 1e-30, so the profile is evaluated from
// q = c^2 r^2 + 1e-30 by the branch-free exp / sqrt of gpx_math.cuh (kernel_profile, gpx_common.cuh).  The gradient kernel
// (mll_grad) keeps plain 1 / lengthscale coordinates (it needs the per-dimension squared differences themselves).
#include <cstdlib>
#include <type_traits>
#include "gpx_common.cuh"

namespace gpx {

constexpr int P_THREADS = 256;
constexpr int P_LD = TILE + 1;   // padded smem row (transposed coordinate tile [d][128])

// Stage rows [row0, row0+128) of X (row-major, ld = D) as z = ((x - a)/range) * inv_ls * cmul, transposed, into Dp >= D rows
// (rows D..Dp-1 are zero: the distance loops run over an even number of dimensions).  Rows >= n_valid are filled with zeros.
// xa/xr may be null (training inputs are already scaled).  Thread t owns tile row t % 128 and every second dimension.
__device__ __forceinline__ void stage_coords(double* sm, const double* __restrict__ X, long row0, long n_valid, int D,
                                             const double* __restrict__ theta, const double* __restrict__ xa,
                                             const double* __restrict__ xr, double cmul = 1.0, int Dp = -1) {
    if (Dp < 0) Dp = D;
    const int r = threadIdx.x & (TILE - 1);
    const bool valid = row0 + r < n_valid;
    for (int d = threadIdx.x >> 7; d < Dp; d += P_THREADS / TILE) {
        double v = 0.0;
        if (valid && d < D) {
            v = X[(row0 + r) * (long)D + d];
            if (xa != nullptr) v = (v - xa[d]) / xr[d];
            v *= theta[d] * cmul;
        }
        sm[d * P_LD + r] = v;
    }
}

// q[b] = 1e-30 + sum_d (sa[d][row] - sb[d][tx + 16 b])^2 over the Dp (even) staged coordinate rows: one rolled loop of two
// dimensions per trip (9 LDS + 16 FP64 per dimension for 8 entries), no remainder code.
__device__ __forceinline__ void row_q(double (&q)[8], const double* sa, const double* sb, int Dp, int row, int tx) {
#pragma unroll
    for (int b = 0; b < 8; ++b) q[b] = 1e-30;
#pragma unroll 1
    for (int d = 0; d < Dp; d += 2) {
        const double x0 = sa[d * P_LD + row], x1 = sa[(d + 1) * P_LD + row];
#pragma unroll
        for (int b = 0; b < 8; ++b) {
            const double d0 = x0 - sb[d * P_LD + tx + 16 * b];
            const double d1 = x1 - sb[(d + 1) * P_LD + tx + 16 * b];
            q[b] = fma(d0, d0, q[b]);
            q[b] = fma(d1, d1, q[b]);
        }
    }
}

// Two rows (row0 and row0 + 16) per pass: every train coordinate read from shared memory feeds two entries.  The distance loop
// is bound by shared-memory wavefronts (an LDS.64 delivers 16 lanes' worth per wavefront), not by the FP64 pipe: 1.25 wavefronts
// per entry and dimension here against 2.25 for one row per pass.
__device__ __forceinline__ void rows_q2(double (&q)[2][8], const double* sa, const double* sb, int Dp, int row0, int tx) {
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int b = 0; b < 8; ++b) q[r][b] = 1e-30;
#pragma unroll 1
    for (int d = 0; d < Dp; d += 2) {
        const double x00 = sa[d * P_LD + row0], x01 = sa[(d + 1) * P_LD + row0];
        const double x10 = sa[d * P_LD + row0 + 16], x11 = sa[(d + 1) * P_LD + row0 + 16];
#pragma unroll
        for (int b = 0; b < 8; ++b) {
            const double y0 = sb[d * P_LD + tx + 16 * b], y1 = sb[(d + 1) * P_LD + tx + 16 * b];
            const double d00 = x00 - y0, d01 = x01 - y1, d10 = x10 - y0, d11 = x11 - y1;
            q[0][b] = fma(d00, d00, q[0][b]);
            q[0][b] = fma(d01, d01, q[0][b]);
            q[1][b] = fma(d10, d10, q[1][b]);
            q[1][b] = fma(d11, d11, q[1][b]);
        }
    }
}

// r2[a][b] for rows ty+16a (from sa) and cols tx+16b (from sb)
__device__ __forceinline__ void tile_r2(double (&r2)[8][8], const double* sa, const double* sb, int D, int ty, int tx) {
#pragma unroll
    for (int a = 0; a < 8; ++a)
#pragma unroll
        for (int b = 0; b < 8; ++b) r2[a][b] = 0.0;
    for (int d = 0; d < D; ++d) {
        double xa_[8], xb_[8];
#pragma unroll
        for (int a = 0; a < 8; ++a) xa_[a] = sa[d * P_LD + ty + 16 * a];
#pragma unroll
        for (int b = 0; b < 8; ++b) xb_[b] = sb[d * P_LD + tx + 16 * b];
#pragma unroll
        for (int a = 0; a < 8; ++a)
#pragma unroll
            for (int b = 0; b < 8; ++b) {
                double df = xa_[a] - xb_[b];
                r2[a][b] = fma(df, df, r2[a][b]);
            }
    }
}

// ------------------------------------------------------------------------------------------------
// One row (8 entries) of the micro-tile at a time, in a ROLLED loop, like kmat_cross below: 64 unrolled FP64 sqrt/exp
// evaluations do not fit the instruction cache (this kernel ran at half the entry rate of kmat_cross before).
template <int KIND>
__global__ void __launch_bounds__(P_THREADS, 2) kmat_sym_kernel(const double* __restrict__ X, int N, int D,
                                                                const double* __restrict__ theta, double* __restrict__ K,
                                                                int Np) {
    const int bi = blockIdx.y, bj = blockIdx.x;
    if (bj > bi) return;
    extern __shared__ double sm[];
    double* sa = sm;
    double* sb = sm + MAXD * P_LD;
    const int Dp = D + (D & 1);
    stage_coords(sa, X, (long)bi * TILE, N, D, theta, nullptr, nullptr, coord_scale<KIND>(), Dp);
    stage_coords(sb, X, (long)bj * TILE, N, D, theta, nullptr, nullptr, coord_scale<KIND>(), Dp);
    __syncthreads();
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const double s = theta[D], noise = theta[D + 1], s3 = s / 3.0;
#pragma unroll 1
    for (int a0 = 0; a0 < 8; a0 += 2) {           // two rows per pass (rows_q2), evaluated one after the other
        double qq[2][8];
        rows_q2(qq, sa, sb, Dp, ty + 16 * a0, tx);
#pragma unroll
        for (int rr = 0; rr < 2; ++rr) {
        const int row = ty + 16 * (a0 + rr);
        const double (&q)[8] = qq[rr];
        const long i = (long)bi * TILE + row;
#pragma unroll
        for (int b = 0; b < 8; ++b) {
            const long j = (long)bj * TILE + tx + 16 * b;
            const bool in = (i < N && j < N);
            double v = kernel_profile<KIND>(q[b], in ? s : 0.0, in ? s3 : 0.0);   // branch-free; the padding block is the identity
            if (i == j) v += in ? noise : 1.0;
            K[i * Np + j] = v;
        }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Cross kernel for one chunk of test points.  grid = (Np/128 j-tiles, Mc/128 m-tiles).
//   Ks[m][j]  (ld = Np)            m in [0, Mc), chunk-local
//   mean_part[jb][m]  (ld = Mc)    partial  sum_{j in tile jb} Ks[m][j] * alpha[j]
// The 8 x 8 micro-tile is walked one row (8 entries) at a time in a ROLLED loop: fully unrolling the 64
// double-precision sqrt/exp evaluations produced ~80 KB of code and the kernel stalled on instruction fetch
// (ncu: "no_instruction" was the top stall, FP64 pipe 18 % busy); one row keeps the body inside the I-cache and
// the register count low enough for several resident CTAs.
template <int KIND>
__global__ void __launch_bounds__(P_THREADS, 2)
kmat_cross_kernel(const double* __restrict__ Xs, long m_valid, const double* __restrict__ X, int N, int D,
                  const double* __restrict__ theta, const double* __restrict__ xa, const double* __restrict__ xr,
                  const double* __restrict__ alpha, double* __restrict__ Ks, int Np, double* __restrict__ mean_part,
                  int Mc, float* __restrict__ Khi, float* __restrict__ Klo) {
    const int jb = blockIdx.x, mb = blockIdx.y;
    extern __shared__ double sm[];
    double* sa = sm;                    // test points (rows m)
    double* sb = sm + MAXD * P_LD;      // train points (cols j)
    double* sal = sb + MAXD * P_LD;     // alpha tile [128]
    double* smean = sal + TILE;         // [128]
    const int Dp = D + (D & 1);
    stage_coords(sa, Xs, (long)mb * TILE, m_valid, D, theta, xa, xr, coord_scale<KIND>(), Dp);
    stage_coords(sb, X, (long)jb * TILE, N, D, theta, nullptr, nullptr, coord_scale<KIND>(), Dp);
    if (threadIdx.x < TILE) sal[threadIdx.x] = alpha[jb * TILE + threadIdx.x];   // alpha is zero-padded to Np
    __syncthreads();
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const double s = theta[D], s3 = s / 3.0;
    double al[8];
#pragma unroll
    for (int b = 0; b < 8; ++b) al[b] = sal[tx + 16 * b];
#pragma unroll 1
    for (int a0 = 0; a0 < 8; a0 += 2) {           // two rows per pass (rows_q2), evaluated one after the other
        double qq[2][8];
        rows_q2(qq, sa, sb, Dp, ty + 16 * a0, tx);
#pragma unroll
        for (int rr = 0; rr < 2; ++rr) {
        const int row = ty + 16 * (a0 + rr);
        const double (&r2)[8] = qq[rr];
        const long m = (long)mb * TILE + row;
        double part = 0.0, part1 = 0.0;          // columns 0..63 / 64..127 as two chains (the order gpx_kmat_cross_i8 sums in)
#pragma unroll
        for (int b = 0; b < 8; ++b) {
            const long j = (long)jb * TILE + tx + 16 * b;
            const bool in = j < N;
            double v = kernel_profile<KIND>(r2[b], in ? s : 0.0, in ? s3 : 0.0);     // masked by the outputscale: no branch around the evaluation
            if (Khi != nullptr) {       // TF32 mode: 3xTF32 operand split (hi = tf32(x), lo = tf32(x - hi))
                uint32_t h, l;
                float fh, fl;
                asm("cvt.rna.tf32.f32 %0, %1;\n" : "=r"(h) : "f"((float)v));
                fh = __uint_as_float(h);
                asm("cvt.rna.tf32.f32 %0, %1;\n" : "=r"(l) : "f"((float)(v - (double)fh)));
                fl = __uint_as_float(l);
                Khi[m * Np + j] = fh;
                Klo[m * Np + j] = fl;
            } else {
                Ks[m * Np + j] = v;
            }
            if (b < 4) part = fma(v, al[b], part); else part1 = fma(v, al[b], part1);
        }
        part += part1;
        // reduce over the 16 tx lanes (a half-warp holds one row)
        part += __shfl_xor_sync(0xffffffffu, part, 1);
        part += __shfl_xor_sync(0xffffffffu, part, 2);
        part += __shfl_xor_sync(0xffffffffu, part, 4);
        part += __shfl_xor_sync(0xffffffffu, part, 8);
        if (tx == 0) smean[row] = part;
        }
    }
    __syncthreads();
    if (threadIdx.x < TILE) mean_part[(long)jb * Mc + (long)mb * TILE + threadIdx.x] = smean[threadIdx.x];
}

// ------------------------------------------------------------------------------------------------
// Sliced-integer mode (gpx_i8.cu): K* evaluated in FP64 exactly like kmat_cross_kernel -- same micro-tile, same order of
// the fused K* alpha partial sums, so the posterior mean is bit-identical -- and written directly as S int8 slices
// (planes [S][plane_rows][Np]) of  v * 2^-(e+1)  (outputscale = f 2^e, k <= outputscale) instead of 8-byte values that a
// second kernel would read back and slice.  BITS = 7: S rounds of round-to-nearest; BITS = 8: one FP64 -> int64 conversion and
// the bytes of the biased integer (gpx_common.cuh: balanced_digits) -- two FP64 instructions per entry instead of 4 S.
// BITS = 8 as built now: the outputscale carries the power-of-two digit scale, so the profile's value IS the integer to be rounded
// and its digits are the low mantissa bytes of ONE add of 1.5 * 2^52 + bias (below).
// A thread owns the eight columns tx + 16 b of two rows per pass (the FP64 kernels' mapping).  The plane layout stores column c
// of a 128-block at byte (c % 16) * 8 + c / 16 (the L^-1 planes use the same permutation), so the 8 bytes of a slice are
// assembled in two registers with one PRMT per digit and leave as ONE 8-byte store per slice: a half-warp writes a full 128-byte
// row segment, no shared-memory staging.  (A variant with four columns x four rows per thread -- 78 registers, three CTAs per
// SM -- measured 3-5 % slower at every shape, profiles/r02_builder_bench.json; occupancy is not what limits this kernel.)
// Shared memory holds only the 2 x D coordinate rows (sized by D, not MAXD), so a CTA of this kernel fits beside the 183 KB CTA
// of the integer tensor-core kernel working on the previous chunk.
template <int KIND, int S, int BITS>
__global__ void __launch_bounds__(P_THREADS, 2)
kmat_cross_i8_kernel(const double* __restrict__ Xs, long m_valid, const double* __restrict__ X, int N, int D,
                     const double* __restrict__ theta, const double* __restrict__ xa, const double* __restrict__ xr,
                     const double* __restrict__ alpha, int8_t* __restrict__ planes, long plane_rows, int Np,
                     double* __restrict__ mean_part, int Mc) {
    const int jb = blockIdx.x, mb = blockIdx.y;
    extern __shared__ double sm[];
    const int Dp = D + (D & 1);
    double* sa = sm;                    // test points (rows m)   [Dp][P_LD]
    double* sb = sm + Dp * P_LD;        // train points (cols j)  [Dp][P_LD]
    double* sal = sb + Dp * P_LD;       // alpha tile [128]
    double* smean = sal + TILE;         // [128]
    stage_coords(sa, Xs, (long)mb * TILE, m_valid, D, theta, xa, xr, coord_scale<KIND>(), Dp);
    stage_coords(sb, X, (long)jb * TILE, N, D, theta, nullptr, nullptr, coord_scale<KIND>(), Dp);
    if (threadIdx.x < TILE) sal[threadIdx.x] = alpha[jb * TILE + threadIdx.x];
    __syncthreads();
    constexpr int EPT = 8;                      // entries (columns) per thread and row
    constexpr int NH = 8 / EPT;                 // threads sharing one 8-byte group of a row (1)
    constexpr int NW = EPT / 4;                 // 32-bit words per slice and thread
    const int tx = threadIdx.x & 15, h = (threadIdx.x >> 4) & (NH - 1), ty = threadIdx.x / (16 * NH);
    const int e = scale_exponent_b<BITS>(theta[D]);
    const double scale = ldexp(1.0, (BITS == 7 ? 7 : 0) - (e + 1));      // 7 bit: first slice = rint(v * 2^-(e+1) * 2^7), |.| <= 64
    // 8 bit: the kernel is evaluated with the outputscale pre-multiplied by the power of two  scale * 2^(8S)  (exact, commutes
    // with every rounding), so v8 = V-to-be-rounded directly, and alpha is divided by the same power: v8 * al8 == v * al bit for
    // bit and the fused mean sums stay those of the FP64 path.
    const double pre = (BITS == 8) ? scale * (double)(1ull << (8 * S)) : 1.0;
    const double s = theta[D] * pre, s3 = s / 3.0;
    const long pitch = plane_rows * (long)Np;
    const int c0 = tx + 16 * EPT * h;                                      // first of the thread's EPT columns (stride 16)
    double al[EPT];
#pragma unroll
    for (int b = 0; b < EPT; ++b) al[b] = sal[c0 + 16 * b] * (1.0 / pre);
    // Only the last train tile can hold columns j >= N (their K* entries must be 0: the padding block of L^-1 is the identity);
    // the choice is uniform per CTA, so every other CTA runs the copy of the row loop without the per-entry mask.
    auto rows = [&](auto masked_tag) {
        constexpr bool MASKED = decltype(masked_tag)::value;
        // R rows per pass share every train coordinate read from shared memory: the distance loop is bound by shared-memory
        // wavefronts, not by the FP64 pipe (an LDS.64 delivers 16 lanes' worth per wavefront; with one row per pass it cost
        // 2.25 wavefronts per entry and dimension = 9 SM cycles against 4 FP64 pipe cycles -- measured 10 cycles per dimension)
        constexpr int R = 2, RSTEP = 16 / NH;
#pragma unroll 1
        for (int a0 = 0; a0 < 8 * NH; a0 += R) {
            double qq[R][EPT];
#pragma unroll
            for (int r = 0; r < R; ++r)
#pragma unroll
                for (int b = 0; b < EPT; ++b) qq[r][b] = 1e-30;
#pragma unroll 1
            for (int d = 0; d < Dp; d += 2) {
                double x0[R], x1[R];
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    x0[r] = sa[d * P_LD + ty + RSTEP * (a0 + r)];
                    x1[r] = sa[(d + 1) * P_LD + ty + RSTEP * (a0 + r)];
                }
#pragma unroll
                for (int b = 0; b < EPT; ++b) {
                    const double y0 = sb[d * P_LD + c0 + 16 * b], y1 = sb[(d + 1) * P_LD + c0 + 16 * b];
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        const double d0 = x0[r] - y0, d1 = x1[r] - y1;
                        qq[r][b] = fma(d0, d0, qq[r][b]);
                        qq[r][b] = fma(d1, d1, qq[r][b]);
                    }
                }
            }
#pragma unroll
            for (int r = 0; r < R; ++r) {
            const int row = ty + RSTEP * (a0 + r);
            const double (&q)[EPT] = qq[r];
            double part = 0.0, part1 = 0.0;
            uint32_t word[S][NW];
#pragma unroll
            for (int k = 0; k < S; ++k)
#pragma unroll
                for (int w = 0; w < NW; ++w) word[k][w] = 0u;
#pragma unroll
            for (int b = 0; b < EPT; ++b) {
                const bool in = !MASKED || jb * TILE + c0 + 16 * b < N;
                const double v = kernel_profile<KIND>(q[b], in ? s : 0.0, in ? s3 : 0.0);   // masked by the outputscale: no branch
                if (b < 4) part = fma(v, al[b], part); else part1 = fma(v, al[b], part1);
                // byte b of the slice words: selector keeps the other three bytes and takes one byte of the digit word
                constexpr uint32_t keep = 0x3210u;
                const uint32_t pos = (uint32_t)(b & 3) * 4;
                const int wi = b >> 2;
                if (BITS == 7) {
                    double r = v * scale;
#pragma unroll
                    for (int k = 0; k < S; ++k) {
                        const double qi = rint(r);
                        r = (r - qi) * 128.0;                            // exact: |r - qi| <= 1/2
                        const uint32_t w = (uint32_t)__double2int_rn(qi);
                        const uint32_t sel = (keep & ~(0xFu << pos)) | (4u << pos);
                        word[k][wi] = __byte_perm(word[k][wi], w, sel);
                    }
                } else {
                    // biased digits (V + BIAS in the low mantissa bits of one DADD); the ^ 0x80 of every byte that makes them
                    // balanced is applied to the assembled slice words below
                    const unsigned long long dg = (unsigned long long)__double_as_longlong(
                        v + (6755399441055744.0 + (double)((S == 5) ? 0x8080808080ull : 0x808080808080ull)));
                    const uint32_t w0 = (uint32_t)dg, w1 = (uint32_t)(dg >> 32);
#pragma unroll
                    for (int k = 0; k < S; ++k) {
                        const int src = S - 1 - k;                       // digit k = byte S-1-k
                        const uint32_t sel = (keep & ~(0xFu << pos)) | ((uint32_t)(4 + (src & 3)) << pos);
                        word[k][wi] = __byte_perm(word[k][wi], src < 4 ? w0 : w1, sel);
                    }
                }
            }
            // same order as the FP64 kernels: (columns 0..63 chain) + (columns 64..127 chain), then the tree over tx
            part += part1;
            part += __shfl_xor_sync(0xffffffffu, part, 1);
            part += __shfl_xor_sync(0xffffffffu, part, 2);
            part += __shfl_xor_sync(0xffffffffu, part, 4);
            part += __shfl_xor_sync(0xffffffffu, part, 8);
            if (tx == 0 && h == 0) smean[row] = part;
            int8_t* dst = planes + ((long)mb * TILE + row) * Np + (long)jb * TILE + tx * 8 + 4 * h;   // 128 contiguous bytes per row
#pragma unroll
            for (int k = 0; k < S; ++k) {
                constexpr uint32_t flip = (BITS == 8) ? 0x80808080u : 0u;
                if (NW == 1) *reinterpret_cast<uint32_t*>(dst + k * pitch) = word[k][0] ^ flip;
                else *reinterpret_cast<uint2*>(dst + k * pitch) = make_uint2(word[k][0] ^ flip, word[k][NW - 1] ^ flip);
            }
            }
        }
    };
    if ((jb + 1) * TILE <= N) rows(std::false_type{}); else rows(std::true_type{});
    __syncthreads();
    if (threadIdx.x < TILE) mean_part[(long)jb * Mc + (long)mb * TILE + threadIdx.x] = smean[threadIdx.x];
}

template <int KIND>
static int launch_kmat_cross_i8(const double* Xs, long m_valid, const double* X, int N, int D, const double* theta,
                                const double* xa, const double* xr, const double* alpha, int8_t* planes, long plane_rows,
                                int nslices, int bits, int Np, double* mean_part, int Mc, cudaStream_t st) {
    const size_t smem = (size_t(2) * (D + (D & 1)) * P_LD + 2 * TILE) * 8;
    const dim3 grid(Np / TILE, Mc / TILE);
#define GPX_KX(S_, B_)                                                                                                        \
    do {                                                                                                                      \
        if (smem > 48 * 1024)                                                                                                 \
            cudaFuncSetAttribute(kmat_cross_i8_kernel<KIND, S_, B_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
        cudaFuncSetAttribute(kmat_cross_i8_kernel<KIND, S_, B_>, cudaFuncAttributePreferredSharedMemoryCarveout,              \
                             cudaSharedmemCarveoutMaxShared);                                                                 \
        kmat_cross_i8_kernel<KIND, S_, B_><<<grid, P_THREADS, smem, st>>>(Xs, m_valid, X, N, D, theta, xa, xr, alpha, planes, \
                                                                          plane_rows, Np, mean_part, Mc);                     \
    } while (0)
    if (nslices == 5 && bits == 7) GPX_KX(5, 7);
    else if (nslices == 6 && bits == 7) GPX_KX(6, 7);
    else if (nslices == 5 && bits == 8) GPX_KX(5, 8);
    else if (nslices == 6 && bits == 8) GPX_KX(6, 8);
    else return -8;
#undef GPX_KX
    GPX_CHECK_LAUNCH();
    return 0;
}

// ------------------------------------------------------------------------------------------------
// Posterior MEAN only:  mu*_m = m(x*_m) + sum_j k(x*_m, X_j) alpha_j, un-standardised.  One CTA owns 128 test points and
// walks all train tiles; K* never exists in memory (8 D bytes in, 8 bytes out per point; N (3 D + c) FP64 FLOP per point).
// Same micro-tile, same fused partial sums (FMA chain over the thread's 8 columns, xor-shuffle tree over the half-warp) and the
// same order of the per-tile partials and of the prior-mean terms as kmat_cross_kernel + predict_finalize_kernel, so the
// result is bit-identical to the mean of gpx_predict_meanvar.  Used where the variance is not needed: Sobol' indices of the
// emulator's posterior mean over 1e7-point Saltelli designs (GPErks/perks/gsa.py:66-88 evaluates the emulator there).
template <int KIND>
__global__ void __launch_bounds__(P_THREADS, 2)
predict_mean_kernel(const double* __restrict__ Xs, long m_valid, const double* __restrict__ X, int N, int D,
                    const double* __restrict__ theta, const double* __restrict__ xa, const double* __restrict__ xr,
                    const double* __restrict__ alpha, int nb, const double* __restrict__ prior_mean,
                    const int* __restrict__ feat_idx, int n_feat, int degree, const double* __restrict__ weights,
                    const double* __restrict__ bias, double y_mu, double y_sigma, double* __restrict__ out_mean) {
    const long mb = blockIdx.x;
    extern __shared__ double sm[];
    const int Dp = D + (D & 1);
    double* sa = sm;                    // test points (rows m)   [Dp][P_LD]
    double* sb = sm + Dp * P_LD;        // train points (cols j)  [Dp][P_LD]
    double* sal = sb + Dp * P_LD;       // alpha tile [128]
    double* smean = sal + TILE;         // [128]
    stage_coords(sa, Xs, mb * TILE, m_valid, D, theta, xa, xr, coord_scale<KIND>(), Dp);
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const double s = theta[D], s3 = s / 3.0;
    if (threadIdx.x < TILE) smean[threadIdx.x] = 0.0;      // running sum over the train tiles, owned by the row's tx == 0 lane
    for (int jb = 0; jb < nb; ++jb) {
        __syncthreads();                // previous tile's sb / sal are no longer read (and sa is staged before the first use)
        stage_coords(sb, X, (long)jb * TILE, N, D, theta, nullptr, nullptr, coord_scale<KIND>(), Dp);
        if (threadIdx.x < TILE) sal[threadIdx.x] = alpha[jb * TILE + threadIdx.x];
        __syncthreads();
        double al[8];
#pragma unroll
        for (int b = 0; b < 8; ++b) al[b] = sal[tx + 16 * b];
#pragma unroll 1
        for (int a0 = 0; a0 < 8; a0 += 2) {           // two rows per pass (rows_q2), evaluated one after the other
            double qq[2][8];
            rows_q2(qq, sa, sb, Dp, ty + 16 * a0, tx);
#pragma unroll
            for (int rr = 0; rr < 2; ++rr) {
            const int row = ty + 16 * (a0 + rr);
            const double (&r2)[8] = qq[rr];
            double part = 0.0, part1 = 0.0;      // columns 0..63 / 64..127 as two chains, like kmat_cross_kernel
#pragma unroll
            for (int b = 0; b < 8; ++b) {
                const long j = (long)jb * TILE + tx + 16 * b;
                const bool in = j < N;
                const double v = kernel_profile<KIND>(r2[b], in ? s : 0.0, in ? s3 : 0.0);   // masked by the outputscale: no branch
                if (b < 4) part = fma(v, al[b], part); else part1 = fma(v, al[b], part1);
            }
            part += part1;
            part += __shfl_xor_sync(0xffffffffu, part, 1);
            part += __shfl_xor_sync(0xffffffffu, part, 2);
            part += __shfl_xor_sync(0xffffffffu, part, 4);
            part += __shfl_xor_sync(0xffffffffu, part, 8);
            if (tx == 0) smean[row] += part;
            }
        }
    }
    __syncthreads();
    if (threadIdx.x < TILE) {
        const long m = mb * TILE + threadIdx.x;
        if (m < m_valid) {
            double mean = 0.0;
            if (prior_mean != nullptr) {
                mean = prior_mean[m];
            } else {
                if (weights != nullptr) {
                    for (int p = 0; p < n_feat; ++p) {
                        double f = 1.0;
                        for (int q = 0; q < degree; ++q) {
                            const int d = feat_idx[p * degree + q];
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
            mean += smean[threadIdx.x];
            out_mean[m] = y_sigma * mean + y_mu;
        }
    }
}

template <int KIND>
static int launch_predict_mean(const double* Xs, long M, const double* X, int N, int D, const double* theta, const double* xa,
                               const double* xr, const double* alpha, int Np, const double* prior_mean, const int* feat_idx,
                               int n_feat, int degree, const double* weights, const double* bias, double y_mu, double y_sigma,
                               double* out_mean, cudaStream_t st) {
    const size_t smem = (size_t(2) * (D + (D & 1)) * P_LD + 2 * TILE) * 8;
    if (smem > 48 * 1024)
        cudaFuncSetAttribute(predict_mean_kernel<KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const long blocks = (M + TILE - 1) / TILE;
    predict_mean_kernel<KIND><<<(unsigned)blocks, P_THREADS, smem, st>>>(Xs, M, X, N, D, theta, xa, xr, alpha, Np / TILE, prior_mean,
                                                                          feat_idx, n_feat, degree, weights, bias, y_mu, y_sigma,
                                                                          out_mean);
    GPX_CHECK_LAUNCH();
    return 0;
}

// ------------------------------------------------------------------------------------------------
// TF32-mode cross kernel: K* evaluated in FP32 (coordinates are scaled in FP64, rounded once) and written as the
// (hi, lo) TF32 operand pair of the 3xTF32 contraction; the posterior-mean partials accumulate in FP64.
// Relative error of an entry ~3e-7 -- far inside the mode's 1e-3 tolerance -- at ~1/5 of the FP64 builder's cost.
__device__ __forceinline__ void stage_coords_f32(float* sm, const double* __restrict__ X, long row0, long n_valid, int D,
                                                 const double* __restrict__ theta, const double* __restrict__ xa,
                                                 const double* __restrict__ xr) {
    for (int e = threadIdx.x; e < TILE * D; e += P_THREADS) {
        int r = e / D, d = e - r * D;
        double v = 0.0;
        if (row0 + r < n_valid) {
            v = X[(row0 + r) * (long)D + d];
            if (xa != nullptr) v = (v - xa[d]) / xr[d];
            v *= theta[d];
        }
        sm[d * P_LD + r] = (float)v;
    }
}

template <int KIND>
__device__ __forceinline__ float kernel_value_f32(float r2, float s) {
    if (KIND == RBF) return s * expf(-0.5f * r2);
    const float r = sqrtf(fmaxf(r2, 1e-30f));
    if (KIND == MATERN12) return s * expf(-r);
    if (KIND == MATERN32) {
        const float c = 1.7320508075688772f;
        return s * (1.0f + c * r) * expf(-c * r);
    }
    const float c = 2.23606797749979f;
    return s * (1.0f + c * r + (5.0f / 3.0f) * r2) * expf(-c * r);
}

template <int KIND>
__global__ void __launch_bounds__(P_THREADS, 2)
kmat_cross_f32_kernel(const double* __restrict__ Xs, long m_valid, const double* __restrict__ X, int N, int D,
                      const double* __restrict__ theta, const double* __restrict__ xa, const double* __restrict__ xr,
                      const double* __restrict__ alpha, float* __restrict__ Khi, float* __restrict__ Klo, int Np,
                      double* __restrict__ mean_part, int Mc) {
    const int jb = blockIdx.x, mb = blockIdx.y;
    extern __shared__ double smd[];
    float* sa = reinterpret_cast<float*>(smd);     // test points  [d][129]
    float* sb = sa + MAXD * P_LD;                   // train points [d][129]
    double* sal = reinterpret_cast<double*>(sb + MAXD * P_LD + 2);   // alpha tile [128] (8-byte aligned: 2*32*129+2 floats)
    double* smean = sal + TILE;
    stage_coords_f32(sa, Xs, (long)mb * TILE, m_valid, D, theta, xa, xr);
    stage_coords_f32(sb, X, (long)jb * TILE, N, D, theta, nullptr, nullptr);
    if (threadIdx.x < TILE) sal[threadIdx.x] = alpha[jb * TILE + threadIdx.x];
    __syncthreads();
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    float r2[8][8];
#pragma unroll
    for (int a = 0; a < 8; ++a)
#pragma unroll
        for (int b = 0; b < 8; ++b) r2[a][b] = 0.f;
    for (int d = 0; d < D; ++d) {
        float xa_[8], xb_[8];
#pragma unroll
        for (int a = 0; a < 8; ++a) xa_[a] = sa[d * P_LD + ty + 16 * a];
#pragma unroll
        for (int b = 0; b < 8; ++b) xb_[b] = sb[d * P_LD + tx + 16 * b];
#pragma unroll
        for (int a = 0; a < 8; ++a)
#pragma unroll
            for (int b = 0; b < 8; ++b) {
                const float df = xa_[a] - xb_[b];
                r2[a][b] = fmaf(df, df, r2[a][b]);
            }
    }
    const float s = (float)theta[D];
#pragma unroll
    for (int a = 0; a < 8; ++a) {
        const long m = (long)mb * TILE + ty + 16 * a;
        double part = 0.0;
#pragma unroll
        for (int b = 0; b < 8; ++b) {
            const long j = (long)jb * TILE + tx + 16 * b;
            const float v = (j < N) ? kernel_value_f32<KIND>(r2[a][b], s) : 0.f;
            uint32_t h, l;
            asm("cvt.rna.tf32.f32 %0, %1;\n" : "=r"(h) : "f"(v));
            const float fh = __uint_as_float(h);
            asm("cvt.rna.tf32.f32 %0, %1;\n" : "=r"(l) : "f"(v - fh));
            Khi[m * Np + j] = fh;
            Klo[m * Np + j] = __uint_as_float(l);
            part = fma((double)v, sal[tx + 16 * b], part);
        }
        part += __shfl_xor_sync(0xffffffffu, part, 1);
        part += __shfl_xor_sync(0xffffffffu, part, 2);
        part += __shfl_xor_sync(0xffffffffu, part, 4);
        part += __shfl_xor_sync(0xffffffffu, part, 8);
        if (tx == 0) smean[ty + 16 * a] = part;
    }
    __syncthreads();
    if (threadIdx.x < TILE) mean_part[(long)jb * Mc + (long)mb * TILE + threadIdx.x] = smean[threadIdx.x];
}

template <int KIND>
static int launch_kmat_cross_f32(const double* Xs, long m_valid, const double* X, int N, int D, const double* theta,
                                 const double* xa, const double* xr, const double* alpha, float* Khi, float* Klo,
                                 int Np, double* mean_part, int Mc, cudaStream_t st) {
    const size_t smem = (size_t(2) * MAXD * P_LD + 2) * 4 + 2 * TILE * 8;
    cudaFuncSetAttribute(kmat_cross_f32_kernel<KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    kmat_cross_f32_kernel<KIND><<<dim3(Np / TILE, Mc / TILE), P_THREADS, smem, st>>>(Xs, m_valid, X, N, D, theta, xa, xr,
                                                                                     alpha, Khi, Klo, Np, mean_part, Mc);
    GPX_CHECK_LAUNCH();
    return 0;
}

// ------------------------------------------------------------------------------------------------
// Gradient reduction.  grid = lower tiles (bj <= bi).  partial[(bi*nb + bj)*(D+2) + p]:
//   p <  D : sum W_ij g(r_ij) dz_ijd^2     (weight 2 for off-diagonal tiles)
//   p == D : sum W_ij k_ij / s
//   p == D+1 : sum_i W_ii
template <int KIND>
__global__ void __launch_bounds__(P_THREADS, 3)
mll_grad_tile_kernel(const double* __restrict__ X, int N, int D, const double* __restrict__ theta,
                     const double* __restrict__ Kinv, int Np, const double* __restrict__ alpha,
                     double* __restrict__ partial) {
    const int bi = blockIdx.y, bj = blockIdx.x, nb = gridDim.y;
    double* out = partial + ((long)bi * nb + bj) * (D + 2);
    if (bj > bi) return;
    extern __shared__ double sm[];
    double* sa = sm;
    double* sb = sm + MAXD * P_LD;
    double* sali = sb + MAXD * P_LD;   // alpha_i [128]
    double* salj = sali + TILE;        // alpha_j [128]
    double* sred = salj + TILE;        // [8 warps][MAXD+2]
    stage_coords(sa, X, (long)bi * TILE, N, D, theta, nullptr, nullptr);
    stage_coords(sb, X, (long)bj * TILE, N, D, theta, nullptr, nullptr);
    if (threadIdx.x < TILE) {
        sali[threadIdx.x] = alpha[bi * TILE + threadIdx.x];
        salj[threadIdx.x] = alpha[bj * TILE + threadIdx.x];
    }
    __syncthreads();
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const double s = theta[D];
    const double tile_w = (bi == bj) ? 1.0 : 2.0;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // One row (8 entries) of the micro-tile at a time in a ROLLED loop (instruction-cache footprint, see kmat_cross):
    // r2 -> weight wg = W*g and the outputscale / noise sums, then the per-dimension sums  sum wg * dz_d^2  into this
    // thread's column of sacc[d][thread] (shared memory: D is a run-time value, a register array could not be indexed).
    double* sacc = sred + 8 * (MAXD + 2);      // [D][256]
    for (int d = 0; d < D; ++d) sacc[d * P_THREADS + threadIdx.x] = 0.0;
    double acc_s = 0.0, acc_n = 0.0;
    // K_hat^-1 entries of the NEXT row are requested before this row's distances and sqrt/exp are evaluated (the tile is
    // full: Np is padded, so the loads need no bounds check); without it every row waited on its own eight loads
    // (ncu: long-scoreboard 3.8 warps per issue, 0.63 TB/s).
    const double* wbase = Kinv + ((long)bi * TILE + ty) * Np + (long)bj * TILE + tx;
    double wnext[8];
#pragma unroll
    for (int b = 0; b < 8; ++b) wnext[b] = __ldcs(wbase + 16 * b);
#pragma unroll 1
    for (int a = 0; a < 8; ++a) {
        const int row = ty + 16 * a;
        double wcur[8];
#pragma unroll
        for (int b = 0; b < 8; ++b) wcur[b] = wnext[b];
        if (a < 7) {
#pragma unroll
            for (int b = 0; b < 8; ++b) wnext[b] = __ldcs(wbase + (long)(16 * (a + 1)) * Np + 16 * b);
        }
        double r2[8];
#pragma unroll
        for (int b = 0; b < 8; ++b) r2[b] = 0.0;
        for (int d = 0; d < D; ++d) {
            const double xa_ = sa[d * P_LD + row];
#pragma unroll
            for (int b = 0; b < 8; ++b) {
                const double df = xa_ - sb[d * P_LD + tx + 16 * b];
                r2[b] = fma(df, df, r2[b]);
            }
        }
        const long i = (long)bi * TILE + row;
        const double ai = sali[row];
        double wg[8];
#pragma unroll
        for (int b = 0; b < 8; ++b) {
            const long j = (long)bj * TILE + tx + 16 * b;
            const bool in = (i < N && j < N);
            double g;
            const double k = kernel_value_grad<KIND>(r2[b], s, g);          // evaluated unconditionally (branch-free); w masks it
            const double w = in ? wcur[b] - ai * salj[tx + 16 * b] : 0.0;
            acc_s = fma(w, k, acc_s);
            if (i == j) acc_n += w;
            wg[b] = w * g;
        }
        for (int d = 0; d < D; ++d) {
            const double xa_ = sa[d * P_LD + row];
            double acc = sacc[d * P_THREADS + threadIdx.x];
#pragma unroll
            for (int b = 0; b < 8; ++b) {
                const double df = xa_ - sb[d * P_LD + tx + 16 * b];
                acc = fma(wg[b], df * df, acc);
            }
            sacc[d * P_THREADS + threadIdx.x] = acc;
        }
    }
    for (int d = 0; d < D; ++d) {
        double acc = sacc[d * P_THREADS + threadIdx.x];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) sred[warp * (MAXD + 2) + d] = acc;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        acc_s += __shfl_xor_sync(0xffffffffu, acc_s, o);
        acc_n += __shfl_xor_sync(0xffffffffu, acc_n, o);
    }
    if (lane == 0) {
        sred[warp * (MAXD + 2) + D] = acc_s / s;
        sred[warp * (MAXD + 2) + D + 1] = acc_n;
    }
    __syncthreads();
    if (threadIdx.x < D + 2) {
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < 8; ++w) t += sred[w * (MAXD + 2) + threadIdx.x];
        out[threadIdx.x] = (threadIdx.x == D + 1) ? t : t * tile_w;
    }
}

// final deterministic reduction over tiles: grad[p] = scale_p * sum_tiles partial
//   grad[d] (lengthscale): * inv_ls_d / (2N)   (dz^2 = delta^2/l^2  =>  delta^2/l^3 = dz^2 / l)
//   NOTE: d k/d l = +g * delta^2 / l^3.
__global__ void mll_grad_final_kernel(const double* __restrict__ partial, int nb, int D, int N,
                                      const double* __restrict__ theta, double* __restrict__ grad) {
    const int p = blockIdx.x;
    __shared__ double red[256];
    double t = 0.0;
    for (int bi = 0; bi < nb; ++bi)
        for (int bj = threadIdx.x; bj <= bi; bj += blockDim.x) t += partial[((long)bi * nb + bj) * (D + 2) + p];
    red[threadIdx.x] = t;
    __syncthreads();
    for (int o = blockDim.x / 2; o > 0; o >>= 1) {
        if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        double v = red[0] / (2.0 * N);
        if (p < D) v *= theta[p];
        grad[p] = v;
    }
}

// ------------------------------------------------------------------------------------------------
// host launchers
// ------------------------------------------------------------------------------------------------
template <int KIND>
static int launch_kmat_sym(const double* X, int N, int D, const double* theta, double* K, int Np, cudaStream_t st) {
    const int nb = Np / TILE;
    const size_t smem = size_t(2) * MAXD * P_LD * 8;
    cudaFuncSetAttribute(kmat_sym_kernel<KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    kmat_sym_kernel<KIND><<<dim3(nb, nb), P_THREADS, smem, st>>>(X, N, D, theta, K, Np);
    GPX_CHECK_LAUNCH();
    return 0;
}

template <int KIND>
static int launch_kmat_cross(const double* Xs, long m_valid, const double* X, int N, int D, const double* theta,
                             const double* xa, const double* xr, const double* alpha, double* Ks, int Np,
                             double* mean_part, int Mc, float* Khi, float* Klo, cudaStream_t st) {
    const size_t smem = size_t(2) * MAXD * P_LD * 8 + 2 * TILE * 8;
    cudaFuncSetAttribute(kmat_cross_kernel<KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    kmat_cross_kernel<KIND><<<dim3(Np / TILE, Mc / TILE), P_THREADS, smem, st>>>(Xs, m_valid, X, N, D, theta, xa, xr,
                                                                                 alpha, Ks, Np, mean_part, Mc, Khi, Klo);
    GPX_CHECK_LAUNCH();
    return 0;
}

template <int KIND>
static int launch_mll_grad(const double* X, int N, int D, const double* theta, const double* Kinv, int Np,
                           const double* alpha, double* partial, double* grad, cudaStream_t st) {
    const int nb = Np / TILE;
    const size_t smem = size_t(2) * MAXD * P_LD * 8 + 2 * TILE * 8 + 8 * (MAXD + 2) * 8 + size_t(D) * P_THREADS * 8;
    cudaFuncSetAttribute(mll_grad_tile_kernel<KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    mll_grad_tile_kernel<KIND><<<dim3(nb, nb), P_THREADS, smem, st>>>(X, N, D, theta, Kinv, Np, alpha, partial);
    GPX_CHECK_LAUNCH();
    mll_grad_final_kernel<<<D + 2, 256, 0, st>>>(partial, nb, D, N, theta, grad);
    GPX_CHECK_LAUNCH();
    return 0;
}

#define GPX_DISPATCH_KIND(kind, CALL)                 \
    switch (kind) {                                   \
        case RBF: return CALL(RBF);                   \
        case MATERN12: return CALL(MATERN12);         \
        case MATERN32: return CALL(MATERN32);         \
        case MATERN52: return CALL(MATERN52);         \
        default: return -5;                           \
    }

int kmat_sym(const double* X, int N, int D, const double* theta, int kind, double* K, int Np, cudaStream_t st) {
#define C_(KD) launch_kmat_sym<KD>(X, N, D, theta, K, Np, st)
    GPX_DISPATCH_KIND(kind, C_)
#undef C_
}

int kmat_cross(const double* Xs, long m_valid, const double* X, int N, int D, const double* theta, const double* xa,
               const double* xr, int kind, const double* alpha, double* Ks, int Np, double* mean_part, int Mc,
               cudaStream_t st) {
#define C_(KD) launch_kmat_cross<KD>(Xs, m_valid, X, N, D, theta, xa, xr, alpha, Ks, Np, mean_part, Mc, nullptr, nullptr, st)
    GPX_DISPATCH_KIND(kind, C_)
#undef C_
}

// Sliced-integer mode: K* written as `nslices` int8 planes [nslices][plane_rows][Np] (gpx_i8.cu).
int kmat_cross_i8(const double* Xs, long m_valid, const double* X, int N, int D, const double* theta, const double* xa,
                  const double* xr, int kind, const double* alpha, int8_t* planes, long plane_rows, int nslices, int bits, int Np,
                  double* mean_part, int Mc, cudaStream_t st) {
#define C_(KD) launch_kmat_cross_i8<KD>(Xs, m_valid, X, N, D, theta, xa, xr, alpha, planes, plane_rows, nslices, bits, Np, mean_part, Mc, st)
    GPX_DISPATCH_KIND(kind, C_)
#undef C_
}

int predict_mean(const double* Xs, long M, const double* X, int N, int D, const double* theta, const double* xa, const double* xr,
                 int kind, const double* alpha, int Np, const double* prior_mean, const int* feat_idx, int n_feat, int degree,
                 const double* weights, const double* bias, double y_mu, double y_sigma, double* out_mean, cudaStream_t st) {
    if (M <= 0) return 0;
#define C_(KD) launch_predict_mean<KD>(Xs, M, X, N, D, theta, xa, xr, alpha, Np, prior_mean, feat_idx, n_feat, degree, weights, bias, y_mu, y_sigma, out_mean, st)
    GPX_DISPATCH_KIND(kind, C_)
#undef C_
}

// TF32 mode: K* written as (hi, lo) float pairs for the 3xTF32 tensor-core contraction.
//   fp32_eval = 1 (default of the mode): kernel evaluated in FP32;  0: evaluated in FP64, then split
int kmat_cross_split(const double* Xs, long m_valid, const double* X, int N, int D, const double* theta,
                     const double* xa, const double* xr, int kind, const double* alpha, float* Khi, float* Klo, int Np,
                     double* mean_part, int Mc, int fp32_eval, cudaStream_t st) {
    if (fp32_eval) {
#define C_(KD) launch_kmat_cross_f32<KD>(Xs, m_valid, X, N, D, theta, xa, xr, alpha, Khi, Klo, Np, mean_part, Mc, st)
        GPX_DISPATCH_KIND(kind, C_)
#undef C_
    }
#define C_(KD) launch_kmat_cross<KD>(Xs, m_valid, X, N, D, theta, xa, xr, alpha, nullptr, Np, mean_part, Mc, Khi, Klo, st)
    GPX_DISPATCH_KIND(kind, C_)
#undef C_
}

int mll_grad(const double* X, int N, int D, const double* theta, int kind, const double* Kinv, int Np,
             const double* alpha, double* partial, double* grad, cudaStream_t st) {
#define C_(KD) launch_mll_grad<KD>(X, N, D, theta, Kinv, Np, alpha, partial, grad, st)
    GPX_DISPATCH_KIND(kind, C_)
#undef C_
}

}  // namespace gpx
