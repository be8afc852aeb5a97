// gperks_b200 -- shared device helpers for the sm_100a exact-GP kernels.
//
// Layout conventions (all FP64, row-major, device memory):
//   * every N x N factor is stored with a padded leading dimension Np = ceil(N/128)*128; the
//     padding block of K_hat is the identity, so L, L^-1 and K_hat^-1 are block-diag(real, I) and
//     no kernel needs edge handling (tiles are always full 128 x 128, K loops full 16-wide).
//   * theta (device): [0,D) = 1/lengthscale_d, [D] = outputscale, [D+1] = noise.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "gpx_math.cuh"

namespace gpx {

constexpr int TILE = 128;          // tile edge of every blocked algorithm
constexpr int MAXD = 32;           // maximum input dimension supported by the pairwise kernels

enum Kind { RBF = 0, MATERN12 = 1, MATERN32 = 2, MATERN52 = 3 };

__host__ __device__ inline int round_up(int n, int m) { return (n + m - 1) / m * m; }

// ---------------------------------------------------------------------------------------------
// async copy + FP64 tensor-core MMA wrappers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem));
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

// D(8x8) += A(8x4) * B(4x8); lane holds A[lane/4][lane%4], B[lane%4][lane/4], C[lane/4][2*(lane%4)+{0,1}]
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
        : "+d"(c0), "+d"(c1)
        : "d"(a), "d"(b));
}

// ---------------------------------------------------------------------------------------------
// stationary kernel profile k(r^2) and the radial derivative factor g(r) with
//   dk/dl_d = g(r) * delta_d^2 / l_d^3           
// ---------------------------------------------------------------------------------------------
// The builders of K_hat and K* (gpx_kmat.cu) stage their coordinates pre-multiplied by coord_scale<KIND>() / lengthscale and
// start every squared-distance accumulator at 1e-30, so that q = c^2 r^2 + 1e-30 arrives ready for the profile:
//   RBF (c^2 = 1/2): s exp(-q)       Matern: u = sqrt(q) = c r;  1/2: s exp(-u);  3/2: (s + s u) exp(-u);
//   5/2: (s + s u + (s/3) q) exp(-u).   q = 1e-30 (a point against itself) gives u = 1e-15 and k = s to the last bit for nu > 1/2.
// s may be 0 (masked entry): every form then returns exactly 0.  s3 = s / 3 is formed once per thread.
template <int KIND>
__host__ __device__ constexpr double coord_scale() {
    return KIND == RBF ? 0.70710678118654752 : KIND == MATERN12 ? 1.0 : KIND == MATERN32 ? 1.7320508075688772 : 2.23606797749979;
}
template <int KIND>
__device__ __forceinline__ double kernel_profile(double q, double s, double s3) {
    if (KIND == RBF) return s * exp_neg(q);
    const double u = sqrt_pos(q);
    if (KIND == MATERN12) return s * exp_neg(u);
    if (KIND == MATERN32) return fma(u, s, s) * exp_neg(u);
    return fma(q, s3, fma(u, s, s)) * exp_neg(u);
}

template <int KIND>
__device__ __forceinline__ double kernel_value(double r2, double s) {
    if (KIND == RBF) return s * exp_neg(0.5 * r2);
    const double r = sqrt_pos(r2 + 1e-30);          // r2 = 0 (a point against itself) -> r = 1e-15, k = s to the last bit
    if (KIND == MATERN12) return s * exp_neg(r);
    if (KIND == MATERN32) {
        const double cr = 1.7320508075688772 * r;
        return (s * (1.0 + cr)) * exp_neg(cr);
    }
    const double cr = 2.23606797749979 * r;
    return (s * fma(r2, 5.0 / 3.0, 1.0 + cr)) * exp_neg(cr);
}

// returns k and writes g
template <int KIND>
__device__ __forceinline__ double kernel_value_grad(double r2, double s, double& g) {
    if (KIND == RBF) {
        double k = s * exp_neg(0.5 * r2);
        g = k;
        return k;
    }
    const double r = sqrt_pos(r2 + 1e-30);
    if (KIND == MATERN12) {
        double e = s * exp_neg(r);
        g = (r2 > 0.0) ? e / r : 0.0;
        return e;
    }
    if (KIND == MATERN32) {
        const double cr = 1.7320508075688772 * r;
        double e = s * exp_neg(cr);
        g = 3.0 * e;
        return (1.0 + cr) * e;
    }
    const double cr = 2.23606797749979 * r;
    double e = s * exp_neg(cr);
    g = (5.0 / 3.0) * (1.0 + cr) * e;
    return fma(r2, 5.0 / 3.0, 1.0 + cr) * e;
}

// ---------------------------------------------------------------------------------------------
// NT tile GEMM engine on FP64 DMMA:   acc(128x128) = sum_k A[m][k] * B[n][k]
//
// Both operands are "row = tile row, k contiguous" in global memory (every dense stage of the GP
// path is arranged to be NT -- see DESIGN.md).  256 threads = 8 warps as 2(m) x 4(n); each warp
// owns a 64 x 32 sub-tile = 8 x 4 DMMA.8x8x4 accumulators (128 registers).  K is consumed in
// 16-wide slices through a 4-stage cp.async ring (rows padded to 18 doubles => conflict-free
// LDS.128 fragment reads; lane c of an MMA takes columns 4c..4c+3 of the slice, one per k4 step).
//
// An operand may switch to an alternate source for one k-window (used to splice the masked
// diagonal block D / D^T of the triangular inverse into an otherwise mirrored matrix).
// ---------------------------------------------------------------------------------------------
constexpr int G_BK = 16;
constexpr int G_LDS = 18;
constexpr int G_STAGES = 4;
constexpr int G_THREADS = 256;
constexpr int G_STAGE_DOUBLES = 2 * TILE * G_LDS;                       // A + B per stage
constexpr size_t G_SMEM_BYTES = size_t(G_STAGES) * G_STAGE_DOUBLES * 8;  // 147456 B

struct Operand {
    const double* base;   // element (row r, k) = base[r*ld + k]        (k in absolute k units)
    long ld;
    const double* alt;    // for k in [alt_k0, alt_k1): alt[r*alt_ld + (k - alt_k0)]
    long alt_ld;
    int alt_k0, alt_k1;
};

__device__ __forceinline__ Operand make_operand(const double* base, long ld) {
    Operand o;
    o.base = base; o.ld = ld; o.alt = nullptr; o.alt_ld = 0; o.alt_k0 = 0; o.alt_k1 = 0;
    return o;
}
__device__ __forceinline__ Operand make_operand_alt(const double* base, long ld, const double* alt, long alt_ld,
                                                    int k0, int k1) {
    Operand o;
    o.base = base; o.ld = ld; o.alt = alt; o.alt_ld = alt_ld; o.alt_k0 = k0; o.alt_k1 = k1;
    return o;
}

__device__ __forceinline__ void gemm_load_stage(double* stage, const Operand& A, const Operand& B, int k0, int tid) {
    const double* a_src; long a_ld;
    if (A.alt != nullptr && k0 >= A.alt_k0 && k0 < A.alt_k1) { a_src = A.alt + (k0 - A.alt_k0); a_ld = A.alt_ld; }
    else { a_src = A.base + k0; a_ld = A.ld; }
    const double* b_src; long b_ld;
    if (B.alt != nullptr && k0 >= B.alt_k0 && k0 < B.alt_k1) { b_src = B.alt + (k0 - B.alt_k0); b_ld = B.alt_ld; }
    else { b_src = B.base + k0; b_ld = B.ld; }
    double* as = stage;
    double* bs = stage + TILE * G_LDS;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int c = tid + i * G_THREADS;       // 1024 16-byte chunks per operand
        int row = c >> 3, kc = (c & 7) * 2;
        cp_async16(as + row * G_LDS + kc, a_src + row * a_ld + kc);
        cp_async16(bs + row * G_LDS + kc, b_src + row * b_ld + kc);
    }
}

// acc[s][n][e]: rows wm*64 + s*8 + lane/4, cols wn*32 + n*8 + 2*(lane%4) + e
__device__ __forceinline__ void gemm_nt_mainloop(double (&acc)[8][4][2], const Operand& A, const Operand& B,
                                                 int k_begin, int k_end, double* smem) {
    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int wm = warp >> 2, wn = warp & 3;
    const int lr = lane >> 2, lc = lane & 3;
#pragma unroll
    for (int s = 0; s < 8; ++s)
#pragma unroll
        for (int n = 0; n < 4; ++n) { acc[s][n][0] = 0.0; acc[s][n][1] = 0.0; }

    const int nk = (k_end - k_begin) / G_BK;
#pragma unroll
    for (int s = 0; s < G_STAGES - 1; ++s) {
        if (s < nk) gemm_load_stage(smem + s * G_STAGE_DOUBLES, A, B, k_begin + s * G_BK, tid);
        cp_async_commit();
    }
    for (int kt = 0; kt < nk; ++kt) {
        cp_async_wait<G_STAGES - 2>();
        __syncthreads();
        {
            int nxt = kt + G_STAGES - 1;
            if (nxt < nk) gemm_load_stage(smem + (nxt % G_STAGES) * G_STAGE_DOUBLES, A, B, k_begin + nxt * G_BK, tid);
            cp_async_commit();
        }
        const double* as = smem + (kt % G_STAGES) * G_STAGE_DOUBLES + (wm * 64 + lr) * G_LDS + lc * 4;
        const double* bs = smem + (kt % G_STAGES) * G_STAGE_DOUBLES + TILE * G_LDS + (wn * 32 + lr) * G_LDS + lc * 4;
        double bf[4][4];
#pragma unroll
        for (int n = 0; n < 4; ++n) {
            double2 v0 = *reinterpret_cast<const double2*>(bs + n * 8 * G_LDS);
            double2 v1 = *reinterpret_cast<const double2*>(bs + n * 8 * G_LDS + 2);
            bf[n][0] = v0.x; bf[n][1] = v0.y; bf[n][2] = v1.x; bf[n][3] = v1.y;
        }
#pragma unroll
        for (int s = 0; s < 8; ++s) {
            double2 a0 = *reinterpret_cast<const double2*>(as + s * 8 * G_LDS);
            double2 a1 = *reinterpret_cast<const double2*>(as + s * 8 * G_LDS + 2);
            double af[4] = {a0.x, a0.y, a1.x, a1.y};
#pragma unroll
            for (int st = 0; st < 4; ++st)
#pragma unroll
                for (int n = 0; n < 4; ++n) dmma884(acc[s][n][0], acc[s][n][1], af[st], bf[n][st]);
        }
    }
    cp_async_wait<0>();
    __syncthreads();   // smem may be reused by the caller's epilogue
}

// Visit every accumulator element with its (row, col) inside the 128 x 128 tile.
// f(row, col, v0, v1): v0 at (row, col), v1 at (row, col+1); col is even.
template <typename F>
__device__ __forceinline__ void gemm_foreach_pair(double (&acc)[8][4][2], F f) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int wm = warp >> 2, wn = warp & 3;
    const int lr = lane >> 2, lc = lane & 3;
#pragma unroll
    for (int s = 0; s < 8; ++s)
#pragma unroll
        for (int n = 0; n < 4; ++n) f(wm * 64 + s * 8 + lr, wn * 32 + n * 8 + 2 * lc, acc[s][n][0], acc[s][n][1]);
}

// same traversal with writable accumulators (used to preload C so that acc = C - A B^T needs no read-modify-write epilogue)
template <typename F>
__device__ __forceinline__ void gemm_foreach_pair_ref(double (&acc)[8][4][2], F f) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int wm = warp >> 2, wn = warp & 3;
    const int lr = lane >> 2, lc = lane & 3;
#pragma unroll
    for (int s = 0; s < 8; ++s)
#pragma unroll
        for (int n = 0; n < 4; ++n) f(wm * 64 + s * 8 + lr, wn * 32 + n * 8 + 2 * lc, acc[s][n][0], acc[s][n][1]);
}

// ---- digit formats of the sliced-integer mode (gpx_i8.cu, gpx_kmat.cu) ----
// BITS = 8: the top balanced digit must stay <= 127, i.e. |x| * scale <= 127.4 / 256: one more bit of headroom when f > 0.9953
template <int BITS>
__device__ __forceinline__ int scale_exponent_b(double amax) {
    if (!(amax > 0.0)) return 0;
    int e;
    const double f = frexp(amax, &e);
    if (BITS == 8 && f > 0.9953) ++e;
    return e;
}
// The S balanced radix-256 digits of v (|v| <= 127.4/256), most significant first: digit q = byte S-1-q of (V + bias) ^ 0x80..80
template <int S>
__device__ __forceinline__ unsigned long long balanced_digits(double v) {
    constexpr unsigned long long BIAS = (S == 5) ? 0x8080808080ull : 0x808080808080ull;
    const long long V = __double2ll_rn(v * (double)(1ull << (8 * S)));
    return ((unsigned long long)V + BIAS) ^ BIAS;
}
__device__ __forceinline__ int perm128(int c) { return (c & ~127) | ((c & 15) << 3) | ((c & 127) >> 4); }

#define GPX_CHECK_LAUNCH()                                  \
    do {                                                    \
        cudaError_t e__ = cudaGetLastError();               \
        if (e__ != cudaSuccess) return -1000 - (int)e__;    \
    } while (0)

}  // namespace gpx
