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
#include "gpx_gemm_tma.cuh"
#include <stdlib.h>
#include <mutex>

namespace gpx {

// ------------------------------------------------------------------------------------------------
// Diagonal block, version 3 (GPX_DIAG_IMPL=3, kept as the A/B reference of v5): the whole 128 x 128 problem lives in
// REGISTERS.  Its predecessors (v1: column by column in shared memory, 186 us; v2: 8-wide sub-panels in shared memory
// with DMMA updates, 83 us, bound by shared-memory traffic -- 12 wavefronts per 2 DMMAs) are in the history only.
// Here the 16 x 16 grid of 8x8 blocks is distributed once: warp w owns
// block rows w and 15-w (17 lower-triangular blocks, perfectly balanced) of BOTH the matrix (A -> L) and the
// right-hand side of L X = I (B -> X = L^-1), each block held as a DMMA C fragment (2 doubles per lane).
// Per 8-wide panel p only the panel itself goes through shared memory:
//   (1) the owner of block (p,p) factors it with warp shuffles, publishes L_pp and T = L_pp^-1;
//   (2) owners of column-p blocks form L_ip = A_ip T^T and owners of row-p X blocks form X_pj = T B_pj on DMMA
//       (C-fragment -> operand-fragment conversion by shuffles) and publish them;
//   (3) every warp updates its own blocks: A_ij -= L_ip L_jp^T, B_ij -= L_ip X_pj (2 DMMAs per block, operands from
//       the published panel, accumulators never leave the registers).
// The published panel is double-buffered by panel parity, so only two CTA barriers per panel are needed.
// ------------------------------------------------------------------------------------------------
constexpr int D3_THREADS = 256;
constexpr int D3_BS = 96;                 // doubles per published 8x8 block: rows padded to 12 (conflict-free fragments)
constexpr int D3_BUF = 33 * D3_BS;        // 16 L blocks + 16 X blocks + T
constexpr int D3_SMEM = (2 * D3_BUF + 2 * TILE) * 8;

__device__ __forceinline__ void d3_slot(int s, int w, int& bi, int& bj) {
    if (s <= w) { bi = w; bj = s; } else { bi = 15 - w; bj = s - w - 1; }
}

// C-fragment (lane (lr,lc) holds M[lr][2lc], M[lr][2lc+1]) -> A-operand fragment (M[lr][lc], M[lr][lc+4])
__device__ __forceinline__ void d3_c_to_a(double v0, double v1, int lr, int lc, double& a0, double& a1) {
    const unsigned full = 0xffffffffu;
    const int src0 = lr * 4 + (lc >> 1), src1 = src0 + 2;
    const double p0 = __shfl_sync(full, v0, src0), p1 = __shfl_sync(full, v1, src0);
    const double q0 = __shfl_sync(full, v0, src1), q1 = __shfl_sync(full, v1, src1);
    a0 = (lc & 1) ? p1 : p0;
    a1 = (lc & 1) ? q1 : q0;
}
// C-fragment -> B-operand fragment of the same matrix used as B[k][n] = M[k][n]: lane (k=lc, n=lr) needs M[lc][lr], M[lc+4][lr]
__device__ __forceinline__ void d3_c_to_b(double v0, double v1, int lr, int lc, double& b0, double& b1) {
    const unsigned full = 0xffffffffu;
    const int src0 = lc * 4 + (lr >> 1), src1 = (lc + 4) * 4 + (lr >> 1);
    const double p0 = __shfl_sync(full, v0, src0), p1 = __shfl_sync(full, v1, src0);
    const double q0 = __shfl_sync(full, v0, src1), q1 = __shfl_sync(full, v1, src1);
    b0 = (lr & 1) ? p1 : p0;
    b1 = (lr & 1) ? q1 : q0;
}

__global__ void __launch_bounds__(D3_THREADS, 1)
potrf_diag3_kernel(double* __restrict__ K, int Np, int blk, int N, double* __restrict__ S, double* __restrict__ DT,
                   double* __restrict__ logdet_blk, int* __restrict__ info) {
    extern __shared__ double sm3[];
    double* rinv = sm3 + 2 * D3_BUF;          // [128]
    double* dg = rinv + TILE;                 // [128]
    __shared__ int bad_pivot;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int lr = lane >> 2, lc = lane & 3;
    const unsigned full = 0xffffffffu;
    const long off = (long)blk * TILE;
    double* Ab = K + off * Np + off;
    if (tid == 0) bad_pivot = 0;
#ifdef GPX_DIAG_TIMING
    __shared__ long long t3_own[3];
    if (tid == 0) { t3_own[0] = 0; t3_own[1] = 0; t3_own[2] = 0; }
    long long t3_mark = clock64(), t3_acc[5] = {0, 0, 0, 0, 0};
#define GPX_TICK3(i) do { long long n_ = clock64(); t3_acc[i] += n_ - t3_mark; t3_mark = n_; } while (0)
#else
#define GPX_TICK3(i)
#endif

    // ---- distribute: 17 blocks of the matrix per warp as C fragments; the right-hand side starts as the identity
    double cl0[17], cl1[17], cx0[17], cx1[17];
#pragma unroll
    for (int s = 0; s < 17; ++s) {
        int bi, bj;
        d3_slot(s, w, bi, bj);
        const double2 v = *reinterpret_cast<const double2*>(Ab + (long)(8 * bi + lr) * Np + 8 * bj + 2 * lc);
        cl0[s] = v.x; cl1[s] = v.y;
        cx0[s] = 0.0; cx1[s] = 0.0;
    }
    __syncthreads();
    GPX_TICK3(0);

    for (int p = 0; p < 16; ++p) {
        double* buf = sm3 + (p & 1) * D3_BUF;
        double* Lc = buf;                      // published column p of L: block bi at Lc + bi*D3_BS, [8][12]
        double* Xr = buf + 16 * D3_BS;         // published row p of X   : block bj at Xr + bj*D3_BS, stored [k][n]
        double* Tb = buf + 32 * D3_BS;         // T = L_pp^-1, [8][12], zeros above the diagonal
        const int wp = (p <= 7) ? p : 15 - p;  // owner warp of block row p
        const int sp = (p <= 7) ? p : 16;      // its slot of the diagonal block (p,p)
        // ---- (1) diagonal 8x8 block
        if (w == wp) {
#ifdef GPX_DIAG_TIMING
            const long long q0 = clock64();
#endif
            double d0 = 0.0, d1 = 0.0;
#pragma unroll
            for (int s = 0; s < 17; ++s) if (s == sp) { d0 = cl0[s]; d1 = cl1[s]; }
            double inv_r = 0.0, sd_r = 0.0;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int jq = j >> 1, je = j & 1;
                const double sel = je ? d1 : d0;
                const double piv = __shfl_sync(full, sel, j * 4 + jq);
                if (!(piv > 0.0) && lane == 0 && bad_pivot == 0) bad_pivot = 8 * p + j + 1;
                const double inv = rsqrt(piv);
                const double sd = piv * inv;
                const double lrj = __shfl_sync(full, sel, lr * 4 + jq) * inv;
                const double lc0 = __shfl_sync(full, sel, (2 * lc) * 4 + jq) * inv;
                const double lc1 = __shfl_sync(full, sel, (2 * lc + 1) * 4 + jq) * inv;
                if (lc == jq && lr >= j) {
                    const double v = (lr == j) ? sd : lrj;
                    if (je) d1 = v; else d0 = v;
                }
                if (2 * lc > j && lr >= 2 * lc) d0 = fma(-lrj, lc0, d0);
                if (2 * lc + 1 > j && lr >= 2 * lc + 1) d1 = fma(-lrj, lc1, d1);
                if (lr == j) { inv_r = inv; sd_r = sd; }
            }
            if (2 * lc > lr) d0 = 0.0;                 // keep L_pp clean above the diagonal
            if (2 * lc + 1 > lr) d1 = 0.0;
#pragma unroll
            for (int s = 0; s < 17; ++s) if (s == sp) { cl0[s] = d0; cl1[s] = d1; }
            *reinterpret_cast<double2*>(Lc + p * D3_BS + lr * 12 + 2 * lc) = make_double2(d0, d1);
            if (lc == 0) { rinv[8 * p + lr] = inv_r; dg[8 * p + lr] = sd_r; }
            __syncwarp();
#ifdef GPX_DIAG_TIMING
            const long long q1 = clock64();
#endif
            if (lane < 8) {        // column `lane` of T by forward substitution
                const double* Lb = Lc + p * D3_BS;
                double x[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    double acc_ = (i == lane) ? 1.0 : 0.0;
#pragma unroll
                    for (int k = 0; k < i; ++k) acc_ = fma(-Lb[i * 12 + k], x[k], acc_);
                    x[i] = acc_ * rinv[8 * p + i];
                }
#pragma unroll
                for (int i = 0; i < 8; ++i) Tb[i * 12 + lane] = x[i];
            }
            __syncwarp();
            // X_pp = T: the owner's right-hand-side diagonal block, also published as row-p block p ([k][n] = T[k][n])
            const double2 t = *reinterpret_cast<const double2*>(Tb + lr * 12 + 2 * lc);
#pragma unroll
            for (int s = 0; s < 17; ++s) if (s == sp) { cx0[s] = t.x; cx1[s] = t.y; }
            *reinterpret_cast<double2*>(Xr + p * D3_BS + lr * 12 + 2 * lc) = t;
#ifdef GPX_DIAG_TIMING
            if (lane == 0) { const long long q2 = clock64(); t3_own[0] += q1 - q0; t3_own[1] += q2 - q1; }
#endif
        }
        __syncthreads();
        GPX_TICK3(1);
        // ---- (2) column-p blocks of L (bi > p) and row-p blocks of X (bj < p) on DMMA
        {
            const double tb0 = Tb[lr * 12 + lc], tb1 = Tb[lr * 12 + lc + 4];     // T as A operand, and T^T as B operand
#pragma unroll
            for (int s = 0; s < 17; ++s) {
                int bi, bj;
                d3_slot(s, w, bi, bj);
                if (bj == p && bi > p) {          // L_ip = A_ip T^T : A = A_ip (from registers), B[k][n] = T[n][k]
                    double a0, a1;
                    d3_c_to_a(cl0[s], cl1[s], lr, lc, a0, a1);
                    double r0 = 0.0, r1 = 0.0;
                    dmma884(r0, r1, a0, tb0);
                    dmma884(r0, r1, a1, tb1);
                    cl0[s] = r0; cl1[s] = r1;
                    *reinterpret_cast<double2*>(Lc + bi * D3_BS + lr * 12 + 2 * lc) = make_double2(r0, r1);
                }
                if (bi == p && bj < p) {          // X_pj = T B_pj : A = T, B[k][n] = B_pj[k][n] (from registers)
                    double b0, b1;
                    d3_c_to_b(cx0[s], cx1[s], lr, lc, b0, b1);
                    double r0 = 0.0, r1 = 0.0;
                    dmma884(r0, r1, tb0, b0);
                    dmma884(r0, r1, tb1, b1);
                    cx0[s] = r0; cx1[s] = r1;
                    *reinterpret_cast<double2*>(Xr + bj * D3_BS + lr * 12 + 2 * lc) = make_double2(r0, r1);
                }
            }
        }
        __syncthreads();
        GPX_TICK3(2);
        // ---- (3) register-resident rank-8 updates of this warp's two block rows
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const int bi = half ? 15 - w : w;
            if (bi > p) {
                const double a0 = -Lc[bi * D3_BS + lr * 12 + lc], a1 = -Lc[bi * D3_BS + lr * 12 + lc + 4];
#pragma unroll
                for (int s = 0; s < 17; ++s) {
                    int sbi, bj;
                    d3_slot(s, w, sbi, bj);
                    if (sbi == bi) {
                        if (bj > p) {             // A_ij -= L_ip L_jp^T : B[k][n] = L_jp[n][k]
                            const double b0 = Lc[bj * D3_BS + lr * 12 + lc], b1 = Lc[bj * D3_BS + lr * 12 + lc + 4];
                            dmma884(cl0[s], cl1[s], a0, b0);
                            dmma884(cl0[s], cl1[s], a1, b1);
                        } else {                  // B_ij -= L_ip X_pj : B[k][n] = X_pj[k][n]
                            const double b0 = Xr[bj * D3_BS + lc * 12 + lr], b1 = Xr[bj * D3_BS + (lc + 4) * 12 + lr];
                            dmma884(cx0[s], cx1[s], a0, b0);
                            dmma884(cx0[s], cx1[s], a1, b1);
                        }
                    }
                }
            }
        }
        GPX_TICK3(3);
        // no barrier: the next panel publishes into the other buffer, and every block a warp needs next is its own
    }
    __syncthreads();
    if (tid == 0 && bad_pivot != 0 && *info == 0) *info = (int)off + bad_pivot;
    // ---- write back from the registers: L -> K (lower blocks), X -> S (lower blocks), X^T -> DT
    double* Sb = S + off * Np + off;
    double* DTb = DT + (long)blk * TILE * TILE;
#pragma unroll
    for (int s = 0; s < 17; ++s) {
        int bi, bj;
        d3_slot(s, w, bi, bj);
        const long r = 8 * bi + lr, c = 8 * bj + 2 * lc;
        *reinterpret_cast<double2*>(Ab + r * Np + c) = make_double2(cl0[s], cl1[s]);
        *reinterpret_cast<double2*>(Sb + r * Np + c) = make_double2(cx0[s], cx1[s]);
        DTb[c * TILE + r] = cx0[s];
        DTb[(c + 1) * TILE + r] = cx1[s];
    }
    // zero the blocks above the diagonal of S and below the diagonal of DT (diagonal blocks are already clean)
    for (int e = tid; e < 120 * 32; e += D3_THREADS) {
        const int b = e >> 5, q = e & 31;               // b-th strictly-upper block, q-th double2 of it
        int bi = 0, rem = b;
        while (rem >= 15 - bi) { rem -= 15 - bi; ++bi; }
        const int bj = bi + 1 + rem;                    // bj > bi
        const int rr = q >> 2, cc = (q & 3) * 2;
        *reinterpret_cast<double2*>(Sb + (long)(8 * bi + rr) * Np + 8 * bj + cc) = make_double2(0.0, 0.0);
        *reinterpret_cast<double2*>(DTb + (8 * bj + rr) * TILE + 8 * bi + cc) = make_double2(0.0, 0.0);
    }
#ifdef GPX_DIAG_TIMING
    __syncthreads();
    GPX_TICK3(4);
    if (tid == 0) {     // timing build only
        for (int i = 0; i < 5; ++i) logdet_blk[8 + i] = (double)t3_acc[i];
        for (int i = 0; i < 2; ++i) logdet_blk[20 + i] = (double)t3_own[i];
    }
#endif
    if (tid < 32) {
        double t = 0.0;
        for (int j = tid; j < TILE; j += 32)
            if (off + j < N) t += log(dg[j]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if (tid == 0) logdet_blk[blk] = t;
    }
}

// ------------------------------------------------------------------------------------------------
// Diagonal block, version 5 (default): v3's register-resident block distribution, restructured around its measured
// critical path (tools/diag_timing.cu: of v3's 123k cycles, 33k were the serial 8x8 factors done with shuffles and ~14k the
// owner's serial X-row finalisation; the 75 KB unrolled loop body missed the instruction cache).
//   * ONE accumulator per block slot: block (i,j) is first the Schur-updated A_ij (panels p < j), is finalised to
//     L_ij at panel j and written out, and then restarts as the right-hand side B_ij of L X = I (panels j <= p < i)
//     until it is finalised to X_ij at panel i.
//   * A dedicated factor warp does nothing but factor the 8x8 diagonal blocks: every lane runs the whole 8x8 LDL^T
//     recurrence redundantly in registers (no shuffles; the pivots chain through a branch-free reciprocal, the rsqrt's
//     are off the chain), and lane c forms column c of T = L_pp^-1.  The owner of block row p+1 updates block (p+1,p+1)
//     FIRST in panel p's update stage and hands it over through a named barrier, so the factor of panel p+1 overlaps
//     the rank-8 updates of panel p.
//   * The right-hand-side updates use M_ip = L_ip T_p (one extra DMMA pair per block row, computed by the row's own
//     warp) against the RAW row B_pj, i.e. B_ij -= (L_ip T_p) B_pj: the owner of row p only has to publish registers,
//     and the outputs X_pj = T_p B_pj are formed from the published blocks by all warps, after the updates.
//   * The 16 panels are unrolled at compile time (template <int P>): a rolled version issued 930 warp-instructions per
//     panel and warp (ncu, profiles/r01_ncu_potrf_diag.txt), ~3x what the arithmetic needs -- select chains that pick
//     "slot p" out of the register arrays, per-slot predicates.  With P static every slot index is an immediate, the
//     phase of a slot is known at compile time and dead slots are never visited (260 instructions).
//   * Published blocks are stored in FRAGMENT order (element `lane` of a block is the pair a DMMA operand needs), so an
//     operand is one conflict-free LDS.128 and the C-fragment -> operand conversions go through a small per-warp
//     shared-memory transposer instead of eight shuffles.
// ------------------------------------------------------------------------------------------------
constexpr int D5_DS = 96;                 // doubles per row-major 8x8 block with rows padded to 12 (factor warp's in/out)
__device__ __forceinline__ void d5_bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void d5_bar_arrive(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }
constexpr int D5_THREADS = 384;           // 12 warps: warp 0 factors alone on its scheduler, 8 compute warps on the other three
constexpr int D5_BS = 64;                 // doubles per published 8x8 block
constexpr int D5_BUF = 34 * D5_BS;        // 16 L blocks | 16 raw right-hand-side blocks | T (A layout) | T (B layout)
constexpr int D5_SMEM = (2 * D5_BUF + 3 * D5_DS + TILE + 16 * D5_BS) * 8;

// index of element (r, c) of an 8x8 block in "A layout": lane lr*4+lc finds (M[lr][lc], M[lr][lc+4]) as one double2.
// The same pair is the B operand of M^T (B[k][n] = M[n][k]).
__device__ __forceinline__ int d5_ia(int r, int c) { return ((r * 4 + (c & 3)) << 1) + (c >> 2); }
// "B layout": lane lr*4+lc finds (M[lc][lr], M[lc+4][lr]), the B operand of M itself (B[k][n] = M[k][n])
__device__ __forceinline__ int d5_ib(int r, int c) { return ((c * 4 + (r & 3)) << 1) + (r >> 2); }

// Reciprocal and reciprocal square root of a positive, normal double without the library's special-case branches: the
// hardware seed (MUFU.RCP64H / RSQ64H, ~2^-20) followed by one third-order step (error e^3) and, for the rsqrt (off the
// pivot chain), a final Newton step.  Branch-free, so the whole 8x8 recurrence is one basic block whose independent
// chains the scheduler interleaves.  Non-positive pivots are reported through `info` before these are used.
__device__ __forceinline__ double d5_rcp(double d) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
    const double e = fma(-d, r, 1.0);
    const double e2 = fma(e, e, e);
    return fma(e2, r, r);
}
__device__ __forceinline__ double d5_rsqrt(double d) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
    const double e = fma(-(d * y), y, 1.0);                 // 1 - d y^2
    y = fma(y * e, fma(0.375, e, 0.5), y);                  // y (1 + e/2 + 3 e^2/8)
    const double e2 = fma(-(d * y), y, 1.0);
    return fma(0.5 * y, e2, y);
}

struct D5Ctx {
    double* sm;            // dynamic shared memory base
    double* scratch;       // this warp's transposer: [2 rows][64]
    double* Ab; double* Sb; double* DTb;
    int Np, w, lane, lr, lc;
    int ia0, ib0;          // A-/B-layout index of this lane's first C-fragment element (the second is +2 / +8)
#ifdef GPX_DIAG_TIMING
    long long* tstat;      // shared: [0] = mark, [1..] = accumulated cycles per phase (thread 0 only)
#endif
};
#ifdef GPX_DIAG_TIMING
#define GPX_TICK5(X, i) do { if ((X).w == 0 && (X).lane == 0) { long long n_ = clock64(); (X).tstat[1 + (i)] += n_ - (X).tstat[0]; (X).tstat[0] = n_; } } while (0)
#else
#define GPX_TICK5(X, i)
#endif

// C fragment -> A-operand pair through the warp's transposer
__device__ __forceinline__ double2 d5_c_to_a(double v0, double v1, double* scr, const D5Ctx& X) {
    __syncwarp();
    scr[X.ia0] = v0; scr[X.ia0 + 2] = v1;
    __syncwarp();
    return *reinterpret_cast<const double2*>(scr + 2 * X.lane);
}

// the owner of row P publishes its raw right-hand-side blocks (B layout)
template <int P, int NJ>
__device__ __forceinline__ void d5_publish_row(const double (&c0)[NJ], const double (&c1)[NJ], int bi, const D5Ctx& X, double* Xr) {
    if (bi != P) return;
#pragma unroll
    for (int j = 0; j < P && j < NJ; ++j) {
        double* Xb = Xr + j * D5_BS;
        Xb[X.ib0] = c0[j]; Xb[X.ib0 + 8] = c1[j];
    }
}

template <int P, int NJ>
__device__ __forceinline__ bool d5_next_diag_row(double (&c0)[NJ], double (&c1)[NJ], int bi, const D5Ctx& X, double2 aL, double* Dnext) {
    if (P + 1 >= NJ) return false;
    constexpr int QJ = (P + 1 < NJ) ? P + 1 : 0;
    if (bi != P + 1) return false;
    dmma884(c0[QJ], c1[QJ], aL.x, -aL.x);              // A_qq -= L_qp L_qp^T
    dmma884(c0[QJ], c1[QJ], aL.y, -aL.y);
    *reinterpret_cast<double2*>(Dnext + X.lr * 12 + 2 * X.lc) = make_double2(c0[QJ], c1[QJ]);
    return true;
}

template <int P, int NJ, bool INV>
__device__ __forceinline__ void d5_stage3_row(double (&c0)[NJ], double (&c1)[NJ], int bi, const D5Ctx& X,
                                              const double* Lc, const double* Xr, double2 aL, double2 aM) {
    if (bi <= P) return;
    const double* xb = Xr + 2 * X.lane;
    const double* lb = Lc + 2 * X.lane;
    // eight slots at a time: operand loads, then the first k-step of every slot, then the second (consecutive DMMAs
    // are independent).  Slots j < P are right-hand sides, B_ij -= M_ip B_pj(raw), always present because j < P < bi;
    // slots j > P are Schur updates, A_ij -= L_ip L_jp^T, present if j <= bi.
#pragma unroll
    for (int j0 = 0; j0 < NJ; j0 += 8) {
        double2 b[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int j = j0 + q;
            if (INV && j < P) b[q] = *reinterpret_cast<const double2*>(xb + j * D5_BS);
            else if (j > P && j <= bi) b[q] = *reinterpret_cast<const double2*>(lb + j * D5_BS);
        }
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int j = j0 + q;
            if (INV && j < P) dmma884(c0[j], c1[j], aM.x, b[q].x);
            else if (j > P && j <= bi) dmma884(c0[j], c1[j], aL.x, b[q].x);
        }
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int j = j0 + q;
            if (INV && j < P) dmma884(c0[j], c1[j], aM.y, b[q].y);
            else if (j > P && j <= bi) dmma884(c0[j], c1[j], aL.y, b[q].y);
        }
    }
}

// index of block (i, j), j <= i, in the packed lower-triangular block list that the Cholesky-only edition exports
__device__ __forceinline__ int d5_tri(int i, int j) { return i * (i + 1) / 2 + j; }

template <int P, bool INV>
__device__ __forceinline__ void d5_panel(double (&cA0)[8], double (&cA1)[8], double (&cB0)[16], double (&cB1)[16], const D5Ctx& X) {
    double* buf = X.sm + (P & 1) * D5_BUF;
    double* Lc = buf;
    double* Xr = buf + 16 * D5_BS;
    const double* TA = buf + 32 * D5_BS;
    const double* TK = buf + 33 * D5_BS;
    double* Dnext = X.sm + 2 * D5_BUF + ((P + 1) & 1) * D5_DS;
    const int rA = X.w, rB = 15 - X.w;
    constexpr bool HAS_A = (P < 7);                    // block row w <= 7 can only be below panel P if P < 7
    constexpr int PA = HAS_A ? P : 0;
    const bool onA = HAS_A && rA > P, onB = rB > P;
    double* scrA = X.scratch;
    double* scrB = X.scratch + D5_BS;
    double* LbA = Lc + rA * D5_BS;
    double* LbB = Lc + rB * D5_BS;
    __syncthreads();                                   // #1: T_P published by the factor warp
    GPX_TICK5(X, 0);
    const double2 ta = *reinterpret_cast<const double2*>(TA + 2 * X.lane);
    const double2 tk = *reinterpret_cast<const double2*>(TK + 2 * X.lane);
    double2 lA = make_double2(0.0, 0.0), lB = lA;
    // ---- stage 2a: L_ip = A_ip T^T for this warp's (<= 2) blocks of column P.  Both rows run as one straight-line
    //      sequence (an absent row computes on garbage and commits nothing), so their transposer round trips and
    //      DMMA pairs overlap.
    if (onA | onB) {
        if (HAS_A && onA) { scrA[X.ia0] = cA0[PA]; scrA[X.ia0 + 2] = cA1[PA]; }
        if (onB) { scrB[X.ia0] = cB0[P]; scrB[X.ia0 + 2] = cB1[P]; }
        __syncwarp();
        double2 aA = make_double2(0.0, 0.0);
        if (HAS_A) aA = *reinterpret_cast<const double2*>(scrA + 2 * X.lane);
        const double2 aB = *reinterpret_cast<const double2*>(scrB + 2 * X.lane);
        double rA0 = 0.0, rA1 = 0.0, rB0 = 0.0, rB1 = 0.0;
        if (HAS_A) dmma884(rA0, rA1, aA.x, ta.x);
        dmma884(rB0, rB1, aB.x, ta.x);
        if (HAS_A) dmma884(rA0, rA1, aA.y, ta.y);
        dmma884(rB0, rB1, aB.y, ta.y);
        if (HAS_A && onA) {
            LbA[X.ia0] = rA0; LbA[X.ia0 + 2] = rA1;
            *reinterpret_cast<double2*>(X.Ab + (long)(8 * rA + X.lr) * X.Np + 8 * P + 2 * X.lc) = make_double2(rA0, rA1);
        }
        if (onB) {
            LbB[X.ia0] = rB0; LbB[X.ia0 + 2] = rB1;
            *reinterpret_cast<double2*>(X.Ab + (long)(8 * rB + X.lr) * X.Np + 8 * P + 2 * X.lc) = make_double2(rB0, rB1);
        }
        __syncwarp();
        if (HAS_A && onA) lA = *reinterpret_cast<const double2*>(LbA + 2 * X.lane);
        if (onB) lB = *reinterpret_cast<const double2*>(LbB + 2 * X.lane);
        if (!INV) {     // export the finished blocks in operand layout for the blocked TRSM panel (scratch = this block's DT)
            if (HAS_A && onA) *reinterpret_cast<double2*>(X.DTb + d5_tri(rA, P) * D5_BS + 2 * X.lane) = lA;
            if (onB) *reinterpret_cast<double2*>(X.DTb + d5_tri(rB, P) * D5_BS + 2 * X.lane) = lB;
        }
    }
    if (INV) {
        d5_publish_row<P, 8>(cA0, cA1, rA, X, Xr);
        d5_publish_row<P, 16>(cB0, cB1, rB, X, Xr);
    }
    GPX_TICK5(X, 1);
    __syncthreads();                                   // #2: column P of L and raw row P complete
    GPX_TICK5(X, 2);
    if (P < 15) {
        const double2 aLA = make_double2(-lA.x, -lA.y), aLB = make_double2(-lB.x, -lB.y);
        // ---- the next diagonal block first: its factor then overlaps everything below
        if (d5_next_diag_row<P, 8>(cA0, cA1, rA, X, aLA, Dnext) | d5_next_diag_row<P, 16>(cB0, cB1, rB, X, aLB, Dnext))
            d5_bar_arrive(1, 64);
        GPX_TICK5(X, 3);
        // ---- stage 2b: M_ip = L_ip T; the slot restarts as B_ip = 0 - L_ip X_pp; A fragments of -M_ip
        double2 aMA = make_double2(0.0, 0.0), aMB = aMA;
        if (INV && (onA | onB)) {
            double mA0 = 0.0, mA1 = 0.0, mB0 = 0.0, mB1 = 0.0;
            if (HAS_A) dmma884(mA0, mA1, lA.x, tk.x);
            dmma884(mB0, mB1, lB.x, tk.x);
            if (HAS_A) dmma884(mA0, mA1, lA.y, tk.y);
            dmma884(mB0, mB1, lB.y, tk.y);
            if (HAS_A && onA) { cA0[PA] = -mA0; cA1[PA] = -mA1; scrA[X.ia0] = mA0; scrA[X.ia0 + 2] = mA1; }
            if (onB) { cB0[P] = -mB0; cB1[P] = -mB1; scrB[X.ia0] = mB0; scrB[X.ia0 + 2] = mB1; }
            __syncwarp();
            if (HAS_A) { const double2 n = *reinterpret_cast<const double2*>(scrA + 2 * X.lane); aMA = make_double2(-n.x, -n.y); }
            { const double2 n = *reinterpret_cast<const double2*>(scrB + 2 * X.lane); aMB = make_double2(-n.x, -n.y); }
            __syncwarp();
        }
        // ---- stage 3: rank-8 updates of this warp's two block rows
        if (HAS_A) d5_stage3_row<P, 8, INV>(cA0, cA1, rA, X, Lc, Xr, aLA, aMA);
        d5_stage3_row<P, 16, INV>(cB0, cB1, rB, X, Lc, Xr, aLB, aMB);
        GPX_TICK5(X, 4);
    }
    // output row P of the inverse, X_pj = T_p B_pj, from the published raw blocks; dealt round-robin to the warps
    for (int j = X.w; INV && j < P; j += 8) {
        const double2 b = *reinterpret_cast<const double2*>(Xr + j * D5_BS + 2 * X.lane);
        double r0 = 0.0, r1 = 0.0;
        dmma884(r0, r1, ta.x, b.x);
        dmma884(r0, r1, ta.y, b.y);
        const long r_ = 8 * P + X.lr, c_ = 8 * j + 2 * X.lc;
        *reinterpret_cast<double2*>(X.Sb + r_ * X.Np + c_) = make_double2(r0, r1);
        X.DTb[c_ * TILE + r_] = r0;
        X.DTb[(c_ + 1) * TILE + r_] = r1;
    }
    GPX_TICK5(X, 5);
}

template <int P, bool INV>
struct D5Unroll {
    static __device__ __forceinline__ void run(double (&cA0)[8], double (&cA1)[8], double (&cB0)[16], double (&cB1)[16], const D5Ctx& X) {
        D5Unroll<P - 1, INV>::run(cA0, cA1, cB0, cB1, X);
        d5_panel<P, INV>(cA0, cA1, cB0, cB1, X);
    }
};
template <bool INV>
struct D5Unroll<-1, INV> {
    static __device__ __forceinline__ void run(double (&)[8], double (&)[8], double (&)[16], double (&)[16], const D5Ctx&) {}
};

// INV = true : L_pp -> K, X_pp = L_pp^-1 -> S and (transposed) DT.
// INV = false: Cholesky only (the right-hand-side half of the rank-8 updates disappears): L_pp -> K, and the 8x8 blocks
//              of L_pp plus the sixteen diagonal 8x8 inverses T_j in DMMA-operand layout -> this block's DT storage,
//              used as scratch by trsm_strips_kernel (blocked TRSM panel; it also forms S / DT at the end of potrf).
template <bool INV>
__global__ void __maxnreg__(168)   // 9 warps are allocated as 12: 168 x 384 = 64512 registers
potrf_diag5_kernel(double* __restrict__ K, int Np, int blk, int N, double* __restrict__ S, double* __restrict__ DT,
                   double* __restrict__ logdet_blk, int* __restrict__ info) {
    extern __shared__ double sm5[];
    double* Draw = sm5 + 2 * D5_BUF;          // [2][8][12] row-major: diagonal block handed to the factor warp
    double* Lpp = Draw + 2 * D5_DS;           // [8][12]
    double* dg = Lpp + D5_DS;                 // [128]
    double* scratch = dg + TILE;              // [8 warps][2][64]
    // The factor warp's ~250 dependent FP64 instructions per panel share the FP64 pipe of their scheduler with the
    // DMMAs of whatever else runs there (measured: 2150 instead of ~600 cycles next to two compute warps), so hardware
    // warp 0 factors alone on scheduler 0 and the compute warps live on schedulers 1-3 (warps 4, 8, 11 exit).
    const int tid = threadIdx.x, lane = tid & 31, hw = tid >> 5;
    if (hw == 4 || hw == 8 || hw == 11) return;
    const int w = (hw == 0) ? 8 : hw - 1 - (hw > 4) - (hw > 8);     // 8 = factor warp, 0..7 = compute warps
    const int lr = lane >> 2, lc = lane & 3;
    const long off = (long)blk * TILE;
    double* Ab = K + off * Np + off;
    double* Sb = S + off * Np + off;
    double* DTb = DT + (long)blk * TILE * TILE;

#ifdef GPX_DIAG_TIMING
#define GPX_TICKF(i) do { long long n_ = clock64(); f_acc[i] += n_ - f_mark; f_mark = n_; } while (0)
#else
#define GPX_TICKF(i)
#endif
#ifdef GPX_DIAG_TIMING
    __shared__ long long t5_stat[8];
    if (tid == 32) { for (int i = 0; i < 8; ++i) t5_stat[i] = 0; t5_stat[0] = clock64(); }
#endif
    if (w == 8) {
        // =================== factor warp (T is published in both operand layouts) ===================
        int bad = 0;
#ifdef GPX_DIAG_TIMING
        long long f_mark = clock64(), f_acc[6] = {0, 0, 0, 0, 0, 0};
#endif
        for (int p = 0; p < 16; ++p) {
            double* TA = sm5 + (p & 1) * D5_BUF + 32 * D5_BS;
            double* TK = TA + D5_BS;
            const double* Dr = Draw + (p & 1) * D5_DS;
            d5_bar_sync(1, 64);                               // block (p,p) published by the owner of row p
            GPX_TICKF(0);
            double a[8][8];
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j <= i; j += 2) {
                    const double2 v = *reinterpret_cast<const double2*>(Dr + i * 12 + j);
                    a[i][j] = v.x;
                    if (j + 1 <= i) a[i][j + 1] = v.y;
                }
            double inv[8], sd[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const double d = a[j][j];
                if (!(d > 0.0) && bad == 0) bad = 8 * p + j + 1;
                const double r = d5_rcp(d);
                inv[j] = d5_rsqrt(d);
                sd[j] = d * inv[j];
                double uj[8];
#pragma unroll
                for (int k = j + 1; k < 8; ++k) uj[k] = a[k][j] * r;
#pragma unroll
                for (int k = j + 1; k < 8; ++k)
#pragma unroll
                    for (int i = k; i < 8; ++i) a[i][k] = fma(-a[i][j], uj[k], a[i][k]);
#pragma unroll
                for (int k = j + 1; k < 8; ++k) a[k][j] = uj[k];
            }
            GPX_TICKF(1);
            const int c = lane & 7;
            double x[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                double acc_ = (i == c) ? 1.0 : 0.0;
#pragma unroll
                for (int k = 0; k < i; ++k) acc_ = fma(-a[i][k], x[k], acc_);
                x[i] = acc_;
            }
            if (lane < 8) {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const double t = x[i] * inv[i];
                    TA[d5_ia(i, c)] = t;
                    if (INV) TK[d5_ib(i, c)] = t;            // T as a B operand is only needed for M = L T
                }
            }
            __syncwarp();
            GPX_TICKF(3);
            __syncthreads();                                  // #1: T_p visible to the compute warps
            GPX_TICKF(4);
            // L_pp itself is only needed for the outputs below: formed after the barrier, off the critical path
            if (lane == 8) {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
#pragma unroll
                    for (int j = 0; j < 8; j += 2) {
                        const double v0 = (j < i) ? a[i][j] * sd[j] : ((j == i) ? sd[i] : 0.0);
                        const double v1 = (j + 1 < i) ? a[i][j + 1] * sd[j + 1] : ((j + 1 == i) ? sd[i] : 0.0);
                        *reinterpret_cast<double2*>(Lpp + i * 12 + j) = make_double2(v0, v1);
                    }
                    dg[8 * p + i] = sd[i];
                }
            }
            __syncwarp();
            {   // outputs of the diagonal 8x8 block: L_pp -> K, T -> S and (transposed) DT
                const double2 lf = *reinterpret_cast<const double2*>(Lpp + lr * 12 + 2 * lc);
                const int i0 = d5_ia(lr, 2 * lc);
                const double t0 = TA[i0], t1 = TA[i0 + 2];
                const long r_ = 8 * p + lr, c_ = 8 * p + 2 * lc;
                *reinterpret_cast<double2*>(Ab + r_ * Np + c_) = lf;
                if (INV) {
                    *reinterpret_cast<double2*>(Sb + r_ * Np + c_) = make_double2(t0, t1);
                    DTb[c_ * TILE + r_] = t0;
                    DTb[(c_ + 1) * TILE + r_] = t1;
                } else {    // scratch: L_pp block and T_p in operand (A) layout
                    const double2 la = make_double2(Lpp[lr * 12 + lc], Lpp[lr * 12 + lc + 4]);
                    *reinterpret_cast<double2*>(DTb + d5_tri(p, p) * D5_BS + 2 * lane) = la;
                    *reinterpret_cast<double2*>(DTb + (136 + p) * D5_BS + 2 * lane) = *reinterpret_cast<const double2*>(TA + 2 * lane);
                }
            }
            __syncthreads();                                  // #2
            GPX_TICKF(5);
        }
#ifdef GPX_DIAG_TIMING
        if (lane == 0) for (int i = 0; i < 6; ++i) logdet_blk[40 + i] = (double)f_acc[i];
#endif
        if (lane == 0 && bad != 0 && *info == 0) *info = (int)off + bad;
        __syncwarp();
        double t = 0.0;
        for (int j = lane; j < TILE; j += 32)
            if (off + j < N) t += log(dg[j]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if (lane == 0) logdet_blk[blk] = t;
        return;
    }

    // =================== compute warps ===================
    D5Ctx X;
    X.sm = sm5; X.scratch = scratch + w * 2 * D5_BS;
    X.Ab = Ab; X.Sb = Sb; X.DTb = DTb; X.Np = Np; X.w = w; X.lane = lane; X.lr = lr; X.lc = lc;
    X.ia0 = d5_ia(lr, 2 * lc); X.ib0 = d5_ib(lr, 2 * lc);
#ifdef GPX_DIAG_TIMING
    X.tstat = t5_stat;
#endif
    double cA0[8], cA1[8], cB0[16], cB1[16];
    const int rA = w, rB = 15 - w;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        cA0[j] = 0.0; cA1[j] = 0.0;
        if (j <= rA) {
            const double2 v = *reinterpret_cast<const double2*>(Ab + (long)(8 * rA + lr) * Np + 8 * j + 2 * lc);
            cA0[j] = v.x; cA1[j] = v.y;
        }
    }
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        cB0[j] = 0.0; cB1[j] = 0.0;
        if (j <= rB) {
            const double2 v = *reinterpret_cast<const double2*>(Ab + (long)(8 * rB + lr) * Np + 8 * j + 2 * lc);
            cB0[j] = v.x; cB1[j] = v.y;
        }
    }
    if (w == 0) {       // block (0,0) goes to the factor warp straight away
        *reinterpret_cast<double2*>(Draw + lr * 12 + 2 * lc) = make_double2(cA0[0], cA1[0]);
        d5_bar_arrive(1, 64);        // bar.arrive/bar.sync order the shared-memory hand-over (PTX producer/consumer idiom)
    }
    // while the first factor runs: zero the blocks above the diagonal of S and below the diagonal of DT
    for (int b = w; INV && b < 256; b += 8) {
        const int bi = b >> 4, bj = b & 15;
        if (bj > bi) {
            const int rr = lane >> 2, cc = (lane & 3) * 2;
            *reinterpret_cast<double2*>(Sb + (long)(8 * bi + rr) * Np + 8 * bj + cc) = make_double2(0.0, 0.0);
            *reinterpret_cast<double2*>(DTb + (8 * bj + rr) * TILE + 8 * bi + cc) = make_double2(0.0, 0.0);
        }
    }
    D5Unroll<15, INV>::run(cA0, cA1, cB0, cB1, X);
#ifdef GPX_DIAG_TIMING
    if (w == 0 && lane == 0) for (int i = 0; i < 6; ++i) logdet_blk[48 + i] = (double)t5_stat[1 + i];
#endif
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

// Trailing update: A[i][j] -= L[i][p] L[j][p]^T for tiles tj = j0 + bx <= ti = i0 + by.
//   next block column only : (j0, i0) = (p+1, p+1), grid (1, n)      -- on the critical path
//   the rest               : (j0, i0) = (p+2, p+2), grid (n-1, n-1)  -- overlapped with the next diag/panel
__global__ void __launch_bounds__(G_THREADS, 1) potrf_trailing_kernel(double* __restrict__ K, int Np, int p, int j0,
                                                                      int i0) {
    const int ti = i0 + blockIdx.y, tj = j0 + blockIdx.x;
    if (tj > ti) return;
    extern __shared__ double smem[];
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
// TMA editions of the dense stages (default; GPX_CHOL_IMPL=cpasync selects the LDGSTS kernels above).
// Tensor maps: tmK = K[Np][Np], tmS = S[Np][Np], tmW = W[Np][Np], tmD = DT viewed as [nb*128][128].
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(GT_THREADS, 1)
potrf_panel_tma_kernel(const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmS,
                       double* __restrict__ K, int Np, int p) {
    extern __shared__ unsigned char smem_raw[];
    const int ti = p + 1 + blockIdx.x;
    TOperand A = t_operand(&tmK, ti * TILE, p * TILE);
    TOperand B = t_operand(&tmS, p * TILE, p * TILE);
    double acc[8][4][2];
    gemm_nt_tma_mainloop(acc, A, B, 0, TILE, smem_raw);
    // every thread must be past its last read of the A tile (== this output tile) before anyone overwrites it:
    // the reads went through the async proxy into smem and were all consumed above, so a CTA barrier suffices
    __syncthreads();
    double* Ct = K + (long)ti * TILE * Np + (long)p * TILE;
    gemm_foreach_pair(acc, [&](int row, int col, double v0, double v1) {
        *reinterpret_cast<double2*>(Ct + (long)row * Np + col) = make_double2(v0, v1);
    });
}

__global__ void __launch_bounds__(GT_THREADS, 1)
potrf_trailing_tma_kernel(const __grid_constant__ CUtensorMap tmK, double* __restrict__ K, int Np, int p, int j0, int i0,
                          int kpanels) {
    // C[ti][tj] -= L[ti][p .. p+kpanels) L[tj][p .. p+kpanels)^T
    const int ti = i0 + blockIdx.y, tj = j0 + blockIdx.x;
    if (tj > ti) return;
    extern __shared__ unsigned char smem_raw[];
    TOperand A = t_operand(&tmK, ti * TILE, p * TILE);
    TOperand B = t_operand(&tmK, tj * TILE, p * TILE);
    // the accumulators start as the C tile (its loads overlap the TMA prologue) and the loop accumulates -A B^T, so the
    // epilogue is a plain store instead of an exposed load-subtract-store pass
    double* Ct = K + (long)ti * TILE * Np + (long)tj * TILE;
    double acc[8][4][2];
    gemm_foreach_pair_ref(acc, [&](int row, int col, double& v0, double& v1) {
        const double2 c = *reinterpret_cast<const double2*>(Ct + (long)row * Np + col);
        v0 = c.x; v1 = c.y;
    });
    gemm_nt_tma_mainloop<false, true>(acc, A, B, 0, kpanels * TILE, smem_raw);
    gemm_foreach_pair(acc, [&](int row, int col, double v0, double v1) {
        *reinterpret_cast<double2*>(Ct + (long)row * Np + col) = make_double2(v0, v1);
    });
}

// ------------------------------------------------------------------------------------------------
// Blocked TRSM against one factored diagonal block:  Y L_pp^T = A,  8 rows per warp, 8x8 blocks, using the sixteen
// diagonal 8x8 inverses T_j the Cholesky-only diagonal kernel exported (with L_pp's blocks, in DMMA-operand layout,
// into that block's DT storage):
//     for kk = 0..15:  Y_kk = A_kk T_kk^T ;  A_j -= Y_kk L(j,kk)^T  for j > kk        (right-looking, in registers)
// MODE 0 (panel, NW = 4 warps = 32 rows per CTA): A = K[rows][p*128 ..], overwritten by L[rows][p] = A L_pp^-T.
//        Replaces the GEMM with the explicit 128x128 inverse: half the flops, no wait for that inverse, ~5 us.
// MODE 1 (inverse, NW = 16 warps = the 128 rows of one block, one CTA per block, run once at the end of potrf):
//        A = I, so Y = L_pp^-T = DT block; S block = Y^T; also clears the scratch.  All 64 blocks in one launch instead
//        of 64 serial in-kernel inversions on the critical path.
// ------------------------------------------------------------------------------------------------
constexpr int TS_SCRATCH = (136 + 16) * D5_BS;      // doubles exported per diagonal block

template <int NW, int MODE>
__global__ void __launch_bounds__(NW * 32)
trsm_strips_kernel(double* __restrict__ K, int Np, int p, double* __restrict__ S, double* __restrict__ DT) {
    extern __shared__ double ts_sm[];
    double* Ls = ts_sm;                              // [136][64] blocks of L_pp (A layout)
    double* Ts = ts_sm + 136 * D5_BS;                // [16][64]  T_j (A layout)
    double* scr = Ts + 16 * D5_BS + (threadIdx.x >> 5) * D5_BS;      // this warp's C-fragment -> A-operand transposer
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int lr = lane >> 2, lc = lane & 3;
    const int blk = (MODE == 0) ? p : (int)blockIdx.x;
    double* DTb = DT + (long)blk * TILE * TILE;
    // 78 KB of operand blocks: asynchronous 16-byte copies, all in flight at once (a load/store loop through registers
    // serialised on the L2 latency and took longer than the solve itself)
    for (int e = threadIdx.x; e < TS_SCRATCH / 2; e += NW * 32)
        cp_async16(ts_sm + 2 * e, DTb + 2 * e);
    cp_async_commit();
    const int ia0 = d5_ia(lr, 2 * lc);
    double a0[16], a1[16];
    long row0 = 0;
    if (MODE == 0) {
        constexpr int PER = TILE / (NW * 8);
        row0 = (long)(p + 1 + blockIdx.x / PER) * TILE + (blockIdx.x % PER) * (NW * 8) + w * 8;
        const double* Ar = K + (row0 + lr) * Np + (long)p * TILE + 2 * lc;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const double2 v = *reinterpret_cast<const double2*>(Ar + 8 * j);
            a0[j] = v.x; a1[j] = v.y;
        }
    } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            a0[j] = (j == w && lr == 2 * lc) ? 1.0 : 0.0;
            a1[j] = (j == w && lr == 2 * lc + 1) ? 1.0 : 0.0;
        }
    }
    cp_async_wait<0>();      // the strip's own rows were loaded meanwhile
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
        double r0 = 0.0, r1 = 0.0;
        double2 xf = make_double2(0.0, 0.0);
        const bool live = (MODE == 0) || (kk >= w);          // identity input: Y_kk = 0 for kk < w
        if (live) {
            __syncwarp();
            scr[ia0] = a0[kk]; scr[ia0 + 2] = a1[kk];
            __syncwarp();
            const double2 xa = *reinterpret_cast<const double2*>(scr + 2 * lane);
            const double2 t = *reinterpret_cast<const double2*>(Ts + kk * D5_BS + 2 * lane);
            dmma884(r0, r1, xa.x, t.x);                       // Y_kk = A_kk T_kk^T
            dmma884(r0, r1, xa.y, t.y);
            __syncwarp();
            scr[ia0] = r0; scr[ia0 + 2] = r1;
            __syncwarp();
            xf = *reinterpret_cast<const double2*>(scr + 2 * lane);
        }
        if (MODE == 0) {
            *reinterpret_cast<double2*>(K + (row0 + lr) * Np + (long)p * TILE + 8 * kk + 2 * lc) = make_double2(r0, r1);
        } else {        // row 8w+lr of Y: DT block directly, S block transposed (zeros where kk < w)
            const long off = (long)blk * TILE;
            const int r_ = 8 * w + lr, c_ = 8 * kk + 2 * lc;
            a0[kk] = r0; a1[kk] = r1;                         // keep: DT is also the scratch other warps still read
            S[(off + c_) * Np + off + r_] = r0;
            S[(off + c_ + 1) * Np + off + r_] = r1;
        }
        if (live) {
#pragma unroll
            for (int j = kk + 1; j < 16; ++j) {
                const double2 l = *reinterpret_cast<const double2*>(Ls + (j * (j + 1) / 2 + kk) * D5_BS + 2 * lane);
                dmma884(a0[j], a1[j], -xf.x, l.x);            // A_j -= Y_kk L(j,kk)^T
                dmma884(a0[j], a1[j], -xf.y, l.y);
            }
        }
    }
    if (MODE == 1) {        // every warp is done with the scratch (it was copied to shared memory before the barrier above)
#pragma unroll
        for (int kk = 0; kk < 16; ++kk)
            *reinterpret_cast<double2*>(DTb + (8 * w + lr) * TILE + 8 * kk + 2 * lc) = make_double2(a0[kk], a1[kk]);
    }
}
constexpr int TS_SMEM = (TS_SCRATCH + 16 * D5_BS) * 8;

// Narrow-tile editions of the two kernels on potrf's critical path (tile = ROWS x 128, ROWS = WM*S*8 in {64, 32}):
// grid.x = (#block rows) * (128 / ROWS).  tmKr is K with a ROWS-row box for the A operand.
template <int WM, int S, int NB>
__global__ void __launch_bounds__(GT_THREADS, 1)
potrf_panel_narrow_kernel(const __grid_constant__ CUtensorMap tmKr, const __grid_constant__ CUtensorMap tmS,
                          double* __restrict__ K, int Np, int p) {
    constexpr int ROWS = WM * S * 8, PER = TILE / ROWS;
    extern __shared__ unsigned char smem_raw[];
    const int row0 = (p + 1 + blockIdx.x / PER) * TILE + (blockIdx.x % PER) * ROWS;
    TOperand A = t_operand(&tmKr, row0, p * TILE);
    TOperand B = t_operand(&tmS, p * TILE, p * TILE);
    double acc[S][NB][2];
    gemm_nt_tma_mainloop_t<WM, S, NB>(acc, A, B, 0, TILE, smem_raw);
    __syncthreads();        // the A tile is this CTA's output: all reads (through smem) are done before it is overwritten
    double* Ct = K + (long)row0 * Np + (long)p * TILE;
    gemm_foreach_pair_t<WM, S, NB>(acc, [&](int row, int col, double& v0, double& v1) {
        *reinterpret_cast<double2*>(Ct + (long)row * Np + col) = make_double2(v0, v1);
    });
}

// look-ahead column: C[rows][p+1] -= L[rows][p] L[p+1][p]^T for block rows >= p+1
template <int WM, int S, int NB>
__global__ void __launch_bounds__(GT_THREADS, 1)
potrf_column_narrow_kernel(const __grid_constant__ CUtensorMap tmKr, const __grid_constant__ CUtensorMap tmK,
                           double* __restrict__ K, int Np, int p) {
    constexpr int ROWS = WM * S * 8, PER = TILE / ROWS;
    extern __shared__ unsigned char smem_raw[];
    const int row0 = (p + 1 + blockIdx.x / PER) * TILE + (blockIdx.x % PER) * ROWS;
    TOperand A = t_operand(&tmKr, row0, p * TILE);
    TOperand B = t_operand(&tmK, (p + 1) * TILE, p * TILE);
    double* Ct = K + (long)row0 * Np + (long)(p + 1) * TILE;
    double acc[S][NB][2];
    gemm_foreach_pair_t<WM, S, NB>(acc, [&](int row, int col, double& v0, double& v1) {
        const double2 c = *reinterpret_cast<const double2*>(Ct + (long)row * Np + col);
        v0 = c.x; v1 = c.y;
    });
    gemm_nt_tma_mainloop_t<WM, S, NB, false, true>(acc, A, B, 0, TILE, smem_raw);
    gemm_foreach_pair_t<WM, S, NB>(acc, [&](int row, int col, double& v0, double& v1) {
        *reinterpret_cast<double2*>(Ct + (long)row * Np + col) = make_double2(v0, v1);
    });
}

__global__ void __launch_bounds__(GT_THREADS, 1)
trtri_step1_tma_kernel(const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmS,
                       const __grid_constant__ CUtensorMap tmD, double* __restrict__ W, int Np, int nb, int h) {
    // grid (pairs, h, h): the k range shrinks with tj, so tj is the slowest index -- heaviest tiles are dispatched first
    const int q = blockIdx.x, ti = blockIdx.y, tj = blockIdx.z;
    const int left = 2 * h * q, right = left + h;
    if (right + ti >= nb) return;
    extern __shared__ unsigned char smem_raw[];
    TOperand A = t_operand(&tmK, (right + ti) * TILE, left * TILE);
    TOperand B = t_operand_alt(&tmS, (left + tj) * TILE, left * TILE, &tmD, (left + tj) * TILE, tj * TILE, (tj + 1) * TILE);
    double acc[8][4][2];
    gemm_nt_tma_mainloop(acc, A, B, tj * TILE, h * TILE, smem_raw);
    double* Wt = W + ((long)(left + tj) * TILE) * Np + (long)(right + ti) * TILE;   // W[j][i]
    gemm_foreach_pair(acc, [&](int row, int col, double v0, double v1) {
        Wt[(long)col * Np + row] = v0;
        Wt[(long)(col + 1) * Np + row] = v1;
    });
}

__global__ void __launch_bounds__(GT_THREADS, 1)
trtri_step2_tma_kernel(const __grid_constant__ CUtensorMap tmS, const __grid_constant__ CUtensorMap tmW,
                       double* __restrict__ S, int Np, int nb, int h) {
    // grid (pairs, h, h): the k range grows with ti, so the largest ti goes first
    const int q = blockIdx.x, tj = blockIdx.y, ti = h - 1 - (int)blockIdx.z;
    const int left = 2 * h * q, right = left + h;
    if (right + ti >= nb) return;
    extern __shared__ unsigned char smem_raw[];
    TOperand A = t_operand(&tmS, (right + ti) * TILE, right * TILE);
    TOperand B = t_operand(&tmW, (left + tj) * TILE, right * TILE);
    double acc[8][4][2];
    gemm_nt_tma_mainloop(acc, A, B, 0, (ti + 1) * TILE, smem_raw);
    double* Xt = S + ((long)(right + ti) * TILE) * Np + (long)(left + tj) * TILE;
    double* Xm = S + ((long)(left + tj) * TILE) * Np + (long)(right + ti) * TILE;
    gemm_foreach_pair(acc, [&](int row, int col, double v0, double v1) {
        *reinterpret_cast<double2*>(Xt + (long)row * Np + col) = make_double2(-v0, -v1);
        Xm[(long)col * Np + row] = -v0;
        Xm[(long)(col + 1) * Np + row] = -v1;
    });
}

__global__ void __launch_bounds__(GT_THREADS, 1)
lauum_tma_kernel(const __grid_constant__ CUtensorMap tmS, const __grid_constant__ CUtensorMap tmD,
                 double* __restrict__ Kinv, int Np) {
    // heaviest (longest k range = smallest ti) tiles first
    const int tj = blockIdx.x, ti = blockIdx.y;
    if (tj > ti) return;
    extern __shared__ unsigned char smem_raw[];
    TOperand A = t_operand_alt(&tmS, ti * TILE, 0, &tmD, ti * TILE, ti * TILE, (ti + 1) * TILE);
    TOperand B = (tj == ti) ? A : t_operand(&tmS, tj * TILE, 0);
    double acc[8][4][2];
    gemm_nt_tma_mainloop(acc, A, B, ti * TILE, Np, smem_raw);
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
// Narrow-tile editions of the inverse kernels (output strip = ROWS x 128, ROWS = WM*S*8 in {64, 32}).  At small and
// medium N a recursion level of trtri, or all of lauum, has fewer 128x128 tiles than the GPU has SMs and its duration
// is that of the tile with the longest k range; ROWS-row strips spread the same work over 2-4x more CTAs.
// blockIdx.y = tile row * (128/ROWS) + strip.  Operand maps with an `r` suffix have a ROWS-row box.
// ------------------------------------------------------------------------------------------------
template <int WM, int S_, int NB>
__global__ void __launch_bounds__(GT_THREADS, 1)
trtri_step1_narrow_kernel(const __grid_constant__ CUtensorMap tmKr, const __grid_constant__ CUtensorMap tmS,
                          const __grid_constant__ CUtensorMap tmD, double* __restrict__ W, int Np, int nb, int h) {
    constexpr int ROWS = WM * S_ * 8, PER = TILE / ROWS;
    const int q = blockIdx.x, ti = blockIdx.y / PER, sub = blockIdx.y % PER, tj = blockIdx.z;
    const int left = 2 * h * q, right = left + h;
    if (right + ti >= nb) return;
    extern __shared__ unsigned char smem_raw[];
    const int arow = (right + ti) * TILE + sub * ROWS;
    TOperand A = t_operand(&tmKr, arow, left * TILE);
    TOperand B = t_operand_alt(&tmS, (left + tj) * TILE, left * TILE, &tmD, (left + tj) * TILE, tj * TILE, (tj + 1) * TILE);
    double acc[S_][NB][2];
    gemm_nt_tma_mainloop_t<WM, S_, NB>(acc, A, B, tj * TILE, h * TILE, smem_raw);
    double* Wt = W + ((long)(left + tj) * TILE) * Np + arow;                 // W[j][i]
    gemm_foreach_pair_t<WM, S_, NB>(acc, [&](int row, int col, double& v0, double& v1) {
        Wt[(long)col * Np + row] = v0;
        Wt[(long)(col + 1) * Np + row] = v1;
    });
}

template <int WM, int S_, int NB>
__global__ void __launch_bounds__(GT_THREADS, 1)
trtri_step2_narrow_kernel(const __grid_constant__ CUtensorMap tmSr, const __grid_constant__ CUtensorMap tmW,
                          double* __restrict__ S, int Np, int nb, int h) {
    constexpr int ROWS = WM * S_ * 8, PER = TILE / ROWS;
    const int q = blockIdx.x, tj = blockIdx.y / PER, sub = blockIdx.y % PER, ti = h - 1 - (int)blockIdx.z;
    const int left = 2 * h * q, right = left + h;
    if (right + ti >= nb) return;
    extern __shared__ unsigned char smem_raw[];
    const int arow = (right + ti) * TILE + sub * ROWS;
    TOperand A = t_operand(&tmSr, arow, right * TILE);
    TOperand B = t_operand(&tmW, (left + tj) * TILE, right * TILE);
    double acc[S_][NB][2];
    gemm_nt_tma_mainloop_t<WM, S_, NB>(acc, A, B, 0, (ti + 1) * TILE, smem_raw);
    double* Xt = S + (long)arow * Np + (long)(left + tj) * TILE;
    double* Xm = S + ((long)(left + tj) * TILE) * Np + arow;
    gemm_foreach_pair_t<WM, S_, NB>(acc, [&](int row, int col, double& v0, double& v1) {
        *reinterpret_cast<double2*>(Xt + (long)row * Np + col) = make_double2(-v0, -v1);
        Xm[(long)col * Np + row] = -v0;
        Xm[(long)(col + 1) * Np + row] = -v1;
    });
}

template <int WM, int S_, int NB>
__global__ void __launch_bounds__(GT_THREADS, 1)
lauum_narrow_kernel(const __grid_constant__ CUtensorMap tmSr, const __grid_constant__ CUtensorMap tmDr,
                    const __grid_constant__ CUtensorMap tmS, const __grid_constant__ CUtensorMap tmD,
                    double* __restrict__ Kinv, int Np) {
    constexpr int ROWS = WM * S_ * 8, PER = TILE / ROWS;
    const int tj = blockIdx.x, ti = blockIdx.y / PER, sub = blockIdx.y % PER;     // small ti (long k range) first
    if (tj > ti) return;
    extern __shared__ unsigned char smem_raw[];
    const int arow = ti * TILE + sub * ROWS;
    TOperand A = t_operand_alt(&tmSr, arow, 0, &tmDr, arow, ti * TILE, (ti + 1) * TILE);
    TOperand B = (tj == ti) ? t_operand_alt(&tmS, tj * TILE, 0, &tmD, tj * TILE, tj * TILE, (tj + 1) * TILE)
                            : t_operand(&tmS, tj * TILE, 0);
    double acc[S_][NB][2];
    gemm_nt_tma_mainloop_t<WM, S_, NB>(acc, A, B, ti * TILE, Np, smem_raw);
    double* Ct = Kinv + (long)arow * Np + (long)tj * TILE;
    double* Cm = Kinv + ((long)tj * TILE) * Np + arow;
    const bool diag = (ti == tj);
    gemm_foreach_pair_t<WM, S_, NB>(acc, [&](int row, int col, double& v0, double& v1) {
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
// 16-byte loads (rows and segment starts are multiples of 128 doubles), four of them in flight per lane and two
// accumulators: the 8-byte edition stalled on its loads at 50 % occupancy (ncu: 4.2 TB/s, profiles/r01_ncu_hbm_stages.txt).
__device__ __forceinline__ void trmv_dot2(const double* __restrict__ a, const double* __restrict__ x, int k0, int k1,
                                          int lane, double& s0, double& s1) {
#pragma unroll 4
    for (int k = k0 + 2 * lane; k < k1; k += 64) {
        const double2 av = __ldcs(reinterpret_cast<const double2*>(a + k));
        const double2 xv = *reinterpret_cast<const double2*>(x + k);
        s0 = fma(av.x, xv.x, s0);
        s1 = fma(av.y, xv.y, s1);
    }
}

__global__ void __launch_bounds__(256) trmv_kernel(const double* __restrict__ S, const double* __restrict__ DT, int Np,
                                                   const double* __restrict__ x, double* __restrict__ y, int mode) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int i = blockIdx.x * 8 + warp;
    if (i >= Np) return;
    const int b = i / TILE;
    const double* row = S + (long)i * Np;
    double s0 = 0.0, s1 = 0.0;
    if (mode == 0) {
        trmv_dot2(row, x, 0, (b + 1) * TILE, lane, s0, s1);    // masked zeros above the diagonal inside the block
    } else {
        const double* drow = DT + (long)b * TILE * TILE + (long)(i - b * TILE) * TILE;
        trmv_dot2(drow, x + b * TILE, 0, TILE, lane, s0, s1);
        trmv_dot2(row, x, (b + 1) * TILE, Np, lane, s0, s1);
    }
    double s = s0 + s1;
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
    cudaFuncSetAttribute(potrf_panel_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GT_SMEM_BYTES);
    cudaFuncSetAttribute(potrf_trailing_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GT_SMEM_BYTES);
    cudaFuncSetAttribute(trtri_step1_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GT_SMEM_BYTES);
    cudaFuncSetAttribute(trtri_step2_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GT_SMEM_BYTES);
    cudaFuncSetAttribute(lauum_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GT_SMEM_BYTES);
    cudaFuncSetAttribute(potrf_diag3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, D3_SMEM);
    cudaFuncSetAttribute(potrf_diag5_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, D5_SMEM);
    cudaFuncSetAttribute(potrf_diag5_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, D5_SMEM);
    cudaFuncSetAttribute(trsm_strips_kernel<4, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, TS_SMEM);
    cudaFuncSetAttribute(trsm_strips_kernel<16, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, TS_SMEM);
    cudaFuncSetAttribute(potrf_panel_narrow_kernel<2, 4, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GT_SMEM_BYTES);
    cudaFuncSetAttribute(potrf_panel_narrow_kernel<1, 4, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GT_SMEM_BYTES);
    cudaFuncSetAttribute(potrf_column_narrow_kernel<2, 4, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GT_SMEM_BYTES);
    cudaFuncSetAttribute(potrf_column_narrow_kernel<1, 4, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GT_SMEM_BYTES);
    cudaFuncSetAttribute(potrf_column_narrow_kernel<1, 2, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GT_SMEM_BYTES);
    cudaFuncSetAttribute(trtri_step1_narrow_kernel<2, 4, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GT_SMEM_BYTES);
    cudaFuncSetAttribute(trtri_step1_narrow_kernel<1, 4, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GT_SMEM_BYTES);
    cudaFuncSetAttribute(trtri_step2_narrow_kernel<2, 4, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GT_SMEM_BYTES);
    cudaFuncSetAttribute(trtri_step2_narrow_kernel<1, 4, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GT_SMEM_BYTES);
    cudaFuncSetAttribute(lauum_narrow_kernel<2, 4, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GT_SMEM_BYTES);
    cudaFuncSetAttribute(lauum_narrow_kernel<1, 4, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GT_SMEM_BYTES);
    if (dev >= 0 && dev < 64) g_attr_set[dev] = true;
}

static int diag_impl() {       // GPX_DIAG_IMPL = 3 | 5 (default 5)
    static int v = -1;
    if (v < 0) { const char* e = getenv("GPX_DIAG_IMPL"); v = (e && e[0] == '3') ? 3 : 5; }
    return v;
}

// Helper stream + events for the look-ahead (per device, created on first use; all work is ordered
// after the caller's stream and joined back into it before potrf returns control of the data).
struct LookAhead {
    std::mutex enqueue;                    // the helper streams/events are shared per device: one potrf enqueues at a time
    cudaStream_t side = nullptr;
    cudaStream_t chain = nullptr;          // highest-priority stream for the critical path (GPX_POTRF_PRIO=1)
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    cudaEvent_t ev_main[2] = {nullptr, nullptr};
    cudaEvent_t ev_side[2] = {nullptr, nullptr};
};
static LookAhead g_la[64];

static LookAhead* lookahead_for_current_device() {
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) return nullptr;
    LookAhead* la = &g_la[dev];
    if (la->side == nullptr) {
        // lowest priority: when SMs free up, pending CTAs of the caller's stream (the critical path) are placed first
        int prio_lo = 0, prio_hi = 0;
        cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
        if (cudaStreamCreateWithPriority(&la->side, cudaStreamNonBlocking, prio_lo) != cudaSuccess) return nullptr;
        if (cudaStreamCreateWithPriority(&la->chain, cudaStreamNonBlocking, prio_hi) != cudaSuccess) return nullptr;
        cudaEventCreateWithFlags(&la->ev_fork, cudaEventDisableTiming);
        cudaEventCreateWithFlags(&la->ev_join, cudaEventDisableTiming);
        for (int i = 0; i < 2; ++i) {
            cudaEventCreateWithFlags(&la->ev_main[i], cudaEventDisableTiming);
            cudaEventCreateWithFlags(&la->ev_side[i], cudaEventDisableTiming);
        }
    }
    return la;
}

static int use_chol_tma() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("GPX_CHOL_IMPL"); v = (e && e[0] == 'c') ? 0 : 1; }
    return v;
}

static int use_lookahead() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("GPX_POTRF_LOOKAHEAD"); v = (e && e[0] == '0') ? 0 : 1; }
    return v;
}

// K (lower tiles valid) -> L in place; diagonal-block inverses into S / DT; per-block log-dets; info.
// Right-looking with a one-panel look-ahead: after panel p, the next block column is updated on the caller's
// stream (critical path: diag p+1, panel p+1), while the rest of the trailing matrix is updated on a helper
// stream.
// gpx_i8.cu: the bulk trailing update of one outer block on the integer tensor cores
int potrf_i8_slice(const double* K, int Np, int c0, int c1, int r0, void* workspace, size_t workspace_bytes, int buf, cudaStream_t st);
int potrf_i8_update(double* K, int Np, int c0, int c1, int r0, int ncol_blocks, void* workspace, int buf, cudaStream_t st);

// i8_ws != null (training mode, gpx_potrf_i8): the trailing updates of the two-level schedule -- once per outer block, the
// look-ahead columns on the critical-path stream and everything right of them on the helper stream -- run as sliced-integer
// products (six 8-bit digits per operand, exact int32 sums; the finished panels are sliced once per outer block, into
// alternating plane buffers) instead of FP64 DMMA; the panels, the in-block updates and the chain-bound tail stay FP64.  Outer
// blocks are then OB = 4 panels wide by default (K = 512 per integer product; 3, 6 and 8 measured within 4 % of it at Np = 8192),
// 6 from 96 block columns and 8 from 192 (Np = 26240: 107 -> 91 ms).
int potrf(double* K, int Np, int N, double* S, double* DT, double* logdet_blk, int* info, cudaStream_t st_caller, void* i8_ws,
          size_t i8_bytes) {
    set_gemm_attrs();
    cudaStream_t st = st_caller;
    const int nb = Np / TILE;
    LookAhead* la = use_lookahead() ? lookahead_for_current_device() : nullptr;
    // Every event record/wait pair below is resolved at enqueue time, so holding the lock while enqueuing is enough to
    // keep concurrent callers (other host threads, other streams) from interleaving their fork/join sequences.
    std::unique_lock<std::mutex> enqueue_lock;
    if (la != nullptr) enqueue_lock = std::unique_lock<std::mutex>(la->enqueue);
    cudaMemsetAsync(info, 0, sizeof(int), st);
    static int prio = -1;
    if (prio < 0) { const char* e = getenv("GPX_POTRF_PRIO"); prio = (e && e[0] == '0') ? 0 : 1; }
    if (la != nullptr && prio) {           // run the critical path on the high-priority stream, forked from the caller's
        cudaEventRecord(la->ev_fork, st_caller);
        cudaStreamWaitEvent(la->chain, la->ev_fork, 0);
        st = la->chain;
    }
    const bool tma = use_chol_tma();
    CUtensorMap tmK, tmS, tmK64, tmK32, tmK16;
    if (tma) {
        int rc;
        if ((rc = encode_f64_rowmajor(&tmK, K, Np, Np, Np, TILE))) return rc;
        if ((rc = encode_f64_rowmajor(&tmS, S, Np, Np, Np, TILE))) return rc;
        if ((rc = encode_f64_rowmajor(&tmK64, K, Np, Np, Np, 64))) return rc;
        if ((rc = encode_f64_rowmajor(&tmK32, K, Np, Np, Np, 32))) return rc;
        if ((rc = encode_f64_rowmajor(&tmK16, K, Np, Np, Np, 16))) return rc;
    }
    // panel / look-ahead column over n block rows: with few rows, narrower tiles put the same work on more SMs
    static int narrow = -1;
    if (narrow < 0) { const char* e = getenv("GPX_POTRF_NARROW"); narrow = (e && e[0] == '0') ? 0 : 1; }
    // GPX_POTRF_PANEL=gemm: panel = GEMM with the explicit inverse of the diagonal block (computed inside the diagonal
    // kernel); default: Cholesky-only diagonal kernel + blocked TRSM panel + one batched inversion launch at the end
    static int panel_trsm = -1;
    if (panel_trsm < 0) { const char* e = getenv("GPX_POTRF_PANEL"); panel_trsm = (e && e[0] == 'g') ? 0 : 1; }
    const bool trsm_mode = tma && panel_trsm && diag_impl() == 5;
    auto panel = [&](int n, int p) {
        if (trsm_mode) trsm_strips_kernel<4, 0><<<4 * n, 128, TS_SMEM, st>>>(K, Np, p, S, DT);
        else if (!tma) potrf_panel_kernel<<<n, G_THREADS, G_SMEM_BYTES, st>>>(K, Np, p, S);
        else if (narrow && 4 * n <= 148) potrf_panel_narrow_kernel<1, 4, 2><<<4 * n, GT_THREADS, GT_SMEM_BYTES, st>>>(tmK32, tmS, K, Np, p);
        else if (narrow && 2 * n <= 148) potrf_panel_narrow_kernel<2, 4, 4><<<2 * n, GT_THREADS, GT_SMEM_BYTES, st>>>(tmK64, tmS, K, Np, p);
        else potrf_panel_tma_kernel<<<n, GT_THREADS, GT_SMEM_BYTES, st>>>(tmK, tmS, K, Np, p);
    };
    auto column = [&](int n, int p) {          // block rows p+1 .. p+n of block column p+1
        if (tma && narrow && 8 * n <= 148) potrf_column_narrow_kernel<1, 2, 2><<<8 * n, GT_THREADS, GT_SMEM_BYTES, st>>>(tmK16, tmK, K, Np, p);
        else if (tma && narrow && 4 * n <= 148) potrf_column_narrow_kernel<1, 4, 2><<<4 * n, GT_THREADS, GT_SMEM_BYTES, st>>>(tmK32, tmK, K, Np, p);
        else if (tma && narrow && 2 * n <= 148) potrf_column_narrow_kernel<2, 4, 4><<<2 * n, GT_THREADS, GT_SMEM_BYTES, st>>>(tmK64, tmK, K, Np, p);
        else if (tma) potrf_trailing_tma_kernel<<<dim3(1, n), GT_THREADS, GT_SMEM_BYTES, st>>>(tmK, K, Np, p, p + 1, p + 1, 1);
        else potrf_trailing_kernel<<<dim3(1, n), G_THREADS, G_SMEM_BYTES, st>>>(K, Np, p, p + 1, p + 1);
    };
    auto trailing = [&](dim3 grid, int p, int j0, int i0, cudaStream_t s_) {
        if (tma) potrf_trailing_tma_kernel<<<grid, GT_THREADS, GT_SMEM_BYTES, s_>>>(tmK, K, Np, p, j0, i0, 1);
        else potrf_trailing_kernel<<<grid, G_THREADS, G_SMEM_BYTES, s_>>>(K, Np, p, j0, i0);
    };
    auto diag = [&](int p) {
        if (diag_impl() == 3)
            potrf_diag3_kernel<<<1, D3_THREADS, D3_SMEM, st>>>(K, Np, p, N, S, DT, logdet_blk, info);
        else if (trsm_mode)
            potrf_diag5_kernel<false><<<1, D5_THREADS, D5_SMEM, st>>>(K, Np, p, N, S, DT, logdet_blk, info);
        else
            potrf_diag5_kernel<true><<<1, D5_THREADS, D5_SMEM, st>>>(K, Np, p, N, S, DT, logdet_blk, info);
    };
    static int ob_env = -1, sw_env = -1;
    if (ob_env < 0) { const char* e = getenv("GPX_POTRF_OB"); ob_env = (e && e[0] >= '1' && e[0] <= '8') ? e[0] - '0' : 0; }
    if (sw_env < 0) { const char* e = getenv("GPX_POTRF_SWITCH"); sw_env = e ? atoi(e) : 0; }
    static int ob8_env = -1;
    if (ob8_env < 0) { const char* e = getenv("GPX_POTRF_OB_I8"); ob8_env = (e && e[0] >= '1' && e[0] <= '8') ? e[0] - '0' : 0; }
    // block columns left to the one-panel loop: 32 with FP64 updates; 24 when the outer blocks' updates are integer products
    // (cheaper, so the two-level schedule pays a little longer: 4.71 -> 4.61 ms at Np = 8192)
    const int sw = sw_env > 0 ? sw_env : (i8_ws != nullptr ? 24 : 32);
    const int OB = i8_ws != nullptr ? (ob8_env ? ob8_env : (nb >= 192 ? 8 : nb >= 96 ? 6 : 4)) : (ob_env ? ob_env : (nb >= 96 ? 4 : 3));
    int p_start = 0;
    if (tma && la != nullptr && st != st_caller && OB >= 2 && nb > sw) {
        // Two-level schedule.  Panels are still 128 wide (diag + panel on the critical path), but the bulk of the
        // trailing matrix is updated once per OUTER block of OB panels with K = OB*128, which halves (OB = 2) the
        // prologue/epilogue share of every output tile.  Inside an outer block the panels only update the block's own
        // columns; afterwards the next outer block's columns are updated on the critical-path stream (look-ahead) and
        // everything to the right of them on the low-priority helper stream.  Only while the trailing matrix is large
        // enough to be throughput-bound: the last `sw` block columns are chain-bound and go through the one-panel
        // look-ahead loop below, whose critical path per panel is shorter.
        bool pending = false;
        int q_idx = 0;
        for (int q = 0; q < nb && nb - q > sw; q += OB, ++q_idx) {
            const int qe = (q + OB < nb) ? q + OB : nb;
            for (int p = q; p < qe; ++p) {
                diag(p);
                const int n = nb - p - 1;
                if (n <= 0) break;
                panel(n, p);
                if (p + 1 < qe)      // the outer block's remaining columns, K = 128
                    potrf_trailing_tma_kernel<<<dim3(qe - p - 1, n), GT_THREADS, GT_SMEM_BYTES, st>>>(tmK, K, Np, p, p + 1, p + 1, 1);
            }
            if (qe >= nb) { p_start = nb; break; }
            const int kp = qe - q;
            const int la_cols = (qe + OB < nb) ? OB : nb - qe;          // look-ahead: the next outer block's columns
            const int rest = nb - qe - la_cols;                          // helper stream: everything right of them
            if (i8_ws != nullptr) {  // digits of the outer block's panels, every row below them (buffer q_idx & 1: the helper's
                                     // update q_idx - 2 that read it last was waited for during outer block q_idx - 1)
                const int rc8 = potrf_i8_slice(K, Np, q * TILE, qe * TILE, qe * TILE, i8_ws, i8_bytes, q_idx & 1, st);
                if (rc8) return rc8;
            }
            if (rest > 0) {          // the helper only needs this outer block's panels
                cudaEventRecord(la->ev_main[q_idx & 1], st);
                cudaStreamWaitEvent(la->side, la->ev_main[q_idx & 1], 0);
            }
            // the look-ahead columns were last written by the helper's previous outer update
            if (pending) cudaStreamWaitEvent(st, la->ev_side[(q_idx - 1) & 1], 0);
            if (rest > 0) {
                if (i8_ws != nullptr) {
                    const int rc8 = potrf_i8_update(K, Np, q * TILE, qe * TILE, (qe + la_cols) * TILE, -1, i8_ws, q_idx & 1, la->side);
                    if (rc8) return rc8;
                } else {
                    potrf_trailing_tma_kernel<<<dim3(rest, rest), GT_THREADS, GT_SMEM_BYTES, la->side>>>(tmK, K, Np, q, qe + la_cols,
                                                                                                     qe + la_cols, kp);
                }
                cudaEventRecord(la->ev_side[q_idx & 1], la->side);
                pending = true;
            } else {
                pending = false;
            }
            if (i8_ws != nullptr) {
                const int rc8 = potrf_i8_update(K, Np, q * TILE, qe * TILE, qe * TILE, la_cols, i8_ws, q_idx & 1, st);
                if (rc8) return rc8;
            } else {
                potrf_trailing_tma_kernel<<<dim3(la_cols, nb - qe), GT_THREADS, GT_SMEM_BYTES, st>>>(tmK, K, Np, q, qe, qe, kp);
            }
            p_start = qe;
        }
        // hand over to the one-panel loop with the whole trailing matrix up to date
        if (pending) cudaStreamWaitEvent(st, la->ev_side[(q_idx - 1) & 1], 0);
    }
    bool side_pending = false;
    static int ev_old = -1;
    if (ev_old < 0) { const char* e = getenv("GPX_POTRF_EV"); ev_old = e ? (e[0] == 'o') : !prio; }    // releasing the helper stream before the look-ahead column only pays with the priority stream
    for (int p = p_start; p < nb; ++p) {
        diag(p);
        const int n = nb - p - 1;
        if (n <= 0) break;
        panel(n, p);
        if (la == nullptr) {
            trailing(dim3(n, n), p, p + 1, p + 1, st);
            continue;
        }
        // the helper stream's update p only needs panel p (it touches block columns >= p+2): release it before the
        // look-ahead column so that it starts right behind update p-1 instead of idling through the column update
        const bool early = !ev_old && n > 1;
        if (early) {
            cudaEventRecord(la->ev_main[p & 1], st);
            cudaStreamWaitEvent(la->side, la->ev_main[p & 1], 0);
        }
        // block column p+1 was also touched by the helper stream's update p-1: wait for it first
        if (side_pending) cudaStreamWaitEvent(st, la->ev_side[(p - 1) & 1], 0);
        if (early) {
            trailing(dim3(n - 1, n - 1), p, p + 2, p + 2, la->side);
            cudaEventRecord(la->ev_side[p & 1], la->side);
        }
        column(n, p);
        if (n > 1) {
            if (!early) {
                cudaEventRecord(la->ev_main[p & 1], st);
                cudaStreamWaitEvent(la->side, la->ev_main[p & 1], 0);
                trailing(dim3(n - 1, n - 1), p, p + 2, p + 2, la->side);
                cudaEventRecord(la->ev_side[p & 1], la->side);
            }
            side_pending = true;
        } else {
            side_pending = false;
        }
    }
    if (trsm_mode)      // S / DT diagonal blocks for all block columns at once, from the exported 8x8 blocks
        trsm_strips_kernel<16, 1><<<nb, 512, TS_SMEM, st>>>(K, Np, 0, S, DT);
    if (la != nullptr && side_pending) cudaStreamWaitEvent(st, la->ev_side[(nb - 3) & 1], 0);   // last helper launch: p = nb-3
    if (st != st_caller) {
        cudaEventRecord(la->ev_join, st);
        cudaStreamWaitEvent(st_caller, la->ev_join, 0);
    }
    GPX_CHECK_LAUNCH();
    return 0;
}

// Completes S (diagonal blocks already hold D_b) to the full mirrored inverse. W = Np x Np scratch.
static int inverse_narrow_tiles() {       // GPX_INV_NARROW=0: 128-row tiles only
    static int v = -1;
    if (v < 0) { const char* e = getenv("GPX_INV_NARROW"); v = (e && e[0] == '0') ? 0 : 1; }
    return v;
}

// gpx_i8.cu: one merge level on the integer tensor cores
int trtri_i8_merge(const double* L, double* Smat, const double* DT, double* W, int Np, int left, int h, int hr, void* workspace,
                   int bits, cudaStream_t st);
size_t trtri_i8_workspace_bytes(int Np, int nslices);
constexpr int TRTRI_I8_MIN_H = 8;        // merge levels whose segments have >= 1024 rows run as sliced-integer GEMMs

static int trtri_impl(const double* L, double* S, const double* DT, double* W, int Np, void* i8_ws, int i8_bits, cudaStream_t st);

int trtri(const double* L, double* S, const double* DT, double* W, int Np, cudaStream_t st) {
    return trtri_impl(L, S, DT, W, Np, nullptr, 0, st);
}

// As trtri; the top merge levels (segments of >= 1024 rows) run on the integer tensor cores from six 7-bit slices per operand
// (entries of L^-1 then agree with the all-FP64 result to ~1e-11 of their row/column scales).
int trtri_i8(const double* L, double* S, const double* DT, double* W, int Np, int bits, void* workspace, size_t workspace_bytes,
             cudaStream_t st) {
    if (bits != 7 && bits != 8) return -6;
    if (Np > (bits == 8 ? 16384 : 32768)) return -4;
    if (workspace_bytes < trtri_i8_workspace_bytes(Np, 6)) return -7;
    return trtri_impl(L, S, DT, W, Np, workspace, bits, st);
}

static int trtri_impl(const double* L, double* S, const double* DT, double* W, int Np, void* i8_ws, int i8_bits, cudaStream_t st) {
    set_gemm_attrs();
    const int nb = Np / TILE;
    const bool tma = use_chol_tma();
    CUtensorMap tmK, tmS, tmW, tmD, tmK64, tmK32, tmS64, tmS32;
    if (tma) {
        int rc;
        if ((rc = encode_f64_rowmajor(&tmK, L, Np, Np, Np, TILE))) return rc;
        if ((rc = encode_f64_rowmajor(&tmS, S, Np, Np, Np, TILE))) return rc;
        if ((rc = encode_f64_rowmajor(&tmW, W, Np, Np, Np, TILE))) return rc;
        if ((rc = encode_f64_rowmajor(&tmD, DT, (uint64_t)nb * TILE, TILE, TILE, TILE))) return rc;
        if ((rc = encode_f64_rowmajor(&tmK64, L, Np, Np, Np, 64))) return rc;
        if ((rc = encode_f64_rowmajor(&tmK32, L, Np, Np, Np, 32))) return rc;
        if ((rc = encode_f64_rowmajor(&tmS64, S, Np, Np, Np, 64))) return rc;
        if ((rc = encode_f64_rowmajor(&tmS32, S, Np, Np, Np, 32))) return rc;
    }
    const int narrow = inverse_narrow_tiles();
    for (int h = 1; h < nb; h *= 2) {
        const int pairs = (nb + 2 * h - 1) / (2 * h);
        dim3 grid(h, h, pairs);
        const int tiles = pairs * h * h;
        if (i8_ws != nullptr && h >= TRTRI_I8_MIN_H) {
            for (int q = 0; q < pairs; ++q) {
                const int left = 2 * h * q, right = left + h;
                if (right >= nb) continue;
                const int hr = (nb - right) < h ? (nb - right) : h;
                const int rc = trtri_i8_merge(L, S, DT, W, Np, left, h, hr, i8_ws, i8_bits, st);
                if (rc) return rc;
            }
        } else if (tma && narrow && tiles <= 300) {             // 32-row strips
            dim3 g(pairs, 4 * h, h);
            trtri_step1_narrow_kernel<1, 4, 2><<<g, GT_THREADS, GT_SMEM_BYTES, st>>>(tmK32, tmS, tmD, W, Np, nb, h);
            trtri_step2_narrow_kernel<1, 4, 2><<<g, GT_THREADS, GT_SMEM_BYTES, st>>>(tmS32, tmW, S, Np, nb, h);
        } else if (tma && narrow && tiles <= 600) {      // 64-row strips
            dim3 g(pairs, 2 * h, h);
            trtri_step1_narrow_kernel<2, 4, 4><<<g, GT_THREADS, GT_SMEM_BYTES, st>>>(tmK64, tmS, tmD, W, Np, nb, h);
            trtri_step2_narrow_kernel<2, 4, 4><<<g, GT_THREADS, GT_SMEM_BYTES, st>>>(tmS64, tmW, S, Np, nb, h);
        } else if (tma) {
            dim3 grid_hf(pairs, h, h);      // heaviest-first dispatch order
            trtri_step1_tma_kernel<<<grid_hf, GT_THREADS, GT_SMEM_BYTES, st>>>(tmK, tmS, tmD, W, Np, nb, h);
            trtri_step2_tma_kernel<<<grid_hf, GT_THREADS, GT_SMEM_BYTES, st>>>(tmS, tmW, S, Np, nb, h);
        } else {
            trtri_step1_kernel<<<grid, G_THREADS, G_SMEM_BYTES, st>>>(L, S, DT, W, Np, nb, h);
            trtri_step2_kernel<<<grid, G_THREADS, G_SMEM_BYTES, st>>>(S, W, Np, nb, h);
        }
    }
    GPX_CHECK_LAUNCH();
    return 0;
}

int lauum(const double* S, const double* DT, double* Kinv, int Np, cudaStream_t st) {
    set_gemm_attrs();
    const int nb = Np / TILE;
    if (use_chol_tma()) {
        CUtensorMap tmS, tmD;
        int rc;
        if ((rc = encode_f64_rowmajor(&tmS, S, Np, Np, Np, TILE))) return rc;
        if ((rc = encode_f64_rowmajor(&tmD, DT, (uint64_t)nb * TILE, TILE, TILE, TILE))) return rc;
        const int tiles = nb * (nb + 1) / 2;
        if (inverse_narrow_tiles() && tiles <= 1200) {
            const int rows = tiles <= 300 ? 32 : 64;
            CUtensorMap tmSr, tmDr;
            if ((rc = encode_f64_rowmajor(&tmSr, S, Np, Np, Np, rows))) return rc;
            if ((rc = encode_f64_rowmajor(&tmDr, DT, (uint64_t)nb * TILE, TILE, TILE, rows))) return rc;
            if (rows == 32)
                lauum_narrow_kernel<1, 4, 2><<<dim3(nb, 4 * nb), GT_THREADS, GT_SMEM_BYTES, st>>>(tmSr, tmDr, tmS, tmD, Kinv, Np);
            else
                lauum_narrow_kernel<2, 4, 4><<<dim3(nb, 2 * nb), GT_THREADS, GT_SMEM_BYTES, st>>>(tmSr, tmDr, tmS, tmD, Kinv, Np);
        } else {
            lauum_tma_kernel<<<dim3(nb, nb), GT_THREADS, GT_SMEM_BYTES, st>>>(tmS, tmD, Kinv, Np);
        }
    } else {
        lauum_kernel<<<dim3(nb, nb), G_THREADS, G_SMEM_BYTES, st>>>(S, DT, Kinv, Np);
    }
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
