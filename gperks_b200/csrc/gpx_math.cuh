// gperks_b200 -- branch-free FP64 exp(-y) and sqrt for the pairwise-kernel builders (gpx_kmat.cu, mll_grad).
//
// Why not libdevice's exp()/sqrt(): every call carries a slow-path guard (BSSY/BSYNC + CALL) that fences the instruction
// scheduler, so the eight independent entries of a micro-tile row were evaluated one after another, and every polynomial
// coefficient was rebuilt from two UMOV immediates per use.  ncu of round 2's K* builder: 168 issued instructions per entry,
// 33 of them FP64, FP64 pipe 46 % busy.  Here the argument ranges are known (r^2 >= 1e-30, exp of a non-positive number
// clamped at -700), so there is no guard at all, and the coefficients live in the constant bank where DFMA reads them as
// an operand (c[3][..]) for free.
//
//   sqrt_pos(x), 1e-300 < x < 1e300:  MUFU.RSQ64H seed (>= 20 bits), one third-order step  y1 = y0 (1 + e/2 + 3 e^2/8),
//       e = 1 - x y0^2, then  g = x y1,  sqrt = g + (x - g^2) y1/2  -- the textbook coupled iteration; the final FMA
//       rounds the exactly-computed residual into g, which is what makes the result correctly rounded.
//   exp_neg(y) = exp(-y), y >= 0 (clamped at 700):  n = rint(-y log2 e) by the 1.5*2^52 shift, r = -y - n ln2 in two FMAs
//       (ln2_hi has 27 trailing zero bits: n ln2_hi is exact), exp(r) = 1 + r + r^2 q(r) with q the degree-9 interpolant of
//       (e^r - 1 - r)/r^2 at the Chebyshev nodes of |r| <= 1.001 ln2/2 (max relative error of the polynomial 1.6e-17,
//       coefficients from tools/fit_exp_poly.py; evaluated as two interleaved Horner chains in r^2), 2^n added into the exponent
//       field (result >= 2^-1011: always normal).
//
// The functions are __host__ __device__ so that tests/test_math_host.py can sweep them against libm on the CPU (the host
// build seeds the square root with 1/sqrtf and reads the coefficients from a plain array).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#if !defined(__CUDA_ARCH__)
#include <cmath>
#include <cstring>
#endif

namespace gpx {

#define GPX_EXP_COEFFS                                                                                              \
    {0.5000000000000001, 0.16666666666666669, 0.041666666666623824, 0.008333333333330039, 0.0013888888917367081,    \
     0.00019841269863171557, 2.4801521057868204e-05, 2.7557268276905696e-06, 2.7620201591031547e-07,                \
     2.5100472505694996e-08, /* [10..13]: log2(e), ln2_hi, ln2_lo, 1.5 * 2^52 */                                    \
     1.4426950408889634, 0.6931471675634384, 1.2996506893889889e-08, 6755399441055744.0}

#if defined(__CUDACC__)
static __constant__ double c_gpx_math[14] = GPX_EXP_COEFFS;
#endif
static const double h_gpx_math[14] = GPX_EXP_COEFFS;

#if defined(__CUDA_ARCH__)
#define GPX_MC(i) c_gpx_math[i]
#else
#define GPX_MC(i) h_gpx_math[i]
#endif

__host__ __device__ __forceinline__ double hi_add(double x, int delta_hi) {   // adds delta_hi to the upper 32 bits of x
#if defined(__CUDA_ARCH__)
    return __hiloint2double(__double2hiint(x) + delta_hi, __double2loint(x));
#else
    uint64_t u;
    std::memcpy(&u, &x, 8);
    u += (uint64_t)(int64_t)delta_hi << 32;
    std::memcpy(&x, &u, 8);
    return x;
#endif
}

__host__ __device__ __forceinline__ double sqrt_pos(double x) {
    double y;
#if defined(__CUDA_ARCH__)
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
#else
    y = (double)(1.0f / std::sqrt((float)x));
#endif
    const double t = y * y;
    const double e = fma(x, -t, 1.0);
    const double h = fma(e, 0.375, 0.5);
    const double ye = y * e;
    const double y1 = fma(h, ye, y);
    const double g = x * y1;
    const double res = fma(-g, g, x);
    return fma(res, hi_add(y1, -0x100000), g);      // y1 / 2 on the integer pipe
}

// min(y, ~700) for y >= 0 on the integer pipe: the upper words of non-negative doubles order like the doubles themselves, so
// clamping the upper word at that of 700.0 (0x4085E000) leaves y untouched below 700 and maps anything above to
// 700 + (low word) * 2^-43 <= 700.0005 (one IMNMX instead of DSETP + 2 SEL).  NaN / negative inputs are the caller's business.
__host__ __device__ __forceinline__ double clamp700(double y) {
#if defined(__CUDA_ARCH__)
    return __hiloint2double(min(__double2hiint(y), 0x4085E000), __double2loint(y));
#else
    uint64_t u;
    std::memcpy(&u, &y, 8);
    uint32_t hi = (uint32_t)(u >> 32);
    if ((int32_t)hi > 0x4085E000) hi = 0x4085E000u;
    u = ((uint64_t)hi << 32) | (u & 0xffffffffull);
    std::memcpy(&y, &u, 8);
    return y;
#endif
}

__host__ __device__ __forceinline__ double exp_neg(double y) {
    y = clamp700(y);
    const double magic = GPX_MC(13);
    const double t = fma(-y, GPX_MC(10), magic);
    const double nd = t - magic;
    double r = fma(nd, -GPX_MC(11), -y);
    r = fma(nd, -GPX_MC(12), r);
    // exp(r) = 1 + (r + r^2 q(r)),  q = qe(r^2) + r qo(r^2): two independent Horner chains of four FMAs (depth 8 instead of 11, so a
    // thread that runs out of registers for interleaving several entries still keeps the FP64 pipe fed)
    const double r2 = r * r;
    double qe = GPX_MC(8), qo = GPX_MC(9);
#pragma unroll
    for (int i = 3; i >= 0; --i) {
        qe = fma(qe, r2, GPX_MC(2 * i));
        qo = fma(qo, r2, GPX_MC(2 * i + 1));
    }
    double p = fma(fma(qo, r, qe), r2, r) + 1.0;
#if defined(__CUDA_ARCH__)
    const int n = __double2loint(t);
#else
    int n;
    {
        uint64_t u;
        std::memcpy(&u, &t, 8);
        n = (int)(uint32_t)u;
    }
#endif
    return hi_add(p, (int)((unsigned)n << 20));
}

}  // namespace gpx
