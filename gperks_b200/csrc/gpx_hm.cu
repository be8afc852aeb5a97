// History-matching epilogue: per-point implausibility over several emulators.
// Replaces the per-point Python loop with two np.sort calls in Wave.compute_impl
// (GPErks/perks/history_matching.py:57-62):
//     I_i  = maxno-th largest_j  sqrt( (M_ij - z_j)^2 / (V_ij + v_j) ),   V = std^2
//     PV_i = maxno-th largest_j  V_ij / v_j
// Inputs are emulator-major (mean[e][m], std[e][m]) exactly as the per-emulator predict calls leave them in
// HBM, so nothing is transposed or copied to the host.  One thread per point; HBM-bound (16 E B read + 16 B
// written per point).  IEEE double sqrt/div/mul in the reference's operation order => bit-identical to numpy.
#include "gpx_common.cuh"

namespace gpx {

constexpr int HM_KMAX = 8;
constexpr int HM_UNROLL = 4;       // emulators whose (mean, std) loads are issued before any of them is consumed

// Insert (a, b) into the descending lists of the MAXNO largest values seen so far.
template <int MAXNO>
__device__ __forceinline__ void hm_insert(double (&topI)[MAXNO], double (&topP)[MAXNO], double a, double b) {
#pragma unroll
    for (int k = 0; k < MAXNO; ++k) {
        const double ti = topI[k], tp = topP[k];
        const bool ci = a > ti, cp = b > tp;
        topI[k] = ci ? a : ti;  a = ci ? ti : a;
        topP[k] = cp ? b : tp;  b = cp ? tp : b;
    }
}

// Explicit round-to-nearest intrinsics: nvcc must not contract sd*sd + v into an FMA, numpy does not.
__device__ __forceinline__ void hm_terms(double mu, double sd, double ze, double ve, double& a, double& b) {
    const double var = __dmul_rn(sd, sd);
    const double d = __dsub_rn(mu, ze);
    a = __dsqrt_rn(__ddiv_rn(__dmul_rn(d, d), __dadd_rn(var, ve)));
    b = __ddiv_rn(var, ve);
}

// One thread per point; the maxno-th largest is kept in MAXNO registers per list (compile-time MAXNO: the edition with a
// run-time bound on 8-entry lists needed 151 registers, ran one block per SM with one load in flight per thread and
// reached 0.68 TB/s, profiles/r01_ncu_hbm_stages.txt).
template <int MAXNO>
__global__ void __launch_bounds__(256, (MAXNO <= 4 ? 4 : 3))
implausibility_kernel(const double* __restrict__ mean, const double* __restrict__ std_, long M, int E,
                      const double* __restrict__ z, const double* __restrict__ v, double* __restrict__ I, double* __restrict__ PV) {
    const long m = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= M) return;
    double topI[MAXNO], topP[MAXNO];          // descending: top[0] largest ... top[MAXNO-1] = answer
#pragma unroll
    for (int k = 0; k < MAXNO; ++k) { topI[k] = -INFINITY; topP[k] = -INFINITY; }
    const double* pm = mean + m;
    const double* ps = std_ + m;
    int e = 0;
    for (; e + HM_UNROLL <= E; e += HM_UNROLL) {
        double mu[HM_UNROLL], sd[HM_UNROLL];
#pragma unroll
        for (int u = 0; u < HM_UNROLL; ++u) {
            mu[u] = __ldcs(pm + (long)(e + u) * M);       // streamed once: evict-first
            sd[u] = __ldcs(ps + (long)(e + u) * M);
        }
#pragma unroll
        for (int u = 0; u < HM_UNROLL; ++u) {
            double a, b;
            hm_terms(mu[u], sd[u], __ldg(z + e + u), __ldg(v + e + u), a, b);
            hm_insert<MAXNO>(topI, topP, a, b);
        }
    }
    for (; e < E; ++e) {
        double a, b;
        hm_terms(__ldcs(pm + (long)e * M), __ldcs(ps + (long)e * M), __ldg(z + e), __ldg(v + e), a, b);
        hm_insert<MAXNO>(topI, topP, a, b);
    }
    __stcs(I + m, topI[MAXNO - 1]);
    __stcs(PV + m, topP[MAXNO - 1]);
}

int implausibility(const double* mean, const double* std_, long M, int E, const double* z, const double* v, int maxno,
                   double* I, double* PV, cudaStream_t st) {
    if (maxno < 1 || maxno > HM_KMAX || maxno > E) return -7;
    if (M == 0) return 0;
    const unsigned grid = (unsigned)((M + 255) / 256);
#define GPX_HM_CASE(K) case K: implausibility_kernel<K><<<grid, 256, 0, st>>>(mean, std_, M, E, z, v, I, PV); break;
    switch (maxno) {
        GPX_HM_CASE(1) GPX_HM_CASE(2) GPX_HM_CASE(3) GPX_HM_CASE(4)
        GPX_HM_CASE(5) GPX_HM_CASE(6) GPX_HM_CASE(7) GPX_HM_CASE(8)
    }
#undef GPX_HM_CASE
    GPX_CHECK_LAUNCH();
    return 0;
}

}  // namespace gpx
