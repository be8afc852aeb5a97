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

__global__ void implausibility_kernel(const double* __restrict__ mean, const double* __restrict__ std_, long M, int E,
                                      const double* __restrict__ z, const double* __restrict__ v, int maxno,
                                      double* __restrict__ I, double* __restrict__ PV) {
    const long m = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= M) return;
    double topI[HM_KMAX], topP[HM_KMAX];      // descending: top[0] largest ... top[maxno-1] = answer
#pragma unroll
    for (int k = 0; k < HM_KMAX; ++k) { topI[k] = -INFINITY; topP[k] = -INFINITY; }
    for (int e = 0; e < E; ++e) {
        const double mu = mean[(long)e * M + m];
        const double sd = std_[(long)e * M + m];
        const double var = sd * sd;
        const double d = mu - z[e];
        double a = sqrt((d * d) / (var + v[e]));
        double b = var / v[e];
#pragma unroll
        for (int k = 0; k < HM_KMAX; ++k) {
            if (k < maxno) {
                if (a > topI[k]) { double t = topI[k]; topI[k] = a; a = t; }
                if (b > topP[k]) { double t = topP[k]; topP[k] = b; b = t; }
            }
        }
    }
    double ri = topI[0], rp = topP[0];
#pragma unroll
    for (int k = 1; k < HM_KMAX; ++k)
        if (k == maxno - 1) { ri = topI[k]; rp = topP[k]; }
    I[m] = ri;
    PV[m] = rp;
}

int implausibility(const double* mean, const double* std_, long M, int E, const double* z, const double* v, int maxno,
                   double* I, double* PV, cudaStream_t st) {
    if (maxno < 1 || maxno > HM_KMAX || maxno > E) return -7;
    if (M == 0) return 0;
    implausibility_kernel<<<(unsigned)((M + 255) / 256), 256, 0, st>>>(mean, std_, M, E, z, v, maxno, I, PV);
    GPX_CHECK_LAUNCH();
    return 0;
}

}  // namespace gpx
