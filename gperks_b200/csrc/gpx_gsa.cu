// Saltelli-2010 sums for Sobol' indices with emulator uncertainty, on the device.
// Replaces the per-draw host loop of SobolGSA.estimate_Sobol_indices_with_emulator (GPErks/perks/gsa.py:66-88: n_draws joint
// draws of n (d + 2) values copied to the host, then scipy's estimators per draw) for designs that do not fit a joint draw
// (n (d + 2) = 1e7 points): with the predictive mean and std of every design point resident in HBM, draw k takes
//     f_k(x) = mean(x) + std(x) z_k(x),   z_k(x) ~ N(0, 1) independent per point (Philox4x32-10 keyed by (seed; point, draw)),
// and only the sums that scipy.stats.sobol_indices needs leave the kernel, per draw:
//     sum f_A, sum f_B, sum f_A^2, sum f_B^2, and per input p:  sum f_B D_p, sum D_p, sum D_p^2   with D_p = f_AB^(p) - f_A,
// from which (mu = mean of [f_A, f_B], var = their population variance)
//     S1_p = (sum f_B D_p - mu sum D_p) / (n var),      ST_p = sum D_p^2 / (2 n var)         [scipy centres by mu first].
// n_draws = 0 evaluates the sums of the posterior mean itself (one noise-free "draw").
// Layout: fA[n], fB[n], fAB[n][d] (row j = point A_j with coordinate p taken from B_j, p contiguous).
// Work split: chunks of GSA_JC rows x blocks of 128 draws; a thread owns one draw of one chunk and keeps its sums in registers
// (inputs in groups of 8 so that d <= 32 fits); partial[chunk][draw][4 + 3 d] is then reduced over the chunks in chunk order
// by a second kernel: deterministic, and independent of how the chunks are dealt over GPUs.
#include "gpx_common.cuh"

namespace gpx {

constexpr int GSA_JC = 4096;        // rows per chunk
constexpr int GSA_TB = 128;         // draws per block
constexpr int GSA_SJ = 32;          // rows staged in shared memory at a time
constexpr int GSA_PG = 8;           // inputs per register group

__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t (&out)[4]) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// standard normal number `slot` (0 .. d+1) of design row j in draw k: two Box-Muller pairs per Philox call
__device__ __forceinline__ double gsa_normal(unsigned long long seed, long j, int k, int slot) {
    uint32_t r[4];
    philox4x32_10((uint32_t)j, (uint32_t)((unsigned long long)j >> 32), (uint32_t)k, (uint32_t)(slot >> 2), (uint32_t)seed,
                  (uint32_t)(seed >> 32), r);
    const int pair = (slot >> 1) & 1;
    const float u1 = ((float)(r[2 * pair] >> 8) + 0.5f) * (1.0f / 16777216.0f);        // (0, 1)
    const float u2 = ((float)(r[2 * pair + 1] >> 8) + 0.5f) * (1.0f / 16777216.0f);
    const float rad = sqrtf(-2.0f * __logf(u1));
    float sn, cs;
    __sincosf(6.2831853071795865f * u2, &sn, &cs);
    return (double)(rad * ((slot & 1) ? sn : cs));
}

// partial[(chunk * n_draws_eff + k) * NS + s],  NS = 4 + 3 d
__global__ void __launch_bounds__(GSA_TB) saltelli_partial_kernel(const double* __restrict__ mA, const double* __restrict__ mB,
                                                                  const double* __restrict__ mAB, const double* __restrict__ sA,
                                                                  const double* __restrict__ sB, const double* __restrict__ sAB,
                                                                  long n, int d, long j_begin, int n_draws, unsigned long long seed,
                                                                  double* __restrict__ partial) {
    extern __shared__ double sm[];                       // [GSA_SJ][2 (d + 2)]: means then stds of A_j, B_j, AB_j[0..d)
    const int kd = n_draws > 0 ? n_draws : 1;
    const int k = blockIdx.y * GSA_TB + threadIdx.x;
    const long chunk = blockIdx.x;
    const long j0 = j_begin + chunk * GSA_JC;
    const long j1 = (j0 + GSA_JC < n) ? j0 + GSA_JC : n;
    const int w = d + 2;
    const bool noisy = n_draws > 0 && sA != nullptr;
    const int NS = 4 + 3 * d;
    double* out = partial + ((size_t)chunk * kd + (k < kd ? k : 0)) * NS;
    for (int p0 = 0; p0 < d; p0 += GSA_PG) {
        const int pn = (d - p0 < GSA_PG) ? (d - p0) : GSA_PG;
        double base[4] = {0.0, 0.0, 0.0, 0.0};
        double acc[GSA_PG][3];
#pragma unroll
        for (int q = 0; q < GSA_PG; ++q) { acc[q][0] = 0.0; acc[q][1] = 0.0; acc[q][2] = 0.0; }
        for (long js = j0; js < j1; js += GSA_SJ) {
            const int rows = (j1 - js < GSA_SJ) ? (int)(j1 - js) : GSA_SJ;
            __syncthreads();
            for (int e = threadIdx.x; e < rows * w; e += GSA_TB) {
                const int r = e / w, c = e - r * w;
                const long j = js + r;
                sm[r * 2 * w + c] = (c == 0) ? mA[j] : (c == 1) ? mB[j] : mAB[j * d + (c - 2)];
                sm[r * 2 * w + w + c] = !noisy ? 0.0 : (c == 0) ? sA[j] : (c == 1) ? sB[j] : sAB[j * d + (c - 2)];
            }
            __syncthreads();
            if (k < kd) {
                for (int r = 0; r < rows; ++r) {
                    const double* mu = sm + r * 2 * w;
                    const double* sd = mu + w;
                    const long j = js + r;
                    double fA = mu[0], fB = mu[1];
                    if (noisy) {
                        fA = fma(sd[0], gsa_normal(seed, j, k, 0), fA);
                        fB = fma(sd[1], gsa_normal(seed, j, k, 1), fB);
                    }
                    if (p0 == 0) {
                        base[0] += fA; base[1] += fB;
                        base[2] = fma(fA, fA, base[2]); base[3] = fma(fB, fB, base[3]);
                    }
#pragma unroll
                    for (int q = 0; q < GSA_PG; ++q) {
                        if (q < pn) {
                            double fAB = mu[2 + p0 + q];
                            if (noisy) fAB = fma(sd[2 + p0 + q], gsa_normal(seed, j, k, 2 + p0 + q), fAB);
                            const double D = fAB - fA;
                            acc[q][0] = fma(fB, D, acc[q][0]);
                            acc[q][1] += D;
                            acc[q][2] = fma(D, D, acc[q][2]);
                        }
                    }
                }
            }
        }
        if (k < kd) {
            if (p0 == 0) { out[0] = base[0]; out[1] = base[1]; out[2] = base[2]; out[3] = base[3]; }
#pragma unroll
            for (int q = 0; q < GSA_PG; ++q)
                if (q < pn) { out[4 + 3 * (p0 + q)] = acc[q][0]; out[5 + 3 * (p0 + q)] = acc[q][1]; out[6 + 3 * (p0 + q)] = acc[q][2]; }
        }
    }
}

// sums[k][s] = sum over chunks (in chunk order) of partial[chunk][k][s]
__global__ void saltelli_reduce_kernel(const double* __restrict__ partial, long n_chunks, long per_chunk, double* __restrict__ sums) {
    const long e = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= per_chunk) return;
    double t = 0.0;
    for (long c = 0; c < n_chunks; ++c) t += partial[c * per_chunk + e];
    sums[e] = t;
}

int saltelli_chunk_rows() { return GSA_JC; }

// rows [j_begin, j_end) of the design (j_begin a multiple of GSA_JC): partial[(chunk - j_begin / GSA_JC)][k][4 + 3 d]
int saltelli_partial(const double* mA, const double* mB, const double* mAB, const double* sA, const double* sB, const double* sAB,
                     long n, int d, long j_begin, long j_end, int n_draws, unsigned long long seed, double* partial, cudaStream_t st) {
    if (d < 1 || d > MAXD || n < 1 || j_begin % GSA_JC || j_begin < 0 || j_end > n || j_end < j_begin) return -2;
    if (j_end == j_begin) return 0;
    const int kd = n_draws > 0 ? n_draws : 1;
    const long chunks = (j_end - j_begin + GSA_JC - 1) / GSA_JC;
    const dim3 grid((unsigned)chunks, (unsigned)((kd + GSA_TB - 1) / GSA_TB));
    const size_t smem = (size_t)GSA_SJ * 2 * (d + 2) * 8;
    saltelli_partial_kernel<<<grid, GSA_TB, smem, st>>>(mA, mB, mAB, sA, sB, sAB, j_end, d, j_begin, n_draws, seed, partial);
    GPX_CHECK_LAUNCH();
    return 0;
}

int saltelli_reduce(const double* partial, long n_chunks, int d, int n_draws, double* sums, cudaStream_t st) {
    const long per_chunk = (long)(n_draws > 0 ? n_draws : 1) * (4 + 3 * d);
    saltelli_reduce_kernel<<<(unsigned)((per_chunk + 255) / 256), 256, 0, st>>>(partial, n_chunks, per_chunk, sums);
    GPX_CHECK_LAUNCH();
    return 0;
}

}  // namespace gpx
