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
implausibility_kernel(const double* __restrict__ mean, const double* __restrict__ std_, long M, long ld, int E,
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
            mu[u] = __ldcs(pm + (long)(e + u) * ld);      // streamed once: evict-first
            sd[u] = __ldcs(ps + (long)(e + u) * ld);
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
        hm_terms(__ldcs(pm + (long)e * ld), __ldcs(ps + (long)e * ld), __ldg(z + e), __ldg(v + e), a, b);
        hm_insert<MAXNO>(topI, topP, a, b);
    }
    __stcs(I + m, topI[MAXNO - 1]);
    __stcs(PV + m, topP[MAXNO - 1]);
}

// mean/std: emulator-major with leading dimension ld >= M (ld = M: the per-emulator predict outputs; ld = chunk size: the
// chunk-local buffers of the multi-emulator wave, gpx_predict.cu)
int implausibility(const double* mean, const double* std_, long M, long ld, int E, const double* z, const double* v, int maxno,
                   double* I, double* PV, cudaStream_t st) {
    if (maxno < 1 || maxno > HM_KMAX || maxno > E || ld < M) return -7;
    if (M == 0) return 0;
    const unsigned grid = (unsigned)((M + 255) / 256);
#define GPX_HM_CASE(K) case K: implausibility_kernel<K><<<grid, 256, 0, st>>>(mean, std_, M, ld, E, z, v, I, PV); break;
    switch (maxno) {
        GPX_HM_CASE(1) GPX_HM_CASE(2) GPX_HM_CASE(3) GPX_HM_CASE(4)
        GPX_HM_CASE(5) GPX_HM_CASE(6) GPX_HM_CASE(7) GPX_HM_CASE(8)
    }
#undef GPX_HM_CASE
    GPX_CHECK_LAUNCH();
    return 0;
}

// ---- deterministic stream compaction: the ascending indices m with I[m] < cutoff -----------------------------------------
// Replaces  np.where(I < cutoff)  of Wave.find_regions (GPErks/perks/history_matching.py:66-77) so that only the kept
// indices (8 B each) ever cross PCIe.  Three launches, no atomics: per-block counts (4096 entries per block) -> exclusive
// scan of the counts by one block -> scatter.  Inside a block, entry t + 256 j of thread t is ranked by a ballot / popc scan.
constexpr int CP_THREADS = 256;
constexpr int CP_ITEMS = 16;
constexpr int CP_TILE = CP_THREADS * CP_ITEMS;

__device__ __forceinline__ int block_exclusive_scan_flag(int flag, int* warp_tot, int& block_total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned bal = __ballot_sync(0xffffffffu, flag);
    const int in_warp = __popc(bal & ((1u << lane) - 1u));
    if (lane == 0) warp_tot[warp] = __popc(bal);
    __syncthreads();
    int before = 0, total = 0;
#pragma unroll
    for (int w = 0; w < CP_THREADS / 32; ++w) {
        const int c = warp_tot[w];
        before += (w < warp) ? c : 0;
        total += c;
    }
    __syncthreads();
    block_total = total;
    return before + in_warp;
}

template <bool SCATTER>
__global__ void __launch_bounds__(CP_THREADS) compact_below_kernel(const double* __restrict__ I, long M, double cutoff,
                                                                   long long* __restrict__ block_off, long long* __restrict__ idx) {
    __shared__ int warp_tot[CP_THREADS / 32];
    const long base = (long)blockIdx.x * CP_TILE;
    long long run = SCATTER ? block_off[blockIdx.x] : 0;
#pragma unroll 1
    for (int j = 0; j < CP_ITEMS; ++j) {
        const long m = base + (long)j * CP_THREADS + threadIdx.x;
        const int keep = (m < M) && (I[m] < cutoff);
        int total;
        const int rank = block_exclusive_scan_flag(keep, warp_tot, total);
        if (SCATTER && keep) idx[run + rank] = m;
        run += total;
    }
    if (!SCATTER && threadIdx.x == 0) block_off[blockIdx.x] = run;
}

// block_off[0 .. nblk) (counts) -> exclusive prefix sums in place; count_out = total
__global__ void __launch_bounds__(1024) compact_scan_kernel(long long* __restrict__ block_off, long nblk, long long* __restrict__ count_out) {
    __shared__ long long wsum[32];
    __shared__ long long carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (long b0 = 0; b0 < nblk; b0 += 1024) {
        const long b = b0 + threadIdx.x;
        const long long c = b < nblk ? block_off[b] : 0;
        long long x = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const long long y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) wsum[warp] = x;
        __syncthreads();
        if (warp == 0) {
            long long w = wsum[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const long long y = __shfl_up_sync(0xffffffffu, w, o);
                if (lane >= o) w += y;
            }
            wsum[lane] = w;
        }
        __syncthreads();
        const long long carry = carry_s;
        const long long incl = x + (warp > 0 ? wsum[warp - 1] : 0);
        if (b < nblk) block_off[b] = carry + incl - c;
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = carry + incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) *count_out = carry_s;
}

size_t compact_below_workspace_bytes(long M) { return (size_t)((M + CP_TILE - 1) / CP_TILE + 1) * sizeof(long long); }

int compact_below(const double* I, long M, double cutoff, long long* idx, long long* count, void* workspace, size_t workspace_bytes,
                  cudaStream_t st) {
    if (M < 0) return -2;
    if (workspace_bytes < compact_below_workspace_bytes(M)) return -7;
    if (M == 0) { cudaMemsetAsync(count, 0, sizeof(long long), st); return 0; }
    long long* block_off = reinterpret_cast<long long*>(workspace);
    const long nblk = (M + CP_TILE - 1) / CP_TILE;
    compact_below_kernel<false><<<(unsigned)nblk, CP_THREADS, 0, st>>>(I, M, cutoff, block_off, idx);
    compact_scan_kernel<<<1, 1024, 0, st>>>(block_off, nblk, count);
    compact_below_kernel<true><<<(unsigned)nblk, CP_THREADS, 0, st>>>(I, M, cutoff, block_off, idx);
    GPX_CHECK_LAUNCH();
    return 0;
}

}  // namespace gpx
