// Host build of gperks_b200/csrc/gpx_math.cuh for tests/test_math_host.py (CPU sweep against libm).
#include "../gperks_b200/csrc/gpx_math.cuh"
extern "C" void gpx_host_exp_neg(const double* y, double* out, long n) {
    for (long i = 0; i < n; ++i) out[i] = gpx::exp_neg(y[i]);
}
extern "C" void gpx_host_sqrt_pos(const double* x, double* out, long n) {
    for (long i = 0; i < n; ++i) out[i] = gpx::sqrt_pos(x[i]);
}
