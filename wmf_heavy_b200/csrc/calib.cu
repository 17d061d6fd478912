// calib.cu -- fitness of a NSGA-II population on the device (C1: wmf_heavy/wmfv2.py:2060-2106, 2194-2198).
// The reference evaluates every individual's hydrograph with two objectives, __eval_nash__ and __eval_q_pico__, whose
// bodies are not in the repository; they are taken in their standard form here:
//   Nash-Sutcliffe efficiency  1 - sum (o - s)^2 / sum (o - mean o)^2          (to be maximised)
//   peak-flow error            |max s - max o| / max o                         (to be minimised)
// over the steps where the observation is finite.  One block per parameter set, fp64 sums in a fixed order.
#include "common.cuh"

namespace {

constexpr int FT = 256;

__global__ void __launch_bounds__(FT) k_fitness(const float *__restrict__ qsim, const float *__restrict__ qobs, int64_t T,
                                                double *__restrict__ out)
{
    __shared__ double s_a[FT], s_b[FT], s_c[FT], s_d[FT];
    __shared__ double s_mean;
    const float *s = qsim + (int64_t)blockIdx.x * T;
    double so = 0.0, n = 0.0;
    for (int64_t t = threadIdx.x; t < T; t += FT) {
        const float o = qobs[t];
        if (isfinite(o)) { so += (double)o; n += 1.0; }
    }
    s_a[threadIdx.x] = so; s_b[threadIdx.x] = n;
    __syncthreads();
    for (int k = FT / 2; k > 0; k >>= 1) {
        if (threadIdx.x < k) { s_a[threadIdx.x] += s_a[threadIdx.x + k]; s_b[threadIdx.x] += s_b[threadIdx.x + k]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) s_mean = s_b[0] > 0.0 ? s_a[0] / s_b[0] : 0.0;
    __syncthreads();
    const double mean = s_mean;
    double num = 0.0, den = 0.0, ms = -1.0e300, mo = -1.0e300;
    for (int64_t t = threadIdx.x; t < T; t += FT) {
        const float o = qobs[t];
        if (!isfinite(o)) continue;
        const double d = (double)o - (double)s[t], e = (double)o - mean;
        num += d * d; den += e * e;
        ms = fmax(ms, (double)s[t]); mo = fmax(mo, (double)o);
    }
    __syncthreads();
    s_a[threadIdx.x] = num; s_b[threadIdx.x] = den; s_c[threadIdx.x] = ms; s_d[threadIdx.x] = mo;
    __syncthreads();
    for (int k = FT / 2; k > 0; k >>= 1) {
        if (threadIdx.x < k) {
            s_a[threadIdx.x] += s_a[threadIdx.x + k]; s_b[threadIdx.x] += s_b[threadIdx.x + k];
            s_c[threadIdx.x] = fmax(s_c[threadIdx.x], s_c[threadIdx.x + k]); s_d[threadIdx.x] = fmax(s_d[threadIdx.x], s_d[threadIdx.x + k]);
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        out[2 * blockIdx.x] = s_b[0] > 0.0 ? 1.0 - s_a[0] / s_b[0] : -INFINITY;
        out[2 * blockIdx.x + 1] = s_d[0] > 0.0 ? fabs(s_c[0] - s_d[0]) / s_d[0] : INFINITY;
    }
}

}  // namespace

extern "C" int wmf_fitness_nash_peak(const float *qsim, const float *qobs, int32_t P, int64_t T, double *fitness, void *stream)
{
    WMF_REQUIRE(qsim && qobs && fitness && P >= 0 && T > 0, WMF_E_ARG, "bad argument");
    if (P == 0) return 0;
    k_fitness<<<P, FT, 0, as_stream(stream)>>>(qsim, qobs, T, fitness);
    WMF_LAUNCH_CHECK();
    return 0;
}
