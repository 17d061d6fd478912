// api.cu -- error reporting, device info, the D8 accumulation dispatcher and synthetic inputs.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

// d8_walk.cu
size_t d8_walk_workspace(int64_t ny, int64_t nx);
int d8_walk(const int8_t *dir, const uint8_t *mask, void *acc, int acc_dtype, int64_t ny, int64_t nx, int encoding,
            void *ws, size_t ws_bytes, cudaStream_t st);
// d8_sweep.cu
size_t d8_sweep_workspace(int64_t ny, int64_t nx, int acc_dtype);
int d8_sweep(const int8_t *dir, const uint8_t *mask, void *acc, int acc_dtype, int64_t ny, int64_t nx, int encoding,
             void *ws, size_t ws_bytes, cudaStream_t st);

int d8_sweep_tile_classes(const void *ws, size_t ws_bytes, int64_t ny, int64_t nx, int64_t *counts8, cudaStream_t st);

static thread_local char g_err[512] = "";

void wmf_set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

namespace {

__host__ __device__ __forceinline__ int scheidegger_code(uint64_t seed, uint64_t idx, int enc)
{
    // splitmix64 finaliser keyed by (seed, global linear index); top two bits pick the direction
    uint64_t x = seed ^ (idx * 0x9E3779B97F4A7C15ull);
    x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ull;
    x ^= x >> 27; x *= 0x94D049BB133111EBull;
    x ^= x >> 31;
    int u = (int)(x >> 62);                     // 0: E (1/4), 1,2: SE (1/2), 3: S (1/4)
    int k = u == 0 ? 0 : (u == 3 ? 2 : 1);
    if (enc == WMF_ENC_INTERNAL) return k;
    return k == 0 ? 6 : (k == 1 ? 3 : 2);
}

__global__ void k_scheidegger(int8_t *dir, int64_t ny, int64_t nx, int64_t row0, int64_t col0, int64_t nx_total,
                              uint64_t seed, int enc)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ny * nx) return;
    int64_t r = i / nx, c = i - r * nx;
    dir[i] = (int8_t)scheidegger_code(seed, (uint64_t)((row0 + r) * nx_total + col0 + c), enc);
}

}  // namespace

extern "C" {

const char *wmf_last_error(void) { return g_err; }
int wmf_version(void) { return 100; }

int wmf_device_info(int64_t *out4)
{
    WMF_REQUIRE(out4, WMF_E_ARG, "null pointer");
    int dev = 0;
    WMF_CUDA_CHECK(cudaGetDevice(&dev));
    cudaDeviceProp p;
    WMF_CUDA_CHECK(cudaGetDeviceProperties(&p, dev));
    out4[0] = p.multiProcessorCount; out4[1] = p.major; out4[2] = p.minor; out4[3] = p.l2CacheSize;
    return 0;
}

size_t wmf_d8_accum_workspace(int64_t ny, int64_t nx, int acc_dtype, int algo)
{
    size_t a = d8_walk_workspace(ny, nx), b = d8_sweep_workspace(ny, nx, acc_dtype);
    if (algo == WMF_ALGO_WALK) return a;
    if (algo == WMF_ALGO_SWEEP) return b;
    return a > b ? a : b;
}

int wmf_d8_accum(const int8_t *dir, const uint8_t *mask, void *acc, int acc_dtype, int64_t ny, int64_t nx,
                 int encoding, int algo, void *workspace, size_t ws_bytes, void *stream)
{
    WMF_REQUIRE(ny >= 0 && nx >= 0, WMF_E_ARG, "negative size");
    if (ny == 0 || nx == 0) return 0;
    WMF_REQUIRE(dir && acc, WMF_E_ARG, "null pointer");
    WMF_REQUIRE(encoding == WMF_ENC_INTERNAL || encoding == WMF_ENC_KEYPAD, WMF_E_ARG, "unknown encoding");
    WMF_REQUIRE(acc_dtype == WMF_ACC_I32 || acc_dtype == WMF_ACC_I64 || acc_dtype == WMF_ACC_F64, WMF_E_ARG,
                "unknown acc_dtype");
    WMF_REQUIRE(algo == WMF_ALGO_AUTO || algo == WMF_ALGO_WALK || algo == WMF_ALGO_SWEEP, WMF_E_ARG, "unknown algo");
    WMF_REQUIRE(!(acc_dtype == WMF_ACC_I32 && ny * nx > 0x7fffffffLL), WMF_E_TOO_LARGE,
                "int32 accumulation can overflow beyond 2^31-1 cells: use WMF_ACC_I64");
    WMF_REQUIRE(workspace || ws_bytes == 0, WMF_E_ARG, "null workspace");
    if (algo == WMF_ALGO_AUTO) algo = WMF_ALGO_SWEEP;
    if (algo == WMF_ALGO_WALK)
        return d8_walk(dir, mask, acc, acc_dtype, ny, nx, encoding, workspace, ws_bytes, as_stream(stream));
    return d8_sweep(dir, mask, acc, acc_dtype, ny, nx, encoding, workspace, ws_bytes, as_stream(stream));
}

int wmf_d8_accum_tile_classes(const void *workspace, size_t ws_bytes, int64_t ny, int64_t nx, int64_t *counts8_host, void *stream)
{
    WMF_REQUIRE(workspace && counts8_host && ny > 0 && nx > 0, WMF_E_ARG, "bad argument");
    return d8_sweep_tile_classes(workspace, ws_bytes, ny, nx, counts8_host, as_stream(stream));
}

int wmf_synth_scheidegger(int8_t *dir, int64_t ny, int64_t nx, int64_t row0, int64_t col0, int64_t nx_total,
                          uint64_t seed, int encoding, void *stream)
{
    WMF_REQUIRE(dir && ny >= 0 && nx >= 0, WMF_E_ARG, "bad argument");
    if (ny * nx == 0) return 0;
    k_scheidegger<<<blocks_for(ny * nx, 256), 256, 0, as_stream(stream)>>>(dir, ny, nx, row0, col0, nx_total, seed, encoding);
    WMF_LAUNCH_CHECK();
    return 0;
}

void wmf_synth_scheidegger_host(int8_t *dir, int64_t ny, int64_t nx, int64_t row0, int64_t col0, int64_t nx_total,
                                uint64_t seed, int encoding)
{
    for (int64_t r = 0; r < ny; r++)
        for (int64_t c = 0; c < nx; c++)
            dir[r * nx + c] = (int8_t)scheidegger_code(seed, (uint64_t)((row0 + r) * nx_total + col0 + c), encoding);
}

}  // extern "C"
