// transform.cu -- direction-code reclassification (survey K1) and basin <-> raster transfer by linear index.
//
//   wmf_dir_reclass_scan / _apply   wmf_py/cu_py/basics.py:73-121 and cuencas_heavy.f90:1086-1119: the raster is
//                                   scanned once for the set of byte values it holds (a 256-bit presence map, plus
//                                   a flag for values beyond int8), the host layer picks identity / table / error
//                                   from that map, and a 256-entry lookup table is applied in a second pass.
//   wmf_index_scan / _repair        the range test and the "row*100 + col" repair of basics.py:271-277
//   wmf_scatter_last                out[idx[i]] = vals[i]; with duplicate indices the LAST i wins, which is what the
//                                   reference's sequential fancy assignment leaves (basics.py:279-280)
//   wmf_gather                      out[i] = flat[idx[i]] (basics.py:296-300)
#include "common.cuh"

namespace {

__device__ __forceinline__ int64_t load_int(const void *p, int eb, int64_t i)
{
    switch (eb) {
    case 1: return ((const int8_t *)p)[i];
    case 2: return ((const int16_t *)p)[i];
    case 4: return ((const int32_t *)p)[i];
    default: return ((const int64_t *)p)[i];
    }
}

__global__ void __launch_bounds__(256) k_reclass_scan(const void *__restrict__ in, int eb, int64_t n,
                                                      uint32_t *__restrict__ presence, int32_t *__restrict__ oor)
{
    __shared__ uint32_t s_pres[8];
    __shared__ int s_oor;
    if (threadIdx.x < 8) s_pres[threadIdx.x] = 0;
    if (threadIdx.x == 0) s_oor = 0;
    __syncthreads();
    uint32_t mine[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    bool far = false;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t v = load_int(in, eb, i);
        if (v < -128 || v > 127) { far = true; continue; }
        const int b = (int)v + 128;
#pragma unroll
        for (int q = 0; q < 8; q++) mine[q] |= (q == (b >> 5)) ? (1u << (b & 31)) : 0u;
    }
#pragma unroll
    for (int q = 0; q < 8; q++) {
        const uint32_t w = __reduce_or_sync(0xffffffffu, mine[q]);
        if ((threadIdx.x & 31) == 0 && w) atomicOr(&s_pres[q], w);
    }
    if (__any_sync(0xffffffffu, far) && (threadIdx.x & 31) == 0) s_oor = 1;
    __syncthreads();
    if (threadIdx.x < 8 && s_pres[threadIdx.x]) atomicOr(&presence[threadIdx.x], s_pres[threadIdx.x]);
    if (threadIdx.x == 0 && s_oor) atomicOr(oor, 1);
}

__global__ void __launch_bounds__(256) k_reclass_apply(const void *__restrict__ in, int eb, int64_t n,
                                                       const int64_t *__restrict__ lut, int keep_oor, int64_t oor_value,
                                                       int64_t *__restrict__ out)
{
    __shared__ int64_t s_lut[256];
    s_lut[threadIdx.x] = lut[threadIdx.x];
    __syncthreads();
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t v = load_int(in, eb, i);
        out[i] = (v < -128 || v > 127) ? (keep_oor ? v : oor_value) : s_lut[(int)v + 128];
    }
}

__global__ void k_index_scan(const int64_t *__restrict__ idx, int64_t n, int64_t ncell, int32_t *__restrict__ flag)
{
    bool bad = false;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t v = idx[i];
        bad = bad || v < 0 || v >= ncell;
    }
    if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) atomicOr(flag, 1);
}

// Python's floor division / modulo: the quotient rounds towards minus infinity, the remainder is never negative
__device__ __forceinline__ int64_t floor_div(int64_t a, int64_t b) { int64_t q = a / b; return (a % b != 0 && ((a < 0) != (b < 0))) ? q - 1 : q; }
__device__ __forceinline__ int64_t floor_mod(int64_t a, int64_t b) { int64_t r = a % b; return (r != 0 && ((r < 0) != (b < 0))) ? r + b : r; }

__global__ void k_index_repair(int64_t *__restrict__ idx, int64_t n, int64_t nx)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t v = idx[i];
        idx[i] = floor_div(v, 100) * nx + floor_mod(floor_mod(v, 100), nx);
    }
}

__global__ void k_scatter_owner(const int64_t *__restrict__ idx, int64_t n, int32_t *__restrict__ owner)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        atomicMax(&owner[idx[i]], (int32_t)i);
}

template <typename E>
__global__ void k_scatter_write(const E *__restrict__ vals, const int64_t *__restrict__ idx, int64_t n,
                                const int32_t *__restrict__ owner, E *__restrict__ out)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t j = idx[i];
        if (owner[j] == (int32_t)i) out[j] = vals[i];
    }
}

template <typename E>
__global__ void k_gather(const E *__restrict__ flat, const int64_t *__restrict__ idx, int64_t n, E *__restrict__ out)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = flat[idx[i]];
}

unsigned int grid_for(int64_t n)
{
    const int64_t b = (n + 255) / 256;
    return (unsigned int)(b < 1 ? 1 : (b > 148 * 16 ? 148 * 16 : b));
}

}  // namespace

extern "C" {

int wmf_dir_reclass_scan(const void *in, int elem_bytes, int64_t n, uint32_t *presence, int32_t *oor, void *stream)
{
    WMF_REQUIRE(in && presence && oor && n >= 0, WMF_E_ARG, "bad argument");
    WMF_REQUIRE(elem_bytes == 1 || elem_bytes == 2 || elem_bytes == 4 || elem_bytes == 8, WMF_E_ARG, "element size must be 1, 2, 4 or 8");
    cudaStream_t st = as_stream(stream);
    WMF_CUDA_CHECK(cudaMemsetAsync(presence, 0, 8 * sizeof(uint32_t), st));
    WMF_CUDA_CHECK(cudaMemsetAsync(oor, 0, sizeof(int32_t), st));
    if (n == 0) return 0;
    k_reclass_scan<<<grid_for(n), 256, 0, st>>>(in, elem_bytes, n, presence, oor);
    WMF_LAUNCH_CHECK();
    return 0;
}

int wmf_dir_reclass_apply(const void *in, int elem_bytes, int64_t n, const int64_t *lut, int keep_oor, int64_t oor_value,
                          int64_t *out, void *stream)
{
    WMF_REQUIRE(in && lut && out && n >= 0, WMF_E_ARG, "bad argument");
    WMF_REQUIRE(elem_bytes == 1 || elem_bytes == 2 || elem_bytes == 4 || elem_bytes == 8, WMF_E_ARG, "element size must be 1, 2, 4 or 8");
    if (n == 0) return 0;
    k_reclass_apply<<<grid_for(n), 256, 0, as_stream(stream)>>>(in, elem_bytes, n, lut, keep_oor, oor_value, out);
    WMF_LAUNCH_CHECK();
    return 0;
}

int wmf_index_scan(const int64_t *idx, int64_t n, int64_t ncell, int32_t *flag, void *stream)
{
    WMF_REQUIRE(idx && flag && n >= 0, WMF_E_ARG, "bad argument");
    cudaStream_t st = as_stream(stream);
    WMF_CUDA_CHECK(cudaMemsetAsync(flag, 0, sizeof(int32_t), st));
    if (n == 0) return 0;
    k_index_scan<<<grid_for(n), 256, 0, st>>>(idx, n, ncell, flag);
    WMF_LAUNCH_CHECK();
    return 0;
}

int wmf_index_repair(int64_t *idx, int64_t n, int64_t nx, void *stream)
{
    WMF_REQUIRE(idx && n >= 0 && nx > 0, WMF_E_ARG, "bad argument");
    if (n == 0) return 0;
    k_index_repair<<<grid_for(n), 256, 0, as_stream(stream)>>>(idx, n, nx);
    WMF_LAUNCH_CHECK();
    return 0;
}

int wmf_scatter_last(const void *vals, int elem_bytes, const int64_t *idx, int64_t n, void *out, int32_t *owner,
                     int64_t ncell, void *stream)
{
    WMF_REQUIRE(vals && idx && out && owner && n >= 0 && ncell >= 0, WMF_E_ARG, "bad argument");
    WMF_REQUIRE(n < ((int64_t)1 << 31), WMF_E_TOO_LARGE, "more than 2^31-1 values");
    WMF_REQUIRE(elem_bytes == 1 || elem_bytes == 2 || elem_bytes == 4 || elem_bytes == 8, WMF_E_ARG, "element size must be 1, 2, 4 or 8");
    cudaStream_t st = as_stream(stream);
    WMF_CUDA_CHECK(cudaMemsetAsync(owner, 0xFF, (size_t)ncell * sizeof(int32_t), st));
    if (n == 0) return 0;
    k_scatter_owner<<<grid_for(n), 256, 0, st>>>(idx, n, owner);
    switch (elem_bytes) {
    case 1: k_scatter_write<uint8_t><<<grid_for(n), 256, 0, st>>>((const uint8_t *)vals, idx, n, owner, (uint8_t *)out); break;
    case 2: k_scatter_write<uint16_t><<<grid_for(n), 256, 0, st>>>((const uint16_t *)vals, idx, n, owner, (uint16_t *)out); break;
    case 4: k_scatter_write<uint32_t><<<grid_for(n), 256, 0, st>>>((const uint32_t *)vals, idx, n, owner, (uint32_t *)out); break;
    default: k_scatter_write<uint64_t><<<grid_for(n), 256, 0, st>>>((const uint64_t *)vals, idx, n, owner, (uint64_t *)out); break;
    }
    WMF_LAUNCH_CHECK();
    return 0;
}

int wmf_gather(const void *flat, int elem_bytes, const int64_t *idx, int64_t n, void *out, void *stream)
{
    WMF_REQUIRE(flat && idx && out && n >= 0, WMF_E_ARG, "bad argument");
    WMF_REQUIRE(elem_bytes == 1 || elem_bytes == 2 || elem_bytes == 4 || elem_bytes == 8, WMF_E_ARG, "element size must be 1, 2, 4 or 8");
    if (n == 0) return 0;
    cudaStream_t st = as_stream(stream);
    switch (elem_bytes) {
    case 1: k_gather<uint8_t><<<grid_for(n), 256, 0, st>>>((const uint8_t *)flat, idx, n, (uint8_t *)out); break;
    case 2: k_gather<uint16_t><<<grid_for(n), 256, 0, st>>>((const uint16_t *)flat, idx, n, (uint16_t *)out); break;
    case 4: k_gather<uint32_t><<<grid_for(n), 256, 0, st>>>((const uint32_t *)flat, idx, n, (uint32_t *)out); break;
    default: k_gather<uint64_t><<<grid_for(n), 256, 0, st>>>((const uint64_t *)flat, idx, n, (uint64_t *)out); break;
    }
    WMF_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
