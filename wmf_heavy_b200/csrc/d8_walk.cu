// d8_walk.cu -- full-raster D8 accumulation, general algorithm (WMF_ALGO_WALK).
// Replaces wmf_py/cu_py/basics.py:124-177 for ANY input (invalid codes, masks, pits, cycles):
//   K1 k_indegree : per cell, pull-count the active neighbours that drain into it (coalesced int8
//                   reads, no atomics) -> one byte per cell (0xFF marks a leaf)
//   K2 k_walk     : every leaf starts a walker; a walker adds its finished sum to its receiver
//                   (L2 atomic), decrements the receiver's pending-donor byte, and carries on only
//                   if it was the last donor ("follow the chain while you are the last donor").
// No level barrier, O(cells) work, bit-exact for any schedule because the adds are integer.
// Cells on/below a drainage cycle are never finalised and keep the partial sum of their finished
// inflows without their own +1 -- exactly the reference's queue semantics (basics.py:166-175).
#include "common.cuh"

namespace {

template <int ENC>
__global__ void k_indegree(const int8_t *__restrict__ dir, const uint8_t *__restrict__ mask,
                           uint8_t *__restrict__ cnt, int64_t ny, int64_t nx)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ny * nx) return;
    uint8_t v = 0;                                   // inactive cells: never touched again
    if (cell_active(mask, i)) {
        int64_t r = i / nx, c = i - r * nx;
        int d = raster_indegree<ENC>(dir, mask, ny, nx, r, c);
        v = d == 0 ? 0xFF : (uint8_t)d;
    }
    cnt[i] = v;
}

template <typename T>
__device__ __forceinline__ T atom_add(T *p, T v);
template <>
__device__ __forceinline__ int32_t atom_add<int32_t>(int32_t *p, int32_t v) { return atomicAdd(p, v); }
template <>
__device__ __forceinline__ int64_t atom_add<int64_t>(int64_t *p, int64_t v)
{
    return (int64_t)atomicAdd((unsigned long long *)p, (unsigned long long)v);
}

template <int ENC, typename T>
__global__ void k_walk(const int8_t *__restrict__ dir, const uint8_t *__restrict__ mask, T *acc,
                       uint32_t *cnt32, int64_t ny, int64_t nx)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ny * nx) return;
    if (((const uint8_t *)cnt32)[i] != 0xFF) return;      // leaves only (0xFF is never produced by a decrement)
    T v = 1;
    acc[i] = 1;
    int64_t cur = i;
    while (true) {
        int64_t r = cur / nx, c = cur - r * nx;
        int64_t j = raster_recv<ENC>(dir, mask, ny, nx, r, c);
        if (j < 0) break;
        atom_add<T>(&acc[j], v);
        __threadfence();
        unsigned sh = 8u * (unsigned)(j & 3);
        uint32_t old = atomicSub(&cnt32[j >> 2], 1u << sh);
        if (((old >> sh) & 0xFFu) != 1u) break;            // not the last donor
        __threadfence();
        v = atom_add<T>(&acc[j], (T)1) + 1;                 // own cell, now final
        cur = j;
    }
}

__global__ void k_i64_to_f64(int64_t *buf, int64_t n)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        int64_t v = buf[i];
        ((double *)buf)[i] = (double)v;
    }
}

template <int ENC>
int run_walk(const int8_t *dir, const uint8_t *mask, void *acc, int acc_dtype, int64_t ny, int64_t nx,
             void *ws, size_t ws_bytes, cudaStream_t st)
{
    int64_t n = ny * nx;
    WsCursor cur(ws, ws_bytes);
    uint8_t *cnt = cur.take<uint8_t>((size_t)((n + 3) / 4 * 4));
    WMF_REQUIRE(cur.ok, WMF_E_WORKSPACE, "workspace too small");
    const int th = 256;
    size_t esz = acc_dtype == WMF_ACC_I32 ? 4 : 8;
    WMF_CUDA_CHECK(cudaMemsetAsync(acc, 0, esz * n, st));
    k_indegree<ENC><<<blocks_for(n, th), th, 0, st>>>(dir, mask, cnt, ny, nx);
    if (acc_dtype == WMF_ACC_I32)
        k_walk<ENC, int32_t><<<blocks_for(n, th), th, 0, st>>>(dir, mask, (int32_t *)acc, (uint32_t *)cnt, ny, nx);
    else
        k_walk<ENC, int64_t><<<blocks_for(n, th), th, 0, st>>>(dir, mask, (int64_t *)acc, (uint32_t *)cnt, ny, nx);
    if (acc_dtype == WMF_ACC_F64) k_i64_to_f64<<<blocks_for(n, th), th, 0, st>>>((int64_t *)acc, n);
    WMF_LAUNCH_CHECK();
    return 0;
}

}  // namespace

size_t d8_walk_workspace(int64_t ny, int64_t nx) { return wmf_align((size_t)(ny * nx + 4)) + 256; }

int d8_walk(const int8_t *dir, const uint8_t *mask, void *acc, int acc_dtype, int64_t ny, int64_t nx, int encoding,
            void *ws, size_t ws_bytes, cudaStream_t st)
{
    if (encoding == WMF_ENC_INTERNAL)
        return run_walk<WMF_ENC_INTERNAL>(dir, mask, acc, acc_dtype, ny, nx, ws, ws_bytes, st);
    return run_walk<WMF_ENC_KEYPAD>(dir, mask, acc, acc_dtype, ny, nx, ws, ws_bytes, st);
}

int d8_convert_i64_to_f64(void *buf, int64_t n, cudaStream_t st)
{
    if (n > 0) k_i64_to_f64<<<blocks_for(n, 256), 256, 0, st>>>((int64_t *)buf, n);
    WMF_LAUNCH_CHECK();
    return 0;
}
