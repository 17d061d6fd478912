// d8_sweep.cu -- placeholder, replaced below
#include "common.cuh"
size_t d8_walk_workspace(int64_t ny, int64_t nx);
int d8_walk(const int8_t *dir, const uint8_t *mask, void *acc, int acc_dtype, int64_t ny, int64_t nx, int encoding,
            void *ws, size_t ws_bytes, cudaStream_t st);
size_t d8_sweep_workspace(int64_t ny, int64_t nx, int acc_dtype) { return d8_walk_workspace(ny, nx); }
int d8_sweep(const int8_t *dir, const uint8_t *mask, void *acc, int acc_dtype, int64_t ny, int64_t nx, int encoding,
             void *ws, size_t ws_bytes, cudaStream_t st)
{
    return d8_walk(dir, mask, acc, acc_dtype, ny, nx, encoding, ws, ws_bytes, st);
}
