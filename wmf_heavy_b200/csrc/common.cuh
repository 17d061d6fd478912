// common.cuh -- shared helpers for libwmfb200.so (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/wmf_b200.h"

void wmf_set_error(const char *fmt, ...);

#define WMF_CUDA_CHECK(expr)                                                                        \
    do {                                                                                            \
        cudaError_t e_ = (expr);                                                                    \
        if (e_ != cudaSuccess) {                                                                    \
            wmf_set_error("%s:%d %s: %s", __FILE__, __LINE__, #expr, cudaGetErrorString(e_));       \
            return (int)e_;                                                                         \
        }                                                                                           \
    } while (0)
#define WMF_LAUNCH_CHECK() WMF_CUDA_CHECK(cudaGetLastError())
#define WMF_REQUIRE(cond, code, msg)                                                                \
    do {                                                                                            \
        if (!(cond)) {                                                                              \
            wmf_set_error("%s: %s", __func__, msg);                                                 \
            return (code);                                                                          \
        }                                                                                           \
    } while (0)

static inline size_t wmf_align(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

// Bump allocator over a caller-owned workspace.
struct WsCursor {
    char *base;
    size_t off, cap;
    bool ok;
    WsCursor(void *p, size_t n) : base((char *)p), off(0), cap(n), ok(true) {}
    template <typename T>
    T *take(size_t count) {
        size_t bytes = wmf_align(count * sizeof(T));
        if (off + bytes > cap) { ok = false; return nullptr; }
        T *r = (T *)(base + off);
        off += bytes;
        return r;
    }
};

// ---------------------------------------------------------------------------------------------
// D8 direction codes.  Both encodings are normalised to k in 0..7 (clockwise from East, the
// wmf_py numbering, basics.py:37-46) or 8 = "no receiver".
//   internal: k = code for 0..7, anything else -> none (basics.py:151-153)
//   keypad  : 6->0 E, 3->1 SE, 2->2 S, 1->3 SW, 4->4 W, 7->5 NW, 8->6 N, 9->7 NE; 5, <=0, >9 -> none
//             (cuencas_heavy.f90:1338-1362, rows increase downwards)
// ---------------------------------------------------------------------------------------------
template <int ENC>
__host__ __device__ __forceinline__ int d8_norm(int code)
{
    if (ENC == WMF_ENC_INTERNAL) return ((unsigned)code <= 7u) ? code : 8;
    // nibble LUT, code 0..9 -> k
    return ((unsigned)code <= 9u) ? (int)((0x7650841238ull >> (4 * code)) & 0xF) : 8;
}
// row / column offset of direction k (0..7)
__host__ __device__ __forceinline__ int d8_dr(int k) { return (int)((0x01A9u >> (2 * k)) & 3u) - 1; }
__host__ __device__ __forceinline__ int d8_dc(int k) { return (int)((0x901Au >> (2 * k)) & 3u) - 1; }

__device__ __forceinline__ bool cell_active(const uint8_t *__restrict__ mask, int64_t i)
{
    return mask == nullptr || mask[i] != 0;
}

// Linear index of the receiver of cell (r, c), or -1: valid code, in bounds, receiver active
// (basics.py:147-157).  The caller has checked that (r, c) itself is active.
template <int ENC>
__device__ __forceinline__ int64_t raster_recv(const int8_t *__restrict__ dir, const uint8_t *__restrict__ mask,
                                               int64_t ny, int64_t nx, int64_t r, int64_t c)
{
    int k = d8_norm<ENC>(dir[r * nx + c]);
    if (k == 8) return -1;
    int64_t rr = r + d8_dr(k), cc = c + d8_dc(k);
    if (rr < 0 || rr >= ny || cc < 0 || cc >= nx) return -1;
    int64_t j = rr * nx + cc;
    return cell_active(mask, j) ? j : -1;
}

// Number of active neighbours draining into active cell (r, c).
template <int ENC>
__device__ __forceinline__ int raster_indegree(const int8_t *__restrict__ dir, const uint8_t *__restrict__ mask,
                                               int64_t ny, int64_t nx, int64_t r, int64_t c)
{
    int n = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        int64_t rr = r - d8_dr(k), cc = c - d8_dc(k);   // the neighbour that would reach us with code k
        if (rr < 0 || rr >= ny || cc < 0 || cc >= nx) continue;
        int64_t j = rr * nx + cc;
        if (cell_active(mask, j) && d8_norm<ENC>(dir[j]) == k) n++;
    }
    return n;
}

static inline cudaStream_t as_stream(void *s) { return (cudaStream_t)s; }

// grid size for a 1-thread-per-item launch
static inline unsigned int blocks_for(int64_t n, int threads) { return (unsigned int)((n + threads - 1) / threads); }

// ---------------------------------------------------------------------------------------------
// forest engine (forest.cu): acc[i] = w[i] + sum over children acc[child], for nodes whose whole
// upstream is acyclic; other nodes keep the partial sum of finished inflows without their own
// weight (the reference's cycle semantics, basics.py:166-175).
//   parent: int32 [n], -1 = none.  w: weights (nullptr = 1).  acc: out.  cnt: int32 [n] scratch.
// ---------------------------------------------------------------------------------------------
// blocked (nullable): nodes that must never finalise (their own value is itself unresolved).
// After the call cnt[i] is -1 (leaf) or 0 for finalised nodes and > 0 for unresolved ones.
int forest_accum_i32(const int32_t *parent, const int32_t *w, int32_t *acc, int32_t *cnt, int64_t n, cudaStream_t st,
                     const uint8_t *blocked = nullptr);
int forest_accum_i64(const int32_t *parent, const int32_t *w, int64_t *acc, int32_t *cnt, int64_t n, cudaStream_t st,
                     const uint8_t *blocked = nullptr);
// acyclic forests only: the same sums in log2(depth) rounds of scatter-adds (scratch: 3 n int32 + 64)
int forest_subtree_sums_i32(const int32_t *parent, const int32_t *w, int32_t *acc, int32_t *scratch, int64_t n, cudaStream_t st);
int forest_accum_i64w(const int32_t *parent, const int64_t *w, int64_t *acc, int32_t *cnt, int64_t n, cudaStream_t st,
                      const uint8_t *blocked = nullptr);
