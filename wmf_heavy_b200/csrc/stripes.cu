// stripes.cu -- the boundary forest of the row-striped accumulation.
// After the one all-gather every rank holds the packed boundary records of all G stripes:
//   recs int64 [G][2][2][nx]: plane 0 = exit_acc; plane 1 = link (low 32 bits) | k << 32 | unres << 40
// (k: normalised D8 code, 8 none, 9 inactive; link: boundary index side*nx+col where flow entering at
// the cell leaves its stripe again, -1 if it ends inside).  Node id = (g*2 + side)*nx + col.
//   k_boundary_build  parent of every stripe exit = link of the cell it drains to in the next stripe; arrival counters
//   k_boundary_walk   totals of all stripe exits (2*G*nx nodes; redundant on every rank, it is tiny)
//   k_boundary_pull   what enters each boundary cell of stripe `me` = sum over its <= 3 donors across
//                     the stripe edge (pull: deterministic, no atomics)
#include "common.cuh"

namespace {

__device__ __forceinline__ int rec_k(int64_t m) { return (int)((m >> 32) & 0xFF); }
__device__ __forceinline__ int32_t rec_link(int64_t m) { return (int32_t)(uint32_t)(m & 0xFFFFFFFFll); }
__device__ __forceinline__ bool rec_unres(int64_t m) { return (m >> 40) & 1; }

// is (g, side, k) a stripe exit, and towards which stripe?
__device__ __forceinline__ bool exit_dir(int g, int side, int k, int G, int &g2)
{
    if (side == 0 && k >= 5 && k <= 7 && g > 0) { g2 = g - 1; return true; }
    if (side == 1 && k >= 1 && k <= 3 && g < G - 1) { g2 = g + 1; return true; }
    return false;
}

// parent of every node + the arrival counters: one token for the node's own walker (a blocked node gets one more,
// which never arrives), one per child
__global__ void k_boundary_build(const int64_t *__restrict__ recs, int G, int64_t nx, int32_t *__restrict__ parent,
                                 int32_t *cnt)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)G * 2 * nx) return;
    const int g = (int)(i / (2 * nx)), side = (int)((i / nx) & 1);
    const int64_t col = i % nx;
    const int64_t *rg = recs + (int64_t)g * 4 * nx;
    const int64_t m = rg[(2 + side) * nx + col];
    int32_t p = -1;
    int g2;
    const int k = rec_k(m);
    if (exit_dir(g, side, k, G, g2)) {
        const int64_t c2 = col + d8_dc(k);
        if (c2 >= 0 && c2 < nx) {
            const int64_t m2 = recs[(int64_t)g2 * 4 * nx + (2 + (1 - side)) * nx + c2];
            const int32_t l2 = rec_link(m2);
            if (rec_k(m2) != 9 && l2 >= 0) p = (int32_t)((int64_t)g2 * 2 * nx + l2);
        }
    }
    parent[i] = p;
    atomicAdd(&cnt[i], rec_unres(m) ? 2 : 1);
    if (p >= 0) atomicAdd(&cnt[p], 1);
}

// every node brings its own weight and token; whoever takes a counter to zero carries the node's total downstream
__global__ void k_boundary_walk(const int64_t *__restrict__ recs, int G, int64_t nx, const int32_t *__restrict__ parent,
                                unsigned long long *acc, int32_t *cnt)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)G * 2 * nx) return;
    const int g = (int)(i / (2 * nx)), side = (int)((i / nx) & 1);
    const int64_t w = recs[(int64_t)g * 4 * nx + side * nx + i % nx];
    int64_t cur = i;
    atomicAdd(&acc[cur], (unsigned long long)w);
    __threadfence();
    while (atomicSub(&cnt[cur], 1) == 1) {                  // the last arrival: cur is final
        __threadfence();
        const unsigned long long total = atomicAdd(&acc[cur], 0ull);
        const int32_t p = parent[cur];
        if (p < 0) break;
        atomicAdd(&acc[p], total);
        __threadfence();
        cur = p;
    }
}

__global__ void k_boundary_pull(const int64_t *__restrict__ recs, int G, int64_t nx, int me, const int64_t *__restrict__ acc,
                                const int32_t *__restrict__ cnt, int64_t *__restrict__ ext_inflow,
                                uint8_t *__restrict__ ext_unres)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 2 * nx) return;
    const int side = (int)(i / nx);
    const int64_t col = i - (int64_t)side * nx;
    int64_t sum = 0;
    uint8_t un = 0;
    const int g2 = side == 0 ? me - 1 : me + 1;
    const bool active = rec_k(recs[(int64_t)me * 4 * nx + (2 + side) * nx + col]) != 9;
    if (active && g2 >= 0 && g2 < G) {
        const int s2 = 1 - side;
        for (int dc = -1; dc <= 1; dc++) {            // donor column = col - dc, its code must move it by dc
            const int64_t c2 = col - dc;
            if (c2 < 0 || c2 >= nx) continue;
            const int k2 = rec_k(recs[(int64_t)g2 * 4 * nx + (2 + s2) * nx + c2]);
            int gg;
            if (!exit_dir(g2, s2, k2, G, gg) || gg != me || d8_dc(k2) != dc) continue;
            const int64_t node = ((int64_t)g2 * 2 + s2) * nx + c2;
            if (cnt[node] == 0) sum += acc[node];     // resolved: every arrival came
            else un = 1;
        }
    }
    ext_inflow[i] = sum;
    ext_unres[i] = un;
}

}  // namespace

extern "C" {

size_t wmf_stripe_boundary_workspace(int G, int64_t nx)
{
    size_t n = (size_t)G * 2 * nx;
    return wmf_align(n * 4) * 2 + wmf_align(n * 8) + 1024;
}

int wmf_stripe_boundary_inflow(const int64_t *recs, int G, int64_t nx, int me, int64_t *ext_inflow, uint8_t *ext_unres,
                               void *workspace, size_t ws_bytes, void *stream)
{
    cudaStream_t st = as_stream(stream);
    WMF_REQUIRE(recs && ext_inflow && ext_unres && workspace, WMF_E_ARG, "null pointer");
    WMF_REQUIRE(G >= 1 && nx > 0 && me >= 0 && me < G, WMF_E_ARG, "bad stripe index");
    const int64_t n = (int64_t)G * 2 * nx;
    WMF_REQUIRE(n < ((int64_t)1 << 31), WMF_E_TOO_LARGE, "boundary forest too large");
    WsCursor cur(workspace, ws_bytes);
    int32_t *parent = cur.take<int32_t>((size_t)n);
    int32_t *cnt = cur.take<int32_t>((size_t)n);             // (cnt and acc are adjacent: one memset clears both)
    int64_t *acc = cur.take<int64_t>((size_t)n);
    WMF_REQUIRE(cur.ok, WMF_E_WORKSPACE, "workspace too small");
    WMF_CUDA_CHECK(cudaMemsetAsync(cnt, 0, (size_t)((char *)(acc + n) - (char *)cnt), st));
    k_boundary_build<<<blocks_for(n, 256), 256, 0, st>>>(recs, G, nx, parent, cnt);
    k_boundary_walk<<<blocks_for(n, 256), 256, 0, st>>>(recs, G, nx, parent, (unsigned long long *)acc, cnt);
    k_boundary_pull<<<blocks_for(2 * nx, 256), 256, 0, st>>>(recs, G, nx, me, acc, cnt, ext_inflow, ext_unres);
    WMF_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
