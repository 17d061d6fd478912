// forest.cu -- accumulation over an explicit forest (parent pointers) with an in-degree driven
// chain walk: every leaf starts a walker that pushes its finished sum into its parent with an L2
// atomic and carries on downstream only if it was the parent's last unfinished child.  Work is
// O(n), there is no level barrier, and integer adds make any schedule bit-exact.
// Used for: the vector-form accumulation of cu.basin_basics / cu.basin_acum
// (cuencas_heavy.f90:593-606), subtree sizes for the BFS ordering of cu.basin_find, and the
// contracted boundary forest of the sweep accumulation.
#include "common.cuh"

namespace {

__global__ void k_forest_count(const int32_t *__restrict__ parent, int32_t *__restrict__ cnt, int64_t n)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int32_t p = parent[i];
    if (p >= 0) atomicAdd(&cnt[p], 1);
}

template <typename T>
struct AtomT;
template <>
struct AtomT<int32_t> {
    static __device__ __forceinline__ int32_t add(int32_t *p, int32_t v) { return atomicAdd(p, v); }
};
template <>
struct AtomT<int64_t> {
    static __device__ __forceinline__ int64_t add(int64_t *p, int64_t v)
    {
        return (int64_t)atomicAdd((unsigned long long *)p, (unsigned long long)v);
    }
};

// cnt holds the number of children after k_forest_count; a node with cnt == 0 is a leaf and is
// never decremented by anybody, so reading cnt[i] == 0 here cannot race with a walker -- but a
// non-leaf can *reach* 0.  Leaves are therefore told apart by a second array-free trick: the
// walk kernel is launched after a marking pass that rewrites leaf counters to -1.
__global__ void k_forest_mark(int32_t *__restrict__ cnt, const uint8_t *__restrict__ blocked, int64_t n)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int32_t c = cnt[i];
    if (blocked && blocked[i]) cnt[i] = c + 1;        // a phantom child that never finishes
    else if (c == 0) cnt[i] = -1;
}

template <typename T, typename W>
__global__ void k_forest_walk(const int32_t *__restrict__ parent, const W *__restrict__ w, T *acc,
                              int32_t *cnt, int64_t n)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (cnt[i] != -1) return;               // only leaves start walkers
    T v = w ? (T)w[i] : (T)1;
    acc[i] = v;                             // nobody else writes a leaf
    int64_t cur = i;
    while (true) {
        int32_t p = parent[cur];
        if (p < 0) break;
        AtomT<T>::add(&acc[p], v);
        __threadfence();
        int32_t old = atomicSub(&cnt[p], 1);
        if (old != 1) break;                // other children still running: they will finish p
        __threadfence();
        T own = w ? (T)w[p] : (T)1;
        v = AtomT<T>::add(&acc[p], own) + own;
        cur = p;
    }
}

template <typename T, typename W>
int forest_accum_t(const int32_t *parent, const W *w, T *acc, int32_t *cnt, const uint8_t *blocked, int64_t n,
                   cudaStream_t st)
{
    if (n <= 0) return 0;
    WMF_CUDA_CHECK(cudaMemsetAsync(cnt, 0, sizeof(int32_t) * n, st));
    WMF_CUDA_CHECK(cudaMemsetAsync(acc, 0, sizeof(T) * n, st));
    const int th = 256;
    k_forest_count<<<blocks_for(n, th), th, 0, st>>>(parent, cnt, n);
    k_forest_mark<<<blocks_for(n, th), th, 0, st>>>(cnt, blocked, n);
    k_forest_walk<T, W><<<blocks_for(n, th), th, 0, st>>>(parent, w, acc, cnt, n);
    WMF_LAUNCH_CHECK();
    return 0;
}

}  // namespace

int forest_accum_i32(const int32_t *parent, const int32_t *w, int32_t *acc, int32_t *cnt, int64_t n, cudaStream_t st,
                     const uint8_t *blocked)
{
    return forest_accum_t<int32_t, int32_t>(parent, w, acc, cnt, blocked, n, st);
}
int forest_accum_i64(const int32_t *parent, const int32_t *w, int64_t *acc, int32_t *cnt, int64_t n, cudaStream_t st,
                     const uint8_t *blocked)
{
    return forest_accum_t<int64_t, int32_t>(parent, w, acc, cnt, blocked, n, st);
}
int forest_accum_i64w(const int32_t *parent, const int64_t *w, int64_t *acc, int32_t *cnt, int64_t n, cudaStream_t st,
                      const uint8_t *blocked)
{
    return forest_accum_t<int64_t, int64_t>(parent, w, acc, cnt, blocked, n, st);
}

// ---- acyclic forests: subtree sums by doubling ---------------------------------------------------
// The chain walk above is O(n) work but its run time is the longest path (the main stem of a basin: thousands of
// dependent L2 round trips).  For a forest that is known to be acyclic -- the BFS structure of a basin -- the sums are
// formed in log2(depth) rounds instead: with ptr_k = the 2^k-th ancestor and V_k[u] = the weight within distance < 2^k
// upstream of u, every node that still has a pointer adds V_k[u] into V_{k+1}[ptr_k[u]] and into its own V_{k+1}.
// The two value buffers swap roles every round; a node whose pointer has run out stops moving its value, so what it
// holds and what still arrives for it end up spread over both buffers and the result is their sum.  One kernel per
// round, no host round trip: a round that finds every pointer exhausted leaves its flag down and the rounds behind it
// return at once.
namespace {
__global__ void k_dbl_init(const int32_t *__restrict__ parent, const int32_t *__restrict__ w, int32_t *__restrict__ a0,
                           int32_t *__restrict__ a1, int32_t *__restrict__ p0, int64_t n)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    a0[i] = w ? w[i] : 1;
    a1[i] = 0;
    p0[i] = parent[i];
}
__global__ void k_dbl_round(const int32_t *__restrict__ pc, int32_t *__restrict__ pn, int32_t *aold, int32_t *anew, int64_t n,
                            int32_t *flags, int r)
{
    if (r > 0 && flags[r - 1] == 0) return;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int32_t p = pc[i];
    if (p < 0) { pn[i] = -1; return; }
    const int32_t a = aold[i];
    aold[i] = 0;                                         // (only this thread reads it; nothing is added to `aold` in this round)
    atomicAdd(&anew[i], a);
    atomicAdd(&anew[p], a);
    pn[i] = pc[p];
    flags[r] = 1;
}
__global__ void k_dbl_final(int32_t *__restrict__ acc, const int32_t *__restrict__ a1, int64_t n)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) acc[i] += a1[i];
}
}  // namespace

// scratch: 3 n int32 + 64.  parent must describe an acyclic forest (-1 = root).
int forest_subtree_sums_i32(const int32_t *parent, const int32_t *w, int32_t *acc, int32_t *scratch, int64_t n, cudaStream_t st)
{
    if (n <= 0) return 0;
    int32_t *a1 = scratch, *p0 = scratch + n, *p1 = scratch + 2 * n, *flags = scratch + 3 * n;
    int rounds = 1;
    while (((int64_t)1 << rounds) < n + 1) rounds++;
    rounds += 1;
    if (rounds > 32) rounds = 32;
    WMF_CUDA_CHECK(cudaMemsetAsync(flags, 0, sizeof(int32_t) * 64, st));
    const int th = 256;
    k_dbl_init<<<blocks_for(n, th), th, 0, st>>>(parent, w, acc, a1, p0, n);
    int32_t *aold = acc, *anew = a1;
    for (int r = 0; r < rounds; r++) {
        k_dbl_round<<<blocks_for(n, th), th, 0, st>>>(p0, p1, aold, anew, n, flags, r);
        int32_t *t = p0; p0 = p1; p1 = t;
        t = aold; aold = anew; anew = t;
    }
    k_dbl_final<<<blocks_for(n, th), th, 0, st>>>(acc, a1, n);
    WMF_LAUNCH_CHECK();
    return 0;
}

namespace {
__global__ void k_resolved(const int32_t *__restrict__ cnt, uint8_t *__restrict__ resolved, int64_t n)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) resolved[i] = cnt[i] <= 0 ? 1 : 0;
}
}  // namespace

extern "C" {

size_t wmf_forest_accum_workspace(int64_t n) { return wmf_align((size_t)n * 4) + 256; }

int wmf_forest_accum(const int32_t *parent, const int64_t *weight, const uint8_t *blocked, int64_t n, int64_t *acc,
                     uint8_t *resolved, void *workspace, size_t ws_bytes, void *stream)
{
    cudaStream_t st = as_stream(stream);
    WMF_REQUIRE(parent && acc && n >= 0, WMF_E_ARG, "bad argument");
    WMF_REQUIRE(n < ((int64_t)1 << 31), WMF_E_TOO_LARGE, "n must be < 2^31");
    if (n == 0) return 0;
    WsCursor cur(workspace, ws_bytes);
    int32_t *cnt = cur.take<int32_t>((size_t)n);
    WMF_REQUIRE(cur.ok, WMF_E_WORKSPACE, "workspace too small");
    int rc = forest_accum_t<int64_t, int64_t>(parent, weight, acc, cnt, blocked, n, st);
    if (rc) return rc;
    if (resolved) {
        k_resolved<<<blocks_for(n, 256), 256, 0, st>>>(cnt, resolved, n);
        WMF_LAUNCH_CHECK();
    }
    return 0;
}

}  // extern "C"
