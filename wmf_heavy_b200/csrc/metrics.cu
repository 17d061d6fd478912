// metrics.cu -- "propagate along the D8 links" consumers of the drainage raster:
//   wmf_stream_receiver  replaces wmf_py/cu_py/metrics.py:189-268 hydro_distance_and_receiver
//                        (first stream cell on every cell's downstream path + the path length)
//   wmf_hand             replaces wmf_py/cu_py/metrics.py:271-323 hand (height above that cell)
//   wmf_path_sum         sum of per-cell weights down every cell's flow path: the distance pass of
//                        basin_ppalstream_find (metrics.py:420-432) and basin_time_to_out (:528-599)
//   wmf_time_to_out      basin_time_to_out itself (Manning travel time per cell, then the path sum)
//   wmf_stream_nodes     node detection and node-to-node links of basin_netxy_find / _cut (streams.py:282-384)
// Both are pointer jumping over the receiver graph with the stream cells as roots (log2(path)
// rounds, no frontier, no atomics); the reference walks a BFS / topological queue.
#include "common.cuh"

namespace {

template <int ENC>
__global__ void k_sr_init(const int8_t *__restrict__ dir, const uint8_t *__restrict__ mask,
                          const uint8_t *__restrict__ stream, const uint8_t *__restrict__ extra_root, int64_t ny,
                          int64_t nx, int32_t *__restrict__ ptr, uint32_t *__restrict__ n_dx, uint32_t *__restrict__ n_dy,
                          uint32_t *__restrict__ n_dg)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ny * nx) return;
    int64_t p = i;
    uint32_t ax = 0, ay = 0, ag = 0;
    const bool root = (stream && stream[i]) || (extra_root && extra_root[i]);
    if (cell_active(mask, i) && !root) {
        int64_t r = i / nx, c = i - r * nx;
        int64_t j = raster_recv<ENC>(dir, mask, ny, nx, r, c);
        if (j >= 0) {
            p = j;
            int k = d8_norm<ENC>(dir[i]);
            int dr = d8_dr(k), dc = d8_dc(k);
            if (dr != 0 && dc != 0) ag = 1;           // metrics.py:249-254: diagonal, else dx when the row stays, else dy
            else if (dr == 0) ax = 1;
            else ay = 1;
        }
    }
    ptr[i] = (int32_t)p;
    if (n_dx) { n_dx[i] = ax; n_dy[i] = ay; n_dg[i] = ag; }
}

__global__ void k_sr_jump(const int32_t *__restrict__ pa, const uint32_t *__restrict__ xa, const uint32_t *__restrict__ ya,
                          const uint32_t *__restrict__ ga, int32_t *__restrict__ pb, uint32_t *__restrict__ xb,
                          uint32_t *__restrict__ yb, uint32_t *__restrict__ gb, int64_t n, int *changed)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int32_t p = pa[i], pp = pa[p];
    pb[i] = pp;
    if (xa) { xb[i] = xa[i] + xa[p]; yb[i] = ya[i] + ya[p]; gb[i] = ga[i] + ga[p]; }
    if (pp != p) *changed = 1;
}

__global__ void k_sr_finish(const int32_t *__restrict__ ptr, const uint32_t *__restrict__ nx_, const uint32_t *__restrict__ ny_,
                            const uint32_t *__restrict__ ng_, const uint8_t *__restrict__ mask,
                            const uint8_t *__restrict__ stream, int64_t n, double dx, double dy, double diag,
                            int32_t *__restrict__ a_quien, double *__restrict__ hdnd)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int32_t aq = -1;
    double h = __longlong_as_double(0x7ff8000000000000ll);      // nan
    if (cell_active(mask, i)) {
        const int32_t r = ptr[i];
        if (ptr[r] == r && stream[r] && cell_active(mask, r)) {
            aq = r;
            h = (double)nx_[i] * dx + (double)ny_[i] * dy + (double)ng_[i] * diag;
        }
    }
    a_quien[i] = aq;
    if (hdnd) hdnd[i] = h;
}

// After R >= log2(n) doublings ptr[i] is the 2^R-th successor of i: a true root (a cell without receiver)
// unless i drains into a cycle, in which case it is a cell ON that cycle -- and every cycle cell is such an
// image.  (ptr[p] == p proves nothing: on a cycle whose length divides 2^R every cell maps to itself.)
template <int ENC>
__global__ void k_mark_cycles(const int8_t *__restrict__ dir, const uint8_t *__restrict__ mask, int64_t ny, int64_t nx,
                              const int32_t *__restrict__ ptr, uint8_t *__restrict__ is_cycle)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ny * nx) return;
    const int64_t p = ptr[i];
    if (!cell_active(mask, p)) return;
    const int64_t r = p / nx, c = p - r * nx;
    if (raster_recv<ENC>(dir, mask, ny, nx, r, c) >= 0) is_cycle[p] = 1;
}

__global__ void k_hand_finish(const int32_t *__restrict__ ptr, const double *__restrict__ dem, const uint8_t *__restrict__ mask,
                              const uint8_t *__restrict__ stream, int64_t n, double *__restrict__ out)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double nan = __longlong_as_double(0x7ff8000000000000ll);
    double v = nan;
    if (cell_active(mask, i)) {
        const int32_t r = ptr[i];
        if (stream[r] && cell_active(mask, r)) {     // metrics.py:282-283, 313-318: base = dem of the first stream cell
            v = dem[i] - dem[r];
            if (v < 0.0) v = 0.0;                    // np.maximum(out, 0): nan stays nan
        }
    }
    out[i] = v;
}


// ---- path sums -------------------------------------------------------------------------------------
// ptr / val for pointer jumping: val[i] = sum of the weights of i .. (the cell before ptr[i]).  Ends (no receiver, or a
// cell on a drainage cycle) point to themselves and hold the end value: 0, or nan when cycles poison their upstream.
template <int ENC>
__global__ void k_ps_init(const int8_t *__restrict__ dir, const uint8_t *__restrict__ active, const uint8_t *__restrict__ is_cycle,
                          const double *__restrict__ w, int64_t ny, int64_t nx, double dx, double dy, double diag, int cycle_nan,
                          int32_t *__restrict__ ptr, double *__restrict__ val)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ny * nx) return;
    const double nan = __longlong_as_double(0x7ff8000000000000ll);
    int64_t p = i;
    double v = 0.0;
    if (!cell_active(active, i)) v = nan;
    else if (is_cycle[i]) v = cycle_nan ? nan : 0.0;
    else {
        int64_t r = i / nx, c = i - r * nx;
        int64_t j = raster_recv<ENC>(dir, active, ny, nx, r, c);
        if (j >= 0) {
            p = j;
            if (w) v = w[i];
            else {
                int k = d8_norm<ENC>(dir[i]);
                int dr = d8_dr(k), dc = d8_dc(k);
                v = (dr != 0 && dc != 0) ? diag : (dc != 0 ? dx : dy);   // metrics.py:422-425
            }
        }
    }
    ptr[i] = (int32_t)p;
    val[i] = v;
}

__global__ void k_ps_jump(const int32_t *__restrict__ pa, const double *__restrict__ va, int32_t *__restrict__ pb,
                          double *__restrict__ vb, int64_t n, int *changed)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int32_t p = pa[i], pp = pa[p];
    pb[i] = pp;
    vb[i] = p != i ? va[i] + va[p] : va[i];          // an end keeps its value
    if (pp != p) *changed = 1;
}

// Manning travel time of a cell towards its receiver, unit hydraulic radius (metrics.py:568-581); nan when the speed is 0
template <int ENC>
__global__ void k_time_weights(const int8_t *__restrict__ dir, const double *__restrict__ slope, const double *__restrict__ manning,
                               int64_t n, double dx, double dy, double diag, double *__restrict__ w)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double nan = __longlong_as_double(0x7ff8000000000000ll);
    const int k = d8_norm<ENC>(dir[i]);
    double t = nan;
    if (k < 8) {
        const int dr = d8_dr(k), dc = d8_dc(k);
        const double len = (dr != 0 && dc != 0) ? diag : (dc != 0 ? dx : dy);
        const double v = (1.0 / fmax(manning[i], 1e-6)) * sqrt(fmax(slope[i], 0.0));
        if (v > 0.0) t = len / v;
    }
    w[i] = t;
}

// ---- stream network nodes (streams.py:282-333) -------------------------------------------------------
// degree = stream donors + (1 if the receiver is a stream cell); a stream cell is a node when degree != 2 or it is the
// outlet.  donor[i] = the unique stream donor of a non-node cell (its upstream neighbour on the edge), else -1.
template <int ENC>
__global__ void k_stream_nodes(const int8_t *__restrict__ dir, const uint8_t *__restrict__ stream, int64_t ny, int64_t nx,
                               int64_t outlet, uint8_t *__restrict__ node, int32_t *__restrict__ recv, int32_t *__restrict__ donor)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ny * nx) return;
    uint8_t is_node = 0;
    int32_t rv = -1, dn = -1;
    if (stream[i]) {
        const int64_t r = i / nx, c = i - r * nx;
        int indeg = 0;
        for (int k = 0; k < 8; k++) {                 // neighbour (r + dr, c + dc) drains here if its code is the opposite move
            const int64_t rr = r + d8_dr(k), cc = c + d8_dc(k);
            if (rr < 0 || rr >= ny || cc < 0 || cc >= nx) continue;
            const int64_t j = rr * nx + cc;
            if (!stream[j]) continue;
            const int kj = d8_norm<ENC>(dir[j]);
            if (kj < 8 && rr + d8_dr(kj) == r && cc + d8_dc(kj) == c) { indeg++; dn = (int32_t)j; }
        }
        const int k = d8_norm<ENC>(dir[i]);
        int outdeg = 0;
        if (k < 8) {
            const int64_t rr = r + d8_dr(k), cc = c + d8_dc(k);
            if (rr >= 0 && rr < ny && cc >= 0 && cc < nx && stream[rr * nx + cc]) { outdeg = 1; rv = (int32_t)(rr * nx + cc); }
        }
        is_node = (i == outlet || indeg + outdeg != 2) ? 1 : 0;
        if (indeg != 1) dn = -1;
    }
    node[i] = is_node;
    recv[i] = rv;
    donor[i] = dn;
}

// pointer jumping towards the next node: g(i) = i for nodes / cells without the link, else link[i]; after the jumps
// tgt[i] = the node reached from link[i], cnt[i] = cells from i to it (i excluded, the node included)
__global__ void k_net_init(const uint8_t *__restrict__ stream, const uint8_t *__restrict__ node, const int32_t *__restrict__ link,
                           int64_t n, int32_t *__restrict__ ptr, int32_t *__restrict__ cnt)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int32_t p = (int32_t)i, c = 0;
    if (stream[i] && !node[i] && link[i] >= 0) { p = link[i]; c = 1; }
    ptr[i] = p;
    cnt[i] = c;
}
__global__ void k_net_jump(const int32_t *__restrict__ pa, const int32_t *__restrict__ ca, int32_t *__restrict__ pb,
                           int32_t *__restrict__ cb, int64_t n, int *changed)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int32_t p = pa[i], pp = pa[p];
    pb[i] = pp;
    cb[i] = ca[i] + (p != i ? ca[p] : 0);
    if (pp != p) *changed = 1;
}
__global__ void k_net_finish(const uint8_t *__restrict__ stream, const uint8_t *__restrict__ node, const int32_t *__restrict__ link,
                             const int32_t *__restrict__ ptr, const int32_t *__restrict__ cnt, int64_t n,
                             int32_t *__restrict__ tgt, int32_t *__restrict__ steps)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int32_t t = -1, s = 0;
    if (stream[i] && link[i] >= 0) {
        const int32_t first = link[i], end = ptr[first];
        if (node[end]) { t = end; s = 1 + cnt[first]; }    // (a node-free loop or a dead end: no node is reached)
    }
    tgt[i] = t;
    steps[i] = s;
}

// jump until nothing moves (or, with cycles present, for ~log2(n) rounds); the result is in *pp_out
template <bool COUNTS>
int run_jumps(int64_t n, int32_t *pA, int32_t *pB, uint32_t *xA, uint32_t *yA, uint32_t *gA, uint32_t *xB, uint32_t *yB,
              uint32_t *gB, int *d_changed, bool until_fixed, int32_t **p_out, uint32_t **x_out, uint32_t **y_out,
              uint32_t **g_out, cudaStream_t st)
{
    int max_rounds = 2;
    while (((int64_t)1 << max_rounds) < n) max_rounds++;
    max_rounds += 2;
    for (int r = 0; r < max_rounds;) {
        WMF_CUDA_CHECK(cudaMemsetAsync(d_changed, 0, sizeof(int), st));
        for (int k = 0; k < 4 && r < max_rounds; k++, r++) {
            k_sr_jump<<<blocks_for(n, 256), 256, 0, st>>>(pA, COUNTS ? xA : nullptr, yA, gA, pB, xB, yB, gB, n, d_changed);
            int32_t *tp = pA; pA = pB; pB = tp;
            uint32_t *t;
            t = xA; xA = xB; xB = t;
            t = yA; yA = yB; yB = t;
            t = gA; gA = gB; gB = t;
        }
        if (until_fixed) {
            int h = 0;
            WMF_CUDA_CHECK(cudaMemcpyAsync(&h, d_changed, sizeof(int), cudaMemcpyDeviceToHost, st));
            WMF_CUDA_CHECK(cudaStreamSynchronize(st));
            if (!h) break;
        }
    }
    WMF_LAUNCH_CHECK();
    *p_out = pA; *x_out = xA; *y_out = yA; *g_out = gA;
    return 0;
}


// cells on drainage cycles of the graph restricted to `active` (they never enter the reference's topological orders)
static int mark_cycles(const int8_t *dir, const uint8_t *active, int64_t ny, int64_t nx, int encoding, int32_t *pA, int32_t *pB,
                       uint8_t *is_cycle, int *d_changed, cudaStream_t st)
{
    const int64_t n = ny * nx;
    int32_t *p;
    uint32_t *x, *y, *g;
    if (encoding == WMF_ENC_INTERNAL) k_sr_init<WMF_ENC_INTERNAL><<<blocks_for(n, 256), 256, 0, st>>>(dir, active, nullptr, nullptr, ny, nx, pA, nullptr, nullptr, nullptr);
    else k_sr_init<WMF_ENC_KEYPAD><<<blocks_for(n, 256), 256, 0, st>>>(dir, active, nullptr, nullptr, ny, nx, pA, nullptr, nullptr, nullptr);
    int rc = run_jumps<false>(n, pA, pB, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, d_changed, false, &p, &x, &y, &g, st);
    if (rc) return rc;
    WMF_CUDA_CHECK(cudaMemsetAsync(is_cycle, 0, (size_t)n, st));
    if (encoding == WMF_ENC_INTERNAL) k_mark_cycles<WMF_ENC_INTERNAL><<<blocks_for(n, 256), 256, 0, st>>>(dir, active, ny, nx, p, is_cycle);
    else k_mark_cycles<WMF_ENC_KEYPAD><<<blocks_for(n, 256), 256, 0, st>>>(dir, active, ny, nx, p, is_cycle);
    WMF_LAUNCH_CHECK();
    return 0;
}

static int path_sum(const int8_t *dir, const uint8_t *active, const double *w, int64_t ny, int64_t nx, int encoding, double dx,
                    double dy, int cycle_nan, double *out, void *workspace, size_t ws_bytes, cudaStream_t st)
{
    const int64_t n = ny * nx;
    WsCursor cur(workspace, ws_bytes);
    int32_t *pA = cur.take<int32_t>((size_t)n), *pB = cur.take<int32_t>((size_t)n);
    double *vB = cur.take<double>((size_t)n);
    uint8_t *is_cycle = cur.take<uint8_t>((size_t)n);
    int *d_changed = cur.take<int>(1);
    WMF_REQUIRE(cur.ok, WMF_E_WORKSPACE, "workspace too small");
    int rc = mark_cycles(dir, active, ny, nx, encoding, pA, pB, is_cycle, d_changed, st);
    if (rc) return rc;
    double *vA = out;                                        // ping-pong between out and the workspace
    if (encoding == WMF_ENC_INTERNAL) k_ps_init<WMF_ENC_INTERNAL><<<blocks_for(n, 256), 256, 0, st>>>(dir, active, is_cycle, w, ny, nx, dx, dy, hypot(dx, dy), cycle_nan, pA, vA);
    else k_ps_init<WMF_ENC_KEYPAD><<<blocks_for(n, 256), 256, 0, st>>>(dir, active, is_cycle, w, ny, nx, dx, dy, hypot(dx, dy), cycle_nan, pA, vA);
    int max_rounds = 2;
    while (((int64_t)1 << max_rounds) < n) max_rounds++;
    max_rounds += 2;
    for (int r = 0; r < max_rounds;) {
        WMF_CUDA_CHECK(cudaMemsetAsync(d_changed, 0, sizeof(int), st));
        for (int k = 0; k < 4 && r < max_rounds; k++, r++) {
            k_ps_jump<<<blocks_for(n, 256), 256, 0, st>>>(pA, vA, pB, vB, n, d_changed);
            int32_t *tp = pA; pA = pB; pB = tp;
            double *tv = vA; vA = vB; vB = tv;
        }
        int h = 0;
        WMF_CUDA_CHECK(cudaMemcpyAsync(&h, d_changed, sizeof(int), cudaMemcpyDeviceToHost, st));
        WMF_CUDA_CHECK(cudaStreamSynchronize(st));
        if (!h) break;
    }
    if (vA != out) WMF_CUDA_CHECK(cudaMemcpyAsync(out, vA, sizeof(double) * (size_t)n, cudaMemcpyDeviceToDevice, st));
    WMF_LAUNCH_CHECK();
    return 0;
}

}  // namespace

extern "C" {

size_t wmf_stream_receiver_workspace(int64_t ny, int64_t nx) { return 8 * wmf_align((size_t)(ny * nx) * 4) + 1024; }

int wmf_stream_receiver(const int8_t *dir, const uint8_t *mask, const uint8_t *stream, int64_t ny, int64_t nx, int encoding,
                        double dx, double dy, int32_t *a_quien, double *hdnd, void *workspace, size_t ws_bytes, void *stream_)
{
    cudaStream_t st = as_stream(stream_);
    WMF_REQUIRE(dir && stream && a_quien && ny > 0 && nx > 0, WMF_E_ARG, "bad argument");
    WMF_REQUIRE(ny * nx < ((int64_t)1 << 31), WMF_E_TOO_LARGE, "raster needs ny*nx < 2^31");
    WMF_REQUIRE(encoding == WMF_ENC_INTERNAL || encoding == WMF_ENC_KEYPAD, WMF_E_ARG, "unknown encoding");
    const int64_t n = ny * nx;
    WsCursor cur(workspace, ws_bytes);
    int32_t *pA = cur.take<int32_t>((size_t)n), *pB = cur.take<int32_t>((size_t)n);
    uint32_t *xA = cur.take<uint32_t>((size_t)n), *yA = cur.take<uint32_t>((size_t)n), *gA = cur.take<uint32_t>((size_t)n);
    uint32_t *xB = cur.take<uint32_t>((size_t)n), *yB = cur.take<uint32_t>((size_t)n), *gB = cur.take<uint32_t>((size_t)n);
    int *d_changed = cur.take<int>(1);
    WMF_REQUIRE(cur.ok, WMF_E_WORKSPACE, "workspace too small");
    if (encoding == WMF_ENC_INTERNAL) k_sr_init<WMF_ENC_INTERNAL><<<blocks_for(n, 256), 256, 0, st>>>(dir, mask, stream, nullptr, ny, nx, pA, xA, yA, gA);
    else k_sr_init<WMF_ENC_KEYPAD><<<blocks_for(n, 256), 256, 0, st>>>(dir, mask, stream, nullptr, ny, nx, pA, xA, yA, gA);
    int32_t *p;
    uint32_t *x, *y, *g;
    int rc = run_jumps<true>(n, pA, pB, xA, yA, gA, xB, yB, gB, d_changed, true, &p, &x, &y, &g, st);
    if (rc) return rc;
    k_sr_finish<<<blocks_for(n, 256), 256, 0, st>>>(p, x, y, g, mask, stream, n, dx, dy, hypot(dx, dy), a_quien, hdnd);
    WMF_LAUNCH_CHECK();
    return 0;
}

size_t wmf_hand_workspace(int64_t ny, int64_t nx) { return 2 * wmf_align((size_t)(ny * nx) * 4) + wmf_align((size_t)(ny * nx)) + 1024; }

int wmf_hand(const double *dem, const int8_t *dir, const uint8_t *mask, const uint8_t *stream, int64_t ny, int64_t nx,
             int encoding, double *hand, void *workspace, size_t ws_bytes, void *stream_)
{
    cudaStream_t st = as_stream(stream_);
    WMF_REQUIRE(dem && dir && stream && hand && ny > 0 && nx > 0, WMF_E_ARG, "bad argument");
    WMF_REQUIRE(ny * nx < ((int64_t)1 << 31), WMF_E_TOO_LARGE, "raster needs ny*nx < 2^31");
    WMF_REQUIRE(encoding == WMF_ENC_INTERNAL || encoding == WMF_ENC_KEYPAD, WMF_E_ARG, "unknown encoding");
    const int64_t n = ny * nx;
    WsCursor cur(workspace, ws_bytes);
    int32_t *pA = cur.take<int32_t>((size_t)n), *pB = cur.take<int32_t>((size_t)n);
    uint8_t *is_cycle = cur.take<uint8_t>((size_t)n);
    int *d_changed = cur.take<int>(1);
    WMF_REQUIRE(cur.ok, WMF_E_WORKSPACE, "workspace too small");
    int32_t *p;
    uint32_t *x, *y, *g;
    // 1. cells on drainage cycles never enter the reference's topological order (metrics.py:296-310): find them
    if (encoding == WMF_ENC_INTERNAL) k_sr_init<WMF_ENC_INTERNAL><<<blocks_for(n, 256), 256, 0, st>>>(dir, mask, nullptr, nullptr, ny, nx, pA, nullptr, nullptr, nullptr);
    else k_sr_init<WMF_ENC_KEYPAD><<<blocks_for(n, 256), 256, 0, st>>>(dir, mask, nullptr, nullptr, ny, nx, pA, nullptr, nullptr, nullptr);
    int rc = run_jumps<false>(n, pA, pB, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, d_changed, false, &p, &x, &y, &g, st);
    if (rc) return rc;
    WMF_CUDA_CHECK(cudaMemsetAsync(is_cycle, 0, (size_t)n, st));
    if (encoding == WMF_ENC_INTERNAL) k_mark_cycles<WMF_ENC_INTERNAL><<<blocks_for(n, 256), 256, 0, st>>>(dir, mask, ny, nx, p, is_cycle);
    else k_mark_cycles<WMF_ENC_KEYPAD><<<blocks_for(n, 256), 256, 0, st>>>(dir, mask, ny, nx, p, is_cycle);
    // 2. first stream cell downstream, with stream cells AND cycle cells as roots
    if (encoding == WMF_ENC_INTERNAL) k_sr_init<WMF_ENC_INTERNAL><<<blocks_for(n, 256), 256, 0, st>>>(dir, mask, stream, is_cycle, ny, nx, pA, nullptr, nullptr, nullptr);
    else k_sr_init<WMF_ENC_KEYPAD><<<blocks_for(n, 256), 256, 0, st>>>(dir, mask, stream, is_cycle, ny, nx, pA, nullptr, nullptr, nullptr);
    rc = run_jumps<false>(n, pA, pB, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, d_changed, true, &p, &x, &y, &g, st);
    if (rc) return rc;
    k_hand_finish<<<blocks_for(n, 256), 256, 0, st>>>(p, dem, mask, stream, n, hand);
    WMF_LAUNCH_CHECK();
    return 0;
}


size_t wmf_path_sum_workspace(int64_t ny, int64_t nx)
{
    const size_t n = (size_t)(ny * nx);
    return 2 * wmf_align(n * 4) + 2 * wmf_align(n * 8) + wmf_align(n) + 1024;
}

int wmf_path_sum(const int8_t *dir, const uint8_t *active, const double *w, int64_t ny, int64_t nx, int encoding, double dx,
                 double dy, int cycle_nan, double *out, void *workspace, size_t ws_bytes, void *stream_)
{
    WMF_REQUIRE(dir && out && workspace && ny > 0 && nx > 0, WMF_E_ARG, "bad argument");
    WMF_REQUIRE(ny * nx < ((int64_t)1 << 31), WMF_E_TOO_LARGE, "raster needs ny*nx < 2^31");
    WMF_REQUIRE(encoding == WMF_ENC_INTERNAL || encoding == WMF_ENC_KEYPAD, WMF_E_ARG, "unknown encoding");
    return path_sum(dir, active, w, ny, nx, encoding, dx, dy, cycle_nan, out, workspace, ws_bytes, as_stream(stream_));
}

size_t wmf_time_to_out_workspace(int64_t ny, int64_t nx) { return wmf_path_sum_workspace(ny, nx) + wmf_align((size_t)(ny * nx) * 8); }

int wmf_time_to_out(const int8_t *dir, const double *slope, const double *manning, int64_t ny, int64_t nx, int encoding,
                    double dx, double dy, double *out, void *workspace, size_t ws_bytes, void *stream_)
{
    cudaStream_t st = as_stream(stream_);
    WMF_REQUIRE(dir && slope && manning && out && workspace && ny > 0 && nx > 0, WMF_E_ARG, "bad argument");
    WMF_REQUIRE(ny * nx < ((int64_t)1 << 31), WMF_E_TOO_LARGE, "raster needs ny*nx < 2^31");
    WMF_REQUIRE(encoding == WMF_ENC_INTERNAL || encoding == WMF_ENC_KEYPAD, WMF_E_ARG, "unknown encoding");
    const int64_t n = ny * nx;
    const size_t ps = wmf_path_sum_workspace(ny, nx);
    WMF_REQUIRE(ws_bytes >= wmf_time_to_out_workspace(ny, nx), WMF_E_WORKSPACE, "workspace too small");
    double *w = reinterpret_cast<double *>((char *)workspace + ps);      // per-cell travel times, behind the path-sum scratch
    if (encoding == WMF_ENC_INTERNAL) k_time_weights<WMF_ENC_INTERNAL><<<blocks_for(n, 256), 256, 0, st>>>(dir, slope, manning, n, dx, dy, hypot(dx, dy), w);
    else k_time_weights<WMF_ENC_KEYPAD><<<blocks_for(n, 256), 256, 0, st>>>(dir, slope, manning, n, dx, dy, hypot(dx, dy), w);
    WMF_LAUNCH_CHECK();
    return path_sum(dir, nullptr, w, ny, nx, encoding, dx, dy, 1, out, workspace, ps, st);
}

size_t wmf_stream_nodes_workspace(int64_t ny, int64_t nx) { return 6 * wmf_align((size_t)(ny * nx) * 4) + 1024; }

int wmf_stream_nodes(const int8_t *dir, const uint8_t *stream, int64_t ny, int64_t nx, int encoding, int64_t outlet,
                     uint8_t *node, int32_t *down_node, int32_t *down_steps, int32_t *up_node, int32_t *up_steps,
                     void *workspace, size_t ws_bytes, void *stream_)
{
    cudaStream_t st = as_stream(stream_);
    WMF_REQUIRE(dir && stream && node && down_node && down_steps && up_node && up_steps && workspace && ny > 0 && nx > 0, WMF_E_ARG, "bad argument");
    WMF_REQUIRE(ny * nx < ((int64_t)1 << 31), WMF_E_TOO_LARGE, "raster needs ny*nx < 2^31");
    WMF_REQUIRE(encoding == WMF_ENC_INTERNAL || encoding == WMF_ENC_KEYPAD, WMF_E_ARG, "unknown encoding");
    const int64_t n = ny * nx;
    WsCursor cur(workspace, ws_bytes);
    int32_t *recv = cur.take<int32_t>((size_t)n), *donor = cur.take<int32_t>((size_t)n);
    int32_t *pA = cur.take<int32_t>((size_t)n), *pB = cur.take<int32_t>((size_t)n);
    int32_t *cA = cur.take<int32_t>((size_t)n), *cB = cur.take<int32_t>((size_t)n);
    int *d_changed = cur.take<int>(1);
    WMF_REQUIRE(cur.ok, WMF_E_WORKSPACE, "workspace too small");
    if (encoding == WMF_ENC_INTERNAL) k_stream_nodes<WMF_ENC_INTERNAL><<<blocks_for(n, 256), 256, 0, st>>>(dir, stream, ny, nx, outlet, node, recv, donor);
    else k_stream_nodes<WMF_ENC_KEYPAD><<<blocks_for(n, 256), 256, 0, st>>>(dir, stream, ny, nx, outlet, node, recv, donor);
    int max_rounds = 2;
    while (((int64_t)1 << max_rounds) < n) max_rounds++;
    max_rounds += 2;
    for (int pass = 0; pass < 2; pass++) {                   // downstream (receivers), then upstream (the unique donors)
        const int32_t *link = pass == 0 ? recv : donor;
        int32_t *a = pA, *ca = cA, *b = pB, *cb = cB;
        k_net_init<<<blocks_for(n, 256), 256, 0, st>>>(stream, node, link, n, a, ca);
        for (int r = 0; r < max_rounds;) {
            WMF_CUDA_CHECK(cudaMemsetAsync(d_changed, 0, sizeof(int), st));
            for (int k = 0; k < 4 && r < max_rounds; k++, r++) {
                k_net_jump<<<blocks_for(n, 256), 256, 0, st>>>(a, ca, b, cb, n, d_changed);
                int32_t *t = a; a = b; b = t;
                t = ca; ca = cb; cb = t;
            }
            int h = 0;
            WMF_CUDA_CHECK(cudaMemcpyAsync(&h, d_changed, sizeof(int), cudaMemcpyDeviceToHost, st));
            WMF_CUDA_CHECK(cudaStreamSynchronize(st));
            if (!h) break;
        }
        k_net_finish<<<blocks_for(n, 256), 256, 0, st>>>(stream, node, link, a, ca, n, pass == 0 ? down_node : up_node,
                                                         pass == 0 ? down_steps : up_steps);
    }
    WMF_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
