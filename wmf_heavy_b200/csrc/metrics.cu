// metrics.cu -- "propagate along the D8 links" consumers of the drainage raster:
//   wmf_stream_receiver  replaces wmf_py/cu_py/metrics.py:189-268 hydro_distance_and_receiver
//                        (first stream cell on every cell's downstream path + the path length)
//   wmf_hand             replaces wmf_py/cu_py/metrics.py:271-323 hand (height above that cell)
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

}  // extern "C"
