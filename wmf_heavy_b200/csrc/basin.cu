// basin.cu -- basin delineation (raster mask and BFS-ordered vector `structure`), vector-form
// accumulation, cell basics and stream/node marks.
//
// Reference routines replaced (uranzea/WMF_heavy):
//   wmf_py/cu_py/basics.py:206-248      basin_cut            -> wmf_basin_mask
//   wmf_heavy/cuencas_heavy.f90:507-576 basin_find/basin_cut -> wmf_basin_structure
//   wmf_heavy/cuencas_heavy.f90:581-627 basin_basics         -> wmf_basin_basics / wmf_basin_acum_vec
//   wmf_heavy/cuencas_heavy.f90:639-646 basin_coordXY        -> wmf_basin_coordxy
//   wmf_heavy/cuencas_heavy.f90:780-809 basin_stream_nod     -> wmf_stream_nod
//
// The reference delineates with a sequential BFS.  Here membership is "my D8 path reaches the
// outlet", resolved by pointer jumping over the raster (log rounds, no frontier), and the BFS
// *order* the Fortran produces is reconstructed without running a BFS: within a BFS level the
// discovery order equals DFS preorder restricted to that level when children are ordered by the
// Fortran scan slot (kf outer, kc inner, f90:533-545).  So: subtree sizes (forest engine) ->
// preorder offsets among siblings -> (depth, preorder) by pointer jumping -> one radix sort.
#include <cub/cub.cuh>

#include "common.cuh"

size_t d8_basin_membership_workspace(int64_t ny, int64_t nx);
int d8_basin_membership(const int8_t *dir, const uint8_t *mask, int64_t ny, int64_t nx, int encoding, int64_t outlet, uint8_t *bm,
                        int32_t *flag32, unsigned long long *count, void *ws, size_t ws_bytes, cudaStream_t st);

namespace {

// ------------------------------------------------------------------ membership
// "My D8 path reaches the outlet": the tile contraction of d8_sweep.cu (labels per tile in shared memory, pointer
// jumping over the tiles' exit cells only) -- see d8_basin_membership there.

// ------------------------------------------------------------------ structure building (keypad)
__global__ void k_scatter_cells(const int32_t *__restrict__ flag32, const int32_t *__restrict__ cidx, int64_t n,
                                int32_t *__restrict__ cell)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && flag32[i]) cell[cidx[i]] = (int32_t)i;
}

__global__ void k_parent_compact(const int8_t *__restrict__ dir, const int32_t *__restrict__ cell,
                                 const int32_t *__restrict__ cidx, int64_t nb, int64_t ny, int64_t nx, int32_t outlet,
                                 int32_t *__restrict__ parentC)
{
    int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nb) return;
    int32_t i = cell[b];
    int32_t p = -1;
    if (i != outlet) {
        int64_t r = i / nx, c = i - r * nx;
        int64_t j = raster_recv<WMF_ENC_KEYPAD>(dir, nullptr, ny, nx, r, c);
        p = cidx[j];                       // a basin cell other than the outlet always has a basin receiver
    }
    parentC[b] = p;
}

// preorder offset of b relative to its parent: 1 + sizes of the siblings scanned before it.
// Scan order of the Fortran BFS (f90:533-545): kf = 1..3 (fil-1, fil, fil+1) outer, kc = 1..3 inner;
// the donor test is DIR(c, f) == 3*kf - kc + 1.
__global__ void k_sibling_offset(const int8_t *__restrict__ dir, const int32_t *__restrict__ cell,
                                 const int32_t *__restrict__ cidx, const int32_t *__restrict__ parentC,
                                 const int32_t *__restrict__ size, int64_t nb, int64_t ny, int64_t nx,
                                 int32_t *__restrict__ off)
{
    int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nb) return;
    int32_t pc = parentC[b];
    if (pc < 0) { off[b] = 0; return; }
    int32_t me = cell[b];
    int32_t P = cell[pc];
    int64_t pr = P / nx, pcx = P - pr * nx;
    int32_t o = 1;
    bool done = false;
    for (int kf = 1; kf <= 3 && !done; kf++)
        for (int kc = 1; kc <= 3; kc++) {
            if (kf == 2 && kc == 2) continue;
            int64_t f = pr + kf - 2, c = pcx + kc - 2;
            if (f < 0 || f >= ny || c < 0 || c >= nx) continue;
            int64_t j = f * nx + c;
            if ((int)dir[j] != 3 * kf - kc + 1) continue;
            if (j == me) { done = true; break; }
            o += size[cidx[j]];
        }
    off[b] = o;
}

__global__ void k_jump_init(const int32_t *__restrict__ parentC, const int32_t *__restrict__ off, int64_t nb,
                            int32_t *ptr, int32_t *pre, int32_t *dep)
{
    int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nb) return;
    ptr[b] = parentC[b];
    pre[b] = off[b];
    dep[b] = parentC[b] >= 0 ? 1 : 0;
}

__global__ void k_jump_round(const int32_t *__restrict__ ptr, const int32_t *__restrict__ pre,
                             const int32_t *__restrict__ dep, int32_t *__restrict__ ptr2, int32_t *__restrict__ pre2,
                             int32_t *__restrict__ dep2, int64_t nb)
{
    int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nb) return;
    int32_t p = ptr[b];
    if (p >= 0) {
        pre2[b] = pre[b] + pre[p];
        dep2[b] = dep[b] + dep[p];
        ptr2[b] = ptr[p];
    } else {
        pre2[b] = pre[b];
        dep2[b] = dep[b];
        ptr2[b] = -1;
    }
}

__global__ void k_make_keys(const int32_t *__restrict__ pre, const int32_t *__restrict__ dep, int64_t nb,
                            unsigned long long *__restrict__ keys, int32_t *__restrict__ vals)
{
    int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nb) return;
    keys[b] = ((unsigned long long)(uint32_t)dep[b] << 32) | (uint32_t)pre[b];
    vals[b] = (int32_t)b;
}

__global__ void k_positions(const int32_t *__restrict__ sorted, int64_t nb, int32_t *__restrict__ pos)
{
    int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p < nb) pos[sorted[p]] = (int32_t)p;
}

__global__ void k_write_structure(const int32_t *__restrict__ sorted, const int32_t *__restrict__ pos,
                                  const int32_t *__restrict__ parentC, const int32_t *__restrict__ cell, int64_t nb,
                                  int64_t nx, int32_t *__restrict__ st, int64_t stride)
{
    int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= nb) return;
    int32_t b = sorted[p];
    int64_t i0 = nb - 1 - p;               // Fortran index n - p (1-based)
    int32_t pc = parentC[b];
    int32_t i = cell[b];
    st[0 * stride + i0] = pc >= 0 ? pos[pc] + 1 : 0;    // pop number of the receiver
    st[1 * stride + i0] = (int32_t)(i % nx) + 1;         // col
    st[2 * stride + i0] = (int32_t)(i / nx) + 1;         // fil
}

// ------------------------------------------------------------------ vector-form kernels
__global__ void k_vec_parent(const int32_t *__restrict__ st, int64_t n, int32_t *__restrict__ parent)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int32_t dr = st[i];
    parent[i] = (dr != 0 && dr <= n && dr > 0) ? (int32_t)(n - dr) : -1;
}

__global__ void k_basics(const int32_t *__restrict__ st, int64_t n, const float *__restrict__ dem,
                         const int32_t *__restrict__ dirv, float dxp, float *__restrict__ lon, float *__restrict__ pend,
                         float *__restrict__ elev)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float z = dem[i];
    elev[i] = z;
    float L = (dirv[i] % 2 == 0) ? dxp : __fmul_rn(dxp, sqrtf(2.0f));      // f90:597-602
    lon[i] = L;
    int32_t dr = st[i];
    float s = 0.0f;
    if (dr != 0) s = __fdiv_rn(fabsf(__fsub_rn(z, dem[n - dr])), L);         // f90:605-606
    if (s == 0.0f) s = 0.001f;                                                // f90:610
    pend[i] = s;
}

// outlet rows (drain id 0) copy pend(i-1) (f90:607-608); by construction that is only cell n.
__global__ void k_basics_outlet(const int32_t *__restrict__ st, int64_t n, float *pend)
{
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        int64_t i = n - 1;
        if (st[i] == 0) {
            float s = n >= 2 ? pend[i - 1] : 0.0f;
            if (s == 0.0f) s = 0.001f;
            pend[i] = s;
        }
    }
}

// deterministic fp64 sum of a float vector: fixed 1024-block partials, then one block.
__global__ void k_sum_partial(const float *__restrict__ x, int64_t n, double *__restrict__ part)
{
    __shared__ double sh[256];
    int64_t per = (n + gridDim.x - 1) / gridDim.x;
    int64_t lo = (int64_t)blockIdx.x * per, hi = lo + per < n ? lo + per : n;
    double s = 0.0;
    for (int64_t i = lo + threadIdx.x; i < hi; i += blockDim.x) s += (double)x[i];
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int k = 128; k > 0; k >>= 1) {
        if ((int)threadIdx.x < k) sh[threadIdx.x] += sh[threadIdx.x + k];
        __syncthreads();
    }
    if (threadIdx.x == 0) part[blockIdx.x] = sh[0];
}
__global__ void k_sum_final(const double *__restrict__ part, int np, double *out)
{
    __shared__ double sh[256];
    double s = 0.0;
    for (int i = threadIdx.x; i < np; i += blockDim.x) s += part[i];
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int k = 128; k > 0; k >>= 1) {
        if ((int)threadIdx.x < k) sh[threadIdx.x] += sh[threadIdx.x + k];
        __syncthreads();
    }
    if (threadIdx.x == 0) *out = sh[0];
}

__global__ void k_coordxy(const int32_t *__restrict__ st, int64_t stride, int64_t n, float xll, float yll, float dx,
                          int32_t nrows, float *__restrict__ X, float *__restrict__ Y)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (X) X[i] = __fadd_rn(xll, __fmul_rn(dx, __fsub_rn((float)st[stride + i], 0.5f)));                 // f90:644
    if (Y) Y[i] = __fadd_rn(yll, __fmul_rn(dx, __fadd_rn((float)(nrows - st[2 * stride + i]), 0.5f)));   // f90:645
}

template <typename T>
__global__ void k_map2basin(const int32_t *__restrict__ st, int64_t stride, int64_t n, const T *__restrict__ raster,
                            int64_t nrows, int64_t ncols, T *__restrict__ vec)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int64_t c = st[stride + i] - 1, f = st[2 * stride + i] - 1;
    if (c >= 0 && c < ncols && f >= 0 && f < nrows) vec[i] = raster[f * ncols + c];
}
template <typename T>
__global__ void k_2map(const int32_t *__restrict__ st, int64_t stride, int64_t n, const T *__restrict__ vec,
                       int64_t nrows, int64_t ncols, T *__restrict__ raster)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int64_t c = st[stride + i] - 1, f = st[2 * stride + i] - 1;
    if (c >= 0 && c < ncols && f >= 0 && f < nrows) raster[f * ncols + c] = vec[i];
}

// stream / node marks, cuencas_heavy.f90:780-809
__global__ void k_nod_cauce(const int32_t *__restrict__ acum, int64_t n, int32_t umbral, int32_t *__restrict__ cauce,
                            int32_t *counts)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool c = i < n && acum[i] >= umbral;
    if (i < n) cauce[i] = c ? 1 : 0;
    unsigned b = __ballot_sync(0xffffffffu, c);
    if ((threadIdx.x & 31) == 0 && b) atomicAdd(&counts[1], __popc(b));
}
__global__ void k_nod_count(const int32_t *__restrict__ st, const int32_t *__restrict__ cauce, int64_t n, int32_t *nodos)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int32_t dr = st[i];
    if (dr != 0 && cauce[i] == 1) atomicAdd(&nodos[n - dr], 1);
}
__global__ void k_nod_fix(const int32_t *__restrict__ cauce, int64_t n, int32_t *__restrict__ nodos, int32_t *counts)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool isn = false;
    if (i < n) {
        int32_t v = nodos[i];
        if (v == 0 && cauce[i] == 1) v = 3;
        if (i == n - 1) v = 2;
        if (v < 2) v = 0;
        nodos[i] = v;
        isn = v > 0;
    }
    unsigned b = __ballot_sync(0xffffffffu, isn);
    if ((threadIdx.x & 31) == 0 && b) atomicAdd(&counts[0], __popc(b));
}
__global__ void k_nod_trazado(const int32_t *__restrict__ st, const int32_t *__restrict__ cauce,
                              const int32_t *__restrict__ nodos, int64_t n, int32_t *__restrict__ trazado)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int32_t dr = st[i];
    int32_t t = (i == n - 1) ? 1 : 0;
    if (dr != 0 && nodos[n - dr] != 0 && cauce[i] == 1) t = 1;
    trazado[i] = t;
}

}  // namespace

// =================================================================== C ABI
extern "C" {

size_t wmf_basin_mask_workspace(int64_t ny, int64_t nx) { return d8_basin_membership_workspace(ny, nx) + 1024; }

int wmf_basin_mask(const int8_t *dir, const uint8_t *mask, int64_t ny, int64_t nx, int encoding, int64_t outlet_r,
                   int64_t outlet_c, uint8_t *basin_mask, int64_t *ncells_host, void *workspace, size_t ws_bytes,
                   void *stream)
{
    cudaStream_t st = as_stream(stream);
    WMF_REQUIRE(dir && basin_mask && ny >= 0 && nx >= 0, WMF_E_ARG, "null pointer or negative size");
    WMF_REQUIRE(encoding == WMF_ENC_INTERNAL || encoding == WMF_ENC_KEYPAD, WMF_E_ARG, "unknown encoding");
    WMF_REQUIRE(ny * nx < ((int64_t)1 << 31), WMF_E_TOO_LARGE, "raster needs ny*nx < 2^31");
    WMF_REQUIRE(outlet_r >= 0 && outlet_r < ny && outlet_c >= 0 && outlet_c < nx, WMF_E_OUTLET_RANGE, "outlet outside grid");
    int64_t n = ny * nx, outlet = outlet_r * nx + outlet_c;
    if (mask) {
        uint8_t m = 0;
        WMF_CUDA_CHECK(cudaMemcpyAsync(&m, mask + outlet, 1, cudaMemcpyDeviceToHost, st));
        WMF_CUDA_CHECK(cudaStreamSynchronize(st));
        WMF_REQUIRE(m != 0, WMF_E_OUTLET_MASK, "outlet not in mask");
    }
    WsCursor cur(workspace, ws_bytes);
    unsigned long long *d_count = cur.take<unsigned long long>(1);
    const size_t mws = d8_basin_membership_workspace(ny, nx);
    void *mw = cur.take<char>(mws);
    WMF_REQUIRE(cur.ok, WMF_E_WORKSPACE, "workspace too small");
    (void)n;
    WMF_CUDA_CHECK(cudaMemsetAsync(d_count, 0, sizeof(unsigned long long), st));
    int rc = d8_basin_membership(dir, mask, ny, nx, encoding, outlet, basin_mask, nullptr, d_count, mw, mws, st);
    if (rc) return rc;
    unsigned long long h = 0;
    WMF_CUDA_CHECK(cudaMemcpyAsync(&h, d_count, sizeof(h), cudaMemcpyDeviceToHost, st));
    WMF_CUDA_CHECK(cudaStreamSynchronize(st));
    if (ncells_host) *ncells_host = (int64_t)h;
    return 0;
}

void wmf_coord2fil_col_host(float x, float y, float xll, float yll, float dx, int32_t ncols, int32_t nrows,
                            int32_t *col, int32_t *fil)
{
    // cuencas_heavy.f90:355-365, fp32 throughout (f2py casts the arguments to `real`)
    volatile float qx = (x - xll) / dx, qy = (y - yll) / dx;
    int32_t c = (int32_t)ceilf(qx);
    if (c < 0) c = -c;
    int32_t f = nrows - (int32_t)floorf(qy);
    if (c > ncols || c <= 0) c = -999;
    if (f > nrows || f <= 0) f = -999;
    *col = c;
    *fil = f;
}

static size_t structure_ws(int64_t n, int64_t nb, int64_t nrows, int64_t ncols)
{
    size_t scan_tmp = 0, sort_tmp = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, scan_tmp, (int32_t *)nullptr, (int32_t *)nullptr, (int)n);
    cub::DeviceRadixSort::SortPairs(nullptr, sort_tmp, (unsigned long long *)nullptr, (unsigned long long *)nullptr,
                                    (int32_t *)nullptr, (int32_t *)nullptr, (int)nb);
    size_t s = 0;
    s += 2 * wmf_align((size_t)n * 4);                   // flag32, cidx
    s += wmf_align(d8_basin_membership_workspace(nrows, ncols));
    s += wmf_align(scan_tmp) + wmf_align(sort_tmp);
    s += 13 * wmf_align((size_t)nb * 4);                 // cell,parentC,size,cnt,off, 2x(ptr,pre,dep), vals x2
    s += 2 * wmf_align((size_t)nb * 8);                  // keys x2
    return s + 4096;
}

size_t wmf_basin_structure_workspace(int64_t nrows, int64_t ncols)
{
    return structure_ws(nrows * ncols, nrows * ncols, nrows, ncols);
}

int wmf_basin_structure(const int8_t *dir, int64_t nrows, int64_t ncols, int32_t col, int32_t fil, int32_t *structure,
                        int64_t capacity, int64_t *ncells_host, void *workspace, size_t ws_bytes, void *stream)
{
    cudaStream_t st = as_stream(stream);
    WMF_REQUIRE(dir && structure && ncells_host && nrows > 0 && ncols > 0, WMF_E_ARG, "null pointer or empty raster");
    WMF_REQUIRE(nrows * ncols < ((int64_t)1 << 31), WMF_E_TOO_LARGE, "raster needs nrows*ncols < 2^31");
    WMF_REQUIRE(col >= 1 && col <= ncols && fil >= 1 && fil <= nrows, WMF_E_OUTLET_RANGE, "outlet outside grid");
    const int64_t n = nrows * ncols, ny = nrows, nx = ncols;
    const int64_t outlet = (int64_t)(fil - 1) * ncols + (col - 1);
    const int th = 256;
    {   // the outlet's own code: 5 would re-enqueue itself (f90:538 with kf=kc=2)
        int8_t code = 0;
        WMF_CUDA_CHECK(cudaMemcpyAsync(&code, dir + outlet, 1, cudaMemcpyDeviceToHost, st));
        WMF_CUDA_CHECK(cudaStreamSynchronize(st));
        WMF_REQUIRE(code != 5, WMF_E_CODE5, "outlet cell has keypad code 5 (drains to itself)");
    }
    WsCursor cur(workspace, ws_bytes);
    const size_t mws = d8_basin_membership_workspace(ny, nx);
    void *mw = cur.take<char>(mws);
    int32_t *flag32 = cur.take<int32_t>((size_t)n);
    int32_t *cidx = cur.take<int32_t>((size_t)n);
    unsigned long long *d_count = cur.take<unsigned long long>(1);
    size_t scan_tmp_b = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, scan_tmp_b, flag32, cidx, (int)n);
    void *scan_tmp = cur.take<char>(scan_tmp_b);
    WMF_REQUIRE(cur.ok, WMF_E_WORKSPACE, "workspace too small");

    WMF_CUDA_CHECK(cudaMemsetAsync(d_count, 0, sizeof(unsigned long long), st));
    int rc = d8_basin_membership(dir, nullptr, ny, nx, WMF_ENC_KEYPAD, outlet, nullptr, flag32, d_count, mw, mws, st);
    if (rc) return rc;
    unsigned long long h_count = 0;
    WMF_CUDA_CHECK(cudaMemcpyAsync(&h_count, d_count, sizeof(h_count), cudaMemcpyDeviceToHost, st));
    // cycle through the outlet: the outlet's own downstream path comes back to it
    int32_t h_ptr_recv = -1;
    int64_t recv_outlet = -1;
    {
        int8_t code = 0;
        WMF_CUDA_CHECK(cudaMemcpyAsync(&code, dir + outlet, 1, cudaMemcpyDeviceToHost, st));
        WMF_CUDA_CHECK(cudaStreamSynchronize(st));
        int k = d8_norm<WMF_ENC_KEYPAD>(code);
        if (k != 8) {
            int64_t rr = (fil - 1) + d8_dr(k), cc = (col - 1) + d8_dc(k);
            if (rr >= 0 && rr < ny && cc >= 0 && cc < nx) recv_outlet = rr * nx + cc;
        }
    }
    if (recv_outlet >= 0) {
        WMF_CUDA_CHECK(cudaMemcpyAsync(&h_ptr_recv, flag32 + recv_outlet, 4, cudaMemcpyDeviceToHost, st));   // is the receiver in the basin?
    }
    WMF_CUDA_CHECK(cudaStreamSynchronize(st));
    WMF_REQUIRE(!(recv_outlet >= 0 && h_ptr_recv == 1), WMF_E_CYCLE, "drainage cycle through the outlet");
    const int64_t nb = (int64_t)h_count;
    *ncells_host = nb;
    WMF_REQUIRE(nb <= capacity, WMF_E_CAPACITY, "structure capacity too small");

    WMF_CUDA_CHECK(cub::DeviceScan::ExclusiveSum(scan_tmp, scan_tmp_b, flag32, cidx, (int)n, st));
    int32_t *cell = cur.take<int32_t>((size_t)nb), *parentC = cur.take<int32_t>((size_t)nb);
    int32_t *size = cur.take<int32_t>((size_t)nb), *cnt = cur.take<int32_t>((size_t)nb);
    int32_t *off = cur.take<int32_t>((size_t)nb);
    int32_t *pA = cur.take<int32_t>((size_t)nb), *qA = cur.take<int32_t>((size_t)nb), *dA = cur.take<int32_t>((size_t)nb);
    int32_t *pB = cur.take<int32_t>((size_t)nb), *qB = cur.take<int32_t>((size_t)nb), *dB = cur.take<int32_t>((size_t)nb);
    int32_t *vals = cur.take<int32_t>((size_t)nb), *vals2 = cur.take<int32_t>((size_t)nb);
    unsigned long long *keys = cur.take<unsigned long long>((size_t)nb), *keys2 = cur.take<unsigned long long>((size_t)nb);
    size_t sort_tmp_b = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, sort_tmp_b, keys, keys2, vals, vals2, (int)nb);
    void *sort_tmp = cur.take<char>(sort_tmp_b);
    WMF_REQUIRE(cur.ok, WMF_E_WORKSPACE, "workspace too small for this basin size");

    k_scatter_cells<<<blocks_for(n, th), th, 0, st>>>(flag32, cidx, n, cell);
    k_parent_compact<<<blocks_for(nb, th), th, 0, st>>>(dir, cell, cidx, nb, ny, nx, (int32_t)outlet, parentC);
    // subtree sizes (the basin is a tree rooted at the outlet: doubling, not the chain walk whose run time is the main stem).
    // Scratch: pA, qA, dA and pB follow each other in the workspace and are not in use yet.
    (void)cnt;
    rc = forest_subtree_sums_i32(parentC, nullptr, size, pA, nb, st);
    if (rc) return rc;
    k_sibling_offset<<<blocks_for(nb, th), th, 0, st>>>(dir, cell, cidx, parentC, size, nb, ny, nx, off);
    k_jump_init<<<blocks_for(nb, th), th, 0, st>>>(parentC, off, nb, pA, qA, dA);
    int rounds = 1;
    while (((int64_t)1 << rounds) < nb + 1) rounds++;
    for (int r = 0; r < rounds; r++) {
        k_jump_round<<<blocks_for(nb, th), th, 0, st>>>(pA, qA, dA, pB, qB, dB, nb);
        int32_t *t;
        t = pA; pA = pB; pB = t;
        t = qA; qA = qB; qB = t;
        t = dA; dA = dB; dB = t;
    }
    k_make_keys<<<blocks_for(nb, th), th, 0, st>>>(qA, dA, nb, keys, vals);
    WMF_LAUNCH_CHECK();
    WMF_CUDA_CHECK(cub::DeviceRadixSort::SortPairs(sort_tmp, sort_tmp_b, keys, keys2, vals, vals2, (int)nb, 0, 64, st));
    int32_t *pos = pB;                                    // reuse
    k_positions<<<blocks_for(nb, th), th, 0, st>>>(vals2, nb, pos);
    k_write_structure<<<blocks_for(nb, th), th, 0, st>>>(vals2, pos, parentC, cell, nb, nx, structure, capacity);
    WMF_LAUNCH_CHECK();
    WMF_CUDA_CHECK(cudaStreamSynchronize(st));
    return 0;
}

size_t wmf_basin_vec_workspace(int64_t n)
{
    size_t sort_tmp = 0;
    cub::DeviceRadixSort::SortKeys(nullptr, sort_tmp, (float *)nullptr, (float *)nullptr, (int)n);
    return 6 * wmf_align((size_t)n * 4) + wmf_align(sort_tmp) + wmf_align(1024 * 8) + 4096;
}

int wmf_basin_acum_vec(const int32_t *structure, int64_t stride, int64_t n, int32_t *acum, void *workspace,
                       size_t ws_bytes, void *stream)
{
    cudaStream_t st = as_stream(stream);
    WMF_REQUIRE(structure && acum && n >= 0 && stride >= n, WMF_E_ARG, "bad argument");
    WMF_REQUIRE(n < ((int64_t)1 << 31), WMF_E_TOO_LARGE, "n must be < 2^31");
    if (n == 0) return 0;
    WsCursor cur(workspace, ws_bytes);
    int32_t *parent = cur.take<int32_t>((size_t)n), *scratch = cur.take<int32_t>((size_t)n);
    cur.take<int32_t>((size_t)n);
    cur.take<int32_t>((size_t)n);
    cur.take<int32_t>((size_t)n);                       // (scratch: 3 n + 64, contiguous)
    WMF_REQUIRE(cur.ok, WMF_E_WORKSPACE, "workspace too small");
    k_vec_parent<<<blocks_for(n, 256), 256, 0, st>>>(structure, n, parent);
    WMF_LAUNCH_CHECK();
    return forest_subtree_sums_i32(parent, nullptr, acum, scratch, n, st);
}

int wmf_basin_coordxy(const int32_t *structure, int64_t stride, int64_t n, float xll, float yll, float dx,
                      int32_t nrows, float *X, float *Y, void *stream)
{
    WMF_REQUIRE(structure && n >= 0 && stride >= n, WMF_E_ARG, "bad argument");
    if (n == 0) return 0;
    k_coordxy<<<blocks_for(n, 256), 256, 0, as_stream(stream)>>>(structure, stride, n, xll, yll, dx, nrows, X, Y);
    WMF_LAUNCH_CHECK();
    return 0;
}

int wmf_basin_basics(const int32_t *structure, int64_t stride, int64_t n, const float *dem, const int32_t *dirv,
                     const float *geo5_host, int32_t *acum, float *long_cel, float *pend, float *elev,
                     float *scalars_host, void *workspace, size_t ws_bytes, void *stream)
{
    cudaStream_t st = as_stream(stream);
    WMF_REQUIRE(structure && dem && dirv && geo5_host && acum && long_cel && pend && elev && scalars_host && n > 0 &&
                    stride >= n, WMF_E_ARG, "bad argument");
    const float dxp = geo5_host[0], xll = geo5_host[1], yll = geo5_host[2], dx = geo5_host[3];
    const int32_t nrows = (int32_t)geo5_host[4];
    int rc = wmf_basin_acum_vec(structure, stride, n, acum, workspace, ws_bytes, stream);
    if (rc) return rc;
    WsCursor cur(workspace, ws_bytes);
    cur.take<int32_t>((size_t)n);
    cur.take<int32_t>((size_t)n);                       // parent, cnt (in use until the walk is done: same stream)
    float *X = cur.take<float>((size_t)n), *Y = cur.take<float>((size_t)n);
    float *Xs = cur.take<float>((size_t)n), *Ys = cur.take<float>((size_t)n);
    double *part = cur.take<double>(1024);
    size_t sort_tmp_b = 0;
    cub::DeviceRadixSort::SortKeys(nullptr, sort_tmp_b, X, Xs, (int)n);
    void *sort_tmp = cur.take<char>(sort_tmp_b);
    WMF_REQUIRE(cur.ok, WMF_E_WORKSPACE, "workspace too small");
    k_basics<<<blocks_for(n, 256), 256, 0, st>>>(structure, n, dem, dirv, dxp, long_cel, pend, elev);
    k_basics_outlet<<<1, 32, 0, st>>>(structure, n, pend);
    int np = (int)(n < 1024 * 256 ? (n + 255) / 256 : 1024);
    double *d_sums = part + 1024 - 2;                    // last two slots hold the results
    np = np > 1000 ? 1000 : np;
    k_sum_partial<<<np, 256, 0, st>>>(pend, n, part);
    k_sum_final<<<1, 256, 0, st>>>(part, np, d_sums);
    double h_sums[2];
    WMF_CUDA_CHECK(cudaMemcpyAsync(&h_sums[0], d_sums, 8, cudaMemcpyDeviceToHost, st));
    WMF_CUDA_CHECK(cudaStreamSynchronize(st));
    k_sum_partial<<<np, 256, 0, st>>>(dem, n, part);
    k_sum_final<<<1, 256, 0, st>>>(part, np, d_sums + 1);
    WMF_CUDA_CHECK(cudaMemcpyAsync(&h_sums[1], d_sums + 1, 8, cudaMemcpyDeviceToHost, st));
    k_coordxy<<<blocks_for(n, 256), 256, 0, st>>>(structure, stride, n, xll, yll, dx, nrows, X, Y);
    WMF_LAUNCH_CHECK();
    WMF_CUDA_CHECK(cub::DeviceRadixSort::SortKeys(sort_tmp, sort_tmp_b, X, Xs, (int)n, 0, 32, st));
    WMF_CUDA_CHECK(cub::DeviceRadixSort::SortKeys(sort_tmp, sort_tmp_b, Y, Ys, (int)n, 0, 32, st));
    int64_t m = (n % 2 == 0) ? n / 2 : (n + 1) / 2;      // 1-based median slot, f90:619-625
    float cx = 0, cy = 0;
    WMF_CUDA_CHECK(cudaMemcpyAsync(&cx, Xs + (m - 1), 4, cudaMemcpyDeviceToHost, st));
    WMF_CUDA_CHECK(cudaMemcpyAsync(&cy, Ys + (m - 1), 4, cudaMemcpyDeviceToHost, st));
    WMF_CUDA_CHECK(cudaStreamSynchronize(st));
    scalars_host[0] = (float)n * (dxp * dxp) / 1e6f;     // f90:613
    scalars_host[1] = (float)(h_sums[0] / (double)n);    // f90:614 (fp64 tree instead of an fp32 running sum)
    scalars_host[2] = (float)(h_sums[1] / (double)n);    // f90:615
    scalars_host[3] = cx;
    scalars_host[4] = cy;
    scalars_host[5] = 0.0f;
    return 0;
}

int wmf_basin_map2basin(const int32_t *structure, int64_t stride, int64_t n, const void *raster, int64_t nrows,
                        int64_t ncols, int elem_size, void *vec, void *stream)
{
    cudaStream_t st = as_stream(stream);
    WMF_REQUIRE(structure && raster && vec && n >= 0 && stride >= n, WMF_E_ARG, "bad argument");
    if (n == 0) return 0;
    unsigned g = blocks_for(n, 256);
    if (elem_size == 1) k_map2basin<uint8_t><<<g, 256, 0, st>>>(structure, stride, n, (const uint8_t *)raster, nrows, ncols, (uint8_t *)vec);
    else if (elem_size == 4) k_map2basin<uint32_t><<<g, 256, 0, st>>>(structure, stride, n, (const uint32_t *)raster, nrows, ncols, (uint32_t *)vec);
    else if (elem_size == 8) k_map2basin<uint64_t><<<g, 256, 0, st>>>(structure, stride, n, (const uint64_t *)raster, nrows, ncols, (uint64_t *)vec);
    else WMF_REQUIRE(false, WMF_E_ARG, "elem_size must be 1, 4 or 8");
    WMF_LAUNCH_CHECK();
    return 0;
}

int wmf_basin_2map(const int32_t *structure, int64_t stride, int64_t n, const void *vec, int64_t nrows, int64_t ncols,
                   int elem_size, void *raster, void *stream)
{
    cudaStream_t st = as_stream(stream);
    WMF_REQUIRE(structure && raster && vec && n >= 0 && stride >= n, WMF_E_ARG, "bad argument");
    if (n == 0) return 0;
    unsigned g = blocks_for(n, 256);
    if (elem_size == 1) k_2map<uint8_t><<<g, 256, 0, st>>>(structure, stride, n, (const uint8_t *)vec, nrows, ncols, (uint8_t *)raster);
    else if (elem_size == 4) k_2map<uint32_t><<<g, 256, 0, st>>>(structure, stride, n, (const uint32_t *)vec, nrows, ncols, (uint32_t *)raster);
    else if (elem_size == 8) k_2map<uint64_t><<<g, 256, 0, st>>>(structure, stride, n, (const uint64_t *)vec, nrows, ncols, (uint64_t *)raster);
    else WMF_REQUIRE(false, WMF_E_ARG, "elem_size must be 1, 4 or 8");
    WMF_LAUNCH_CHECK();
    return 0;
}

int wmf_stream_nod(const int32_t *structure, int64_t stride, int64_t n, const int32_t *acum, int32_t umbral,
                   int32_t *cauce, int32_t *nodos, int32_t *trazado, int32_t *counts_host, void *stream)
{
    cudaStream_t st = as_stream(stream);
    WMF_REQUIRE(structure && acum && cauce && nodos && trazado && counts_host && n > 0 && stride >= n, WMF_E_ARG, "bad argument");
    int32_t *d_counts = nullptr;
    WMF_CUDA_CHECK(cudaMallocAsync((void **)&d_counts, 2 * sizeof(int32_t), st));
    WMF_CUDA_CHECK(cudaMemsetAsync(d_counts, 0, 2 * sizeof(int32_t), st));
    WMF_CUDA_CHECK(cudaMemsetAsync(nodos, 0, sizeof(int32_t) * n, st));
    unsigned g = blocks_for(n, 256);
    k_nod_cauce<<<g, 256, 0, st>>>(acum, n, umbral, cauce, d_counts);
    k_nod_count<<<g, 256, 0, st>>>(structure, cauce, n, nodos);
    k_nod_fix<<<g, 256, 0, st>>>(cauce, n, nodos, d_counts);
    k_nod_trazado<<<g, 256, 0, st>>>(structure, cauce, nodos, n, trazado);
    WMF_LAUNCH_CHECK();
    WMF_CUDA_CHECK(cudaMemcpyAsync(counts_host, d_counts, 2 * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    WMF_CUDA_CHECK(cudaFreeAsync(d_counts, st));
    WMF_CUDA_CHECK(cudaStreamSynchronize(st));
    return 0;
}

}  // extern "C"
