/*
 * wmf_b200.h -- C ABI of libwmfb200.so: the B200 (sm_100a) implementation of WMF's
 * raster-drainage hot path (D8 accumulation + basin delineation, SHIA cell update).
 *
 * This is the drop-in boundary.  The reference reaches its kernels through the
 * namespace `cu`, the name `shia_v1` and module-level state on `cu` / `models`
 * (wmf_heavy/wmfv2.py:36-79); each entry point below names the reference routine it
 * replaces.  INTEGRATION.md shows the ctypes binding a maintainer adds at wmfv2.py:36-79.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in `_host`;
 *   - `stream` is a cudaStream_t passed as void* (NULL = default stream); calls are
 *     asynchronous unless noted ("syncs") and re-entrant per stream;
 *   - return value: 0 ok, <0 argument/semantic error (WMF_E_*), >0 a cudaError_t;
 *     wmf_last_error() returns a thread-local message for the last failure;
 *   - the library keeps no data between calls (the f2py-style module state of the reference lives in the
 *     Python shim, not here); the only lasting objects are helper streams and events per device, created on
 *     first use (wmf_d8_stripe_local: one stream, two events; the optional band schedule of wmf_d8_accum: two
 *     streams and their events).  Enqueueing through them is serialised by a per-device mutex inside the call;
 *   - threading: a call only touches the buffers it is given, so calls from several host threads are
 *     safe as long as they do not share a workspace or an output (give every thread its own workspace; two
 *     calls that share one must be ordered on one stream).  wmf_last_error() is per thread;
 *   - workspaces are caller-owned: query the size, allocate, pass in;
 *   - environment switches, read at every call (all off by default; every setting gives bit-identical
 *     results and is covered by the GPU tests):  WMF_ACC_WIDE=1 keeps 64-bit arithmetic for 64-bit outputs of
 *     rasters below 2^32 cells;  WMF_D8_BULK=1 stages tile rows with TMA bulk copies (cp.async.bulk) instead
 *     of cp.async;  WMF_D8_BANDS=B runs the forest phase of wmf_d8_accum band by band beside phase A
 *     (WMF_D8_BAND_SMEM: bytes of shared-memory padding that hold phase A to fewer blocks per SM there).
 */
#ifndef WMF_B200_H
#define WMF_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define WMF_OK 0
#define WMF_E_ARG (-1)        /* bad argument (null pointer, negative size, unknown enum) */
#define WMF_E_WORKSPACE (-2)  /* workspace too small */
#define WMF_E_OUTLET_RANGE (-3) /* outlet outside grid  (basics.py:225-226) */
#define WMF_E_OUTLET_MASK (-4)  /* outlet not in mask   (basics.py:227-228) */
#define WMF_E_CODE5 (-5)      /* keypad code 5 inside the basin (self-draining cell, f90:538) */
#define WMF_E_CYCLE (-6)      /* drainage cycle through the outlet (Fortran BFS would not end) */
#define WMF_E_TOO_LARGE (-7)  /* raster exceeds the index width of this entry point */
#define WMF_E_CAPACITY (-8)   /* output capacity too small */

/* DIR encodings */
#define WMF_ENC_INTERNAL 0 /* wmf_py: 0 E,1 SE,2 S,3 SW,4 W,5 NW,6 N,7 NE (basics.py:37-46) */
#define WMF_ENC_KEYPAD 1   /* Fortran: 7 8 9 / 4 5 6 / 1 2 3, rows downwards (cuencas_heavy.f90:1338-1362) */

/* accumulation output types */
#define WMF_ACC_I32 0
#define WMF_ACC_I64 1
#define WMF_ACC_F64 2 /* what wmf_py.basin_acum returns (integral values) */

/* accumulation algorithms (wmf_d8_accum `algo`) */
#define WMF_ALGO_AUTO 0
#define WMF_ALGO_WALK 1  /* in-degree driven chain walk with L2 atomics: any input */
#define WMF_ALGO_SWEEP 2 /* warp strip sweeps + contracted boundary forest: any input, fastest when flow is sweep-coherent */

const char *wmf_last_error(void);
int wmf_version(void);
/* device properties the host layer needs: out[0]=SM count, out[1]=cc major, out[2]=cc minor, out[3]=L2 bytes */
int wmf_device_info(int64_t *out4);

/* ---------------------------------------------------------------- D1: full-raster D8 accumulation
 * Replaces wmf_py/cu_py/basics.py:124-177 basin_acum(flowdir, mask).
 * dir: int8 [ny][nx]; mask: uint8 [ny][nx] (nonzero = active) or NULL (= all active);
 * acc: [ny][nx] of acc_dtype.  Cells outside the mask get 0; cells on or below a
 * drainage cycle keep the partial sum of their finished tree inflows (reference semantics).
 * Bit-exact for any schedule (integer adds). */
size_t wmf_d8_accum_workspace(int64_t ny, int64_t nx, int acc_dtype, int algo);
int wmf_d8_accum(const int8_t *dir, const uint8_t *mask, void *acc, int acc_dtype, int64_t ny, int64_t nx,
                 int encoding, int algo, void *workspace, size_t ws_bytes, void *stream);

/* Diagnostic: how the tiles of the last WMF_ALGO_SWEEP accumulation (32-bit arithmetic: fewer than 2^32 cells) that used
 * `workspace` were solved.  counts8_host[c] (HOST pointer, 8 entries) = tiles of class c: 0 one-pass row sweeps, 4 jump solver
 * (tiles with upward flow), 1 / 3 chain walk (cycles), 2 chain walk because an upstream never resolves.  Syncs the stream. */
int wmf_d8_accum_tile_classes(const void *workspace, size_t ws_bytes, int64_t ny, int64_t nx, int64_t *counts8_host, void *stream);

/* ---------------------------------------------------------------- D1 on row stripes
 * A raster of ny_total rows is cut into row stripes, one per GPU; this stripe holds rows
 * [row0, row0 + ny), ny >= 2 and fewer than 2^31 cells.  Device phases around ONE exchange:
 *   wmf_d8_stripe_local  solves the stripe in isolation and condenses it to packed boundary records,
 *                        int64 [2][2][nx] (index side*nx + col in a plane; side 0 = first row, 1 = last row):
 *                          plane 0  exit_acc: stripe-local accumulation of a cell draining across the
 *                                   stripe edge (0 elsewhere)
 *                          plane 1  link (low 32 bits: boundary index where flow ENTERING at this cell
 *                                   leaves the stripe again, -1 if it ends inside) | k << 32 (normalised D8
 *                                   code 0..7 clockwise from E, 8 none, 9 inactive) | unres << 40 (that
 *                                   accumulation is unresolved: cycle upstream)
 *   (exchange: ONE all-gather of the records -> recs_all int64 [G][2][2][nx])
 *   wmf_stripe_boundary_inflow  every rank solves the 2*G*nx-node boundary forest and sums, for each
 *                        boundary cell of stripe `me`, what arrives from the neighbouring stripes:
 *                        ext_inflow int64 [2][nx], ext_unres uint8 [2][nx]
 *   wmf_d8_stripe_finish injects that inflow and writes the final acc [ny][nx] (WMF_ACC_I32 only when
 *                        ny_total*nx < 2^31, WMF_ACC_I64, WMF_ACC_F64).
 * The SAME workspace must be passed to stripe_local and stripe_finish (phase-A state lives in it).  * wmf_d8_stripe_finish CONSUMES the workspace of the local pass (it adds the routed inflow into arrays that pass cleared):
 * one finish per local. */
size_t wmf_d8_stripe_workspace(int64_t ny, int64_t nx);
int wmf_d8_stripe_local(const int8_t *dir, const uint8_t *mask, int64_t ny, int64_t nx, int64_t row0, int64_t ny_total,
                        int encoding, int64_t *records, void *workspace, size_t ws_bytes, void *stream);
size_t wmf_stripe_boundary_workspace(int G, int64_t nx);
int wmf_stripe_boundary_inflow(const int64_t *recs_all, int G, int64_t nx, int me, int64_t *ext_inflow,
                               uint8_t *ext_unres, void *workspace, size_t ws_bytes, void *stream);
int wmf_d8_stripe_finish(const int8_t *dir, const uint8_t *mask, void *acc, int acc_dtype, int64_t ny, int64_t nx,
                         int64_t row0, int64_t ny_total, int encoding, const int64_t *ext_inflow,
                         const uint8_t *ext_unres, void *workspace, size_t ws_bytes, void *stream);

/* Accumulation over an explicit forest: acc[i] = weight[i] + sum over children acc[child] (parent int32,
 * -1 = root; weight nullable = 1; blocked nullable: nodes that never finalise).  Nodes with an
 * unfinished (cyclic or blocked) upstream keep the partial sum of finished children without their own
 * weight and get resolved[i] = 0 (resolved nullable).  The vector-form accumulation of cu.basin_basics
 * (cuencas_heavy.f90:593-606) and the stripe boundary forest are instances. */
size_t wmf_forest_accum_workspace(int64_t n);
int wmf_forest_accum(const int32_t *parent, const int64_t *weight, const uint8_t *blocked, int64_t n, int64_t *acc,
                     uint8_t *resolved, void *workspace, size_t ws_bytes, void *stream);

/* ---------------------------------------------------------------- D2: basin mask (raster form)
 * Replaces the traversal of wmf_py/cu_py/basics.py:206-248 basin_cut: basin_mask[c] = 1 iff c is
 * active and its D8 path through active cells reaches (outlet_r, outlet_c).  Syncs (reads back the
 * cell count).  Needs ny*nx < 2^31. */
size_t wmf_basin_mask_workspace(int64_t ny, int64_t nx);
int wmf_basin_mask(const int8_t *dir, const uint8_t *mask, int64_t ny, int64_t nx, int encoding,
                   int64_t outlet_r, int64_t outlet_c, uint8_t *basin_mask, int64_t *ncells_host,
                   void *workspace, size_t ws_bytes, void *stream);

/* ---------------------------------------------------------------- D3: basin structure (vector form)
 * Replaces cu.basin_find + cu.basin_cut (wmf_heavy/cuencas_heavy.f90:507-576; call sites
 * wmfv2.py:823-824).  dir: keypad int8 [nrows][ncols] (== Fortran DIR(col, fil)); outlet
 * (col, fil) 1-based, from wmf_coord2fil_col_host.  Output structure int32 [3][capacity]:
 * rows = drain id, col, fil, written with row stride `capacity`; *ncells_host = n.  Order is
 * bit-identical to the Fortran BFS (upstream-first).  Syncs. */
size_t wmf_basin_structure_workspace(int64_t nrows, int64_t ncols);
int wmf_basin_structure(const int8_t *dir, int64_t nrows, int64_t ncols, int32_t col, int32_t fil,
                        int32_t *structure, int64_t capacity, int64_t *ncells_host, void *workspace,
                        size_t ws_bytes, void *stream);
/* cuencas_heavy.f90:355-365 in fp32, host side; col/fil = -999 when out of range. */
void wmf_coord2fil_col_host(float x, float y, float xll, float yll, float dx, int32_t ncols, int32_t nrows,
                            int32_t *col, int32_t *fil);

/* ---------------------------------------------------------------- D4/D6: vector-form accumulation + basics
 * structure: int32 [3][n] with row stride `stride`.  wmf_basin_acum_vec == cu.basin_acum(structure,
 * ncells) (wmfv2.py:1868) == the acum of basin_basics (cuencas_heavy.f90:593-606). */
size_t wmf_basin_vec_workspace(int64_t n);
int wmf_basin_acum_vec(const int32_t *structure, int64_t stride, int64_t n, int32_t *acum, void *workspace,
                       size_t ws_bytes, void *stream);
/* cuencas_heavy.f90:581-627.  geo = {dxp, xll, yll, dx, nrows(as float)}.  scalars_host[6] = area,
 * pend_media, elevacion, centroX, centroY, 0.  Syncs. */
int wmf_basin_basics(const int32_t *structure, int64_t stride, int64_t n, const float *dem, const int32_t *dirv,
                     const float *geo5_host, int32_t *acum, float *long_cel, float *pend, float *elev,
                     float *scalars_host, void *workspace, size_t ws_bytes, void *stream);
/* cuencas_heavy.f90:639-646 */
int wmf_basin_coordxy(const int32_t *structure, int64_t stride, int64_t n, float xll, float yll, float dx,
                      int32_t nrows, float *X, float *Y, void *stream);
/* cu.basin_map2basin / cu.basin_2map (wmfv2.py:1085,1167-1169): gather/scatter by (col, fil).
 * elem_size 1, 4 or 8 bytes; raster is [nrows][ncols]. */
int wmf_basin_map2basin(const int32_t *structure, int64_t stride, int64_t n, const void *raster, int64_t nrows,
                        int64_t ncols, int elem_size, void *vec, void *stream);
int wmf_basin_2map(const int32_t *structure, int64_t stride, int64_t n, const void *vec, int64_t nrows,
                   int64_t ncols, int elem_size, void *raster /* pre-filled with nodata */, void *stream);

/* ---------------------------------------------------------------- D5: stream / node marks
 * Replaces cu.basin_stream_nod (cuencas_heavy.f90:780-809).  counts_host = {n_nodos, n_cauce}.  Syncs. */
int wmf_stream_nod(const int32_t *structure, int64_t stride, int64_t n, const int32_t *acum, int32_t umbral,
                   int32_t *cauce, int32_t *nodos, int32_t *trazado, int32_t *counts_host, void *stream);

/* ---------------------------------------------------------------- N2: stream receiver / HAND
 * wmf_stream_receiver replaces wmf_py/cu_py/metrics.py:189-268 hydro_distance_and_receiver: for every active
 * cell the linear index of the first stream cell on its downstream path (a_quien; -1 outside the mask or when
 * the path never meets a stream; a stream cell is its own receiver) and the path length to it (hdnd, nullable;
 * nan where a_quien is -1; steps cost dx when the row stays, dy when the column stays, hypot(dx, dy) diagonally).
 * wmf_hand replaces metrics.py:271-323 hand: max(dem - dem[first stream cell downstream], 0); nan outside the
 * mask, where no stream is met, and -- like the reference's topological order -- behind a non-stream cell of a
 * drainage cycle.  stream: uint8 raster, nonzero = stream cell.  Both sync (convergence checks).  ny*nx < 2^31. */
size_t wmf_stream_receiver_workspace(int64_t ny, int64_t nx);
int wmf_stream_receiver(const int8_t *dir, const uint8_t *mask, const uint8_t *stream, int64_t ny, int64_t nx, int encoding,
                        double dx, double dy, int32_t *a_quien, double *hdnd, void *workspace, size_t ws_bytes, void *stream_);
size_t wmf_hand_workspace(int64_t ny, int64_t nx);
int wmf_hand(const double *dem, const int8_t *dir, const uint8_t *mask, const uint8_t *stream, int64_t ny, int64_t nx,
             int encoding, double *hand, void *workspace, size_t ws_bytes, void *stream_);

/* wmf_path_sum: out[c] = sum of per-cell weights over c's downstream path, c included, the path's last cell (no
 * receiver) excluded -- the distance pass of basin_ppalstream_find (wmf_py/cu_py/metrics.py:420-432, with
 * active = the stream mask, w = NULL = step lengths dx / dy / hypot, cycle_nan = 0: a cell on a cycle never enters the
 * reference's topological order and keeps distance 0) and of basin_time_to_out (:583-597, cycle_nan = 1: nan for
 * cycle cells and everything that drains into them).  active (nullable uint8): inactive cells do not exist (nan out,
 * never a receiver).  Sums are formed by pointer jumping, i.e. in a different order than the reference's
 * outlet-to-source recurrence: fp64, 1e-12 relative.  wmf_time_to_out replaces basin_time_to_out (:528-599): Manning
 * travel time per cell with unit hydraulic radius (nan when the speed is not positive), then that path sum.
 * Both sync (convergence checks).  ny*nx < 2^31. */
size_t wmf_path_sum_workspace(int64_t ny, int64_t nx);
int wmf_path_sum(const int8_t *dir, const uint8_t *active, const double *w, int64_t ny, int64_t nx, int encoding, double dx,
                 double dy, int cycle_nan, double *out, void *workspace, size_t ws_bytes, void *stream_);
size_t wmf_time_to_out_workspace(int64_t ny, int64_t nx);
int wmf_time_to_out(const int8_t *dir, const double *slope, const double *manning, int64_t ny, int64_t nx, int encoding,
                    double dx, double dy, double *out, void *workspace, size_t ws_bytes, void *stream_);

/* N1: network nodes and node-to-node links (wmf_py/cu_py/streams.py:282-384 basin_netxy_find / basin_netxy_cut).
 * node[c] = 1 for stream cells whose degree (stream donors + stream receiver) differs from 2, and for the outlet
 * (linear index; -1 for none).  For every stream cell with a stream receiver, down_node / down_steps = the first node
 * met downstream and the number of moves to it (-1 / 0 when the path meets none: a dead end or a node-free loop);
 * up_node / up_steps likewise along the unique stream donors of non-node cells (the upstream end of the cell's edge).
 * The host layer numbers the nodes in row-major order and assembles the edge dictionaries.  Syncs.  ny*nx < 2^31. */
size_t wmf_stream_nodes_workspace(int64_t ny, int64_t nx);
int wmf_stream_nodes(const int8_t *dir, const uint8_t *stream, int64_t ny, int64_t nx, int encoding, int64_t outlet,
                     uint8_t *node, int32_t *down_node, int32_t *down_steps, int32_t *up_node, int32_t *up_steps,
                     void *workspace, size_t ws_bytes, void *stream_);

/* ---------------------------------------------------------------- S1/S2: SHIA 5-tank (Fortran semantics)
 * Replaces models.shia_v1 as shipped (wmf_heavy/Modelos_heavy.f90:169-548) incl. calc_speed
 * (:907-919): vertical cascade + lateral outflow of tanks 2..4 (discarded), fp32.
 * P parameter sets are advanced in one launch (NSGA-II populations, wmfv2.py:2060-2106).
 * Layouts (N cells, SoA):
 *   cell_static float [17][N]: rows 0-3 v_coef, 4-7 h_coef, 8-11 h_exp, 12 max_capilar,
 *                              13 max_gravita, 14 max_aquifer, 15 hill_long, 16 elem_area
 *   rain_records int32 [n_rec][N] (record k at row k-1); pos_evento int32 [N_reg], 1-based,
 *   1 == "no rain" (no read, Modelos:366-367); evp float [N_reg]; calib float [P][11];
 *   sto float [P][5][N] in/out; hspeed float [P][4][N] in/out (read only if hspeed_given);
 *   outputs: balance [P][N_reg]; mean_rain [N_reg]; acum_rain [N];
 *   mean_storage [P][5][N_reg] / mean_speed [P][4][N_reg] nullable.
 * Per-step diagnostics are reduced deterministically (fixed tree, fp64 partials). */
typedef struct {
    float dt;
    float retorno_gr;
    float retorno_aq;
    int32_t calc_niter;
    int32_t speed_type[3];
    int32_t hspeed_given;
    int32_t strict;       /* 1: the Fortran's own operations (RainInt / 1000., (S1 / H1) ** 0.6 through powf) instead of the
                           * fast forms (x * 0.001f, S1 * (1 / H1), exp2(0.6 * log2 x) on the SFU): ~2x slower, bit-level parity */
} wmf_shia5_params;
/* diagnostics != 0: any of the optional per-step series / per-cell accumulators is requested (mean_storage, mean_speed,
 * mean_retorno, mean_vfluxes, vfluxes, rc_coef, retorned) */
size_t wmf_shia5_workspace(int64_t N, int64_t N_reg, int32_t P, int32_t diagnostics);
int wmf_shia5_run(const wmf_shia5_params *params_host, int64_t N, int64_t N_reg, int64_t n_rec, int32_t P,
                  const float *cell_static, const int32_t *rain_records, const int32_t *pos_evento,
                  const float *evp, const float *calib, float *sto, float *hspeed, float *balance,
                  float *mean_rain, float *acum_rain, float *mean_storage, float *mean_speed,
                  void *workspace, size_t ws_bytes, void *stream);

/* N3: the same run with storage snapshots -- the save_storage / guarda_cond mechanism of
 * Modelos_heavy.f90:317-328, 496-500, where the storages after selected steps go to records of a direct-access
 * file.  snap_slot int32 [N_reg]: -1, or the slot s in [0, n_snap) that receives the five storages after that
 * step; snap_out float [P][n_snap][5][N].  The host layer writes the slots as records (models.write_float_basin). */
int wmf_shia5_run_snap(const wmf_shia5_params *params_host, int64_t N, int64_t N_reg, int64_t n_rec, int32_t P,
                       const float *cell_static, const int32_t *rain_records, const int32_t *pos_evento,
                       const float *evp, const float *calib, float *sto, float *hspeed, float *balance,
                       float *mean_rain, float *acum_rain, float *mean_storage, float *mean_speed,
                       const int32_t *snap_slot, int32_t n_snap, float *snap_out,
                       void *workspace, size_t ws_bytes, void *stream);

/* The same run with the optional outputs of the shipped shia_v1 (all nullable): mean_retorno float [P][N_reg]
 * (show_mean_retorno, Modelos_heavy.f90:525-527), mean_vfluxes float [P][4][N_reg] and vfluxes float [P][4][N] of the last
 * step (save_vfluxes, :429-433, 502-507), rc_coef float [P][2][N] (save_rc, :435-438, 543-544: accumulated vflux2 - vflux3
 * capped by the accumulated rain; the accumulated rain), retorned float [P][N] (Retorned, :424; retorned_per_step = 1: the
 * array is cleared after every step, as show_mean_retorno / save_retorno make the Fortran do, so it ends as zeros). */
int wmf_shia5_run_ext(const wmf_shia5_params *params_host, int64_t N, int64_t N_reg, int64_t n_rec, int32_t P,
                      const float *cell_static, const int32_t *rain_records, const int32_t *pos_evento,
                      const float *evp, const float *calib, float *sto, float *hspeed, float *balance,
                      float *mean_rain, float *acum_rain, float *mean_storage, float *mean_speed,
                      const int32_t *snap_slot, int32_t n_snap, float *snap_out,
                      float *mean_retorno, float *mean_vfluxes, float *vfluxes, float *rc_coef, float *retorned,
                      int32_t retorned_per_step, void *workspace, size_t ws_bytes, void *stream);

/* N4 (EXTENSION: the reference has no behaviour to match here): channel routing behind the shipped 5-tank
 * update.  Modelos_heavy.f90 declares tank 5, `drena` and the control points (:52-55, :91-92, :202-211) but never moves
 * water between cells, so Q == 0 as shipped.  Definition (the tests hold a CPU restatement of it): after the
 * shipped update of a cell, the lateral outflows of tanks 2-4 enter its tank 5 together with the outflow volumes of its
 * donors of the SAME step; tank 5 is a linear reservoir like tanks 2-4 (speed row 4, hill_long) and drains to the
 * receiver.  Q [n_ctrl][N_reg] in m3/s: row 0 = the outlet (the structure's last cell), row j > 0 = the cell with
 * ctrl == j.  Cells in the structure's upstream-first order; rank = max depth - depth (wmf_basin_depth_host), donors in
 * CSR form with ascending indices.  Speed type 1 only.  A skewed level/time wavefront of N_reg + max_rank diagonals:
 * with level_start (int32 [max_rank + 2], nullable: first cell of every rank -- the ranks ascend with the cell index in the
 * structure's BFS order) ONE cooperative launch walks them all, the blocks meeting at a barrier between diagonals; without
 * it, one launch per diagonal. */
int32_t wmf_basin_depth_host(const int32_t *drain, int64_t n, int32_t *depth);
size_t wmf_shia5_route_workspace(int64_t N);
int wmf_shia5_route_run(const wmf_shia5_params *params_host, int64_t N, int64_t N_reg, int64_t n_rec,
                        const float *cell_static, const int32_t *rain_records, const int32_t *pos_evento,
                        const float *evp, const float *calib, const int32_t *rank, int32_t max_rank,
                        const int32_t *don_ptr, const int32_t *don_idx, const int32_t *ctrl, int32_t n_ctrl,
                        const int32_t *level_start, float *sto, float *acum_rain, float *Q, void *workspace,
                        size_t ws_bytes, void *stream);

/* ---------------------------------------------------------------- C1: fitness of a calibration population
 * The two objectives nsgaii_element.__evalfunc__ applies to every individual's hydrograph (wmf_heavy/wmfv2.py:2194-2198;
 * the bodies of __eval_nash__ / __eval_q_pico__ are not in the reference, so they take their standard form): qsim float
 * [P][T] (one row per parameter set), qobs float [T] (non-finite entries are skipped), fitness double [P][2] =
 * {Nash-Sutcliffe efficiency, |max qsim - max qobs| / max qobs}. */
int wmf_fitness_nash_peak(const float *qsim, const float *qobs, int32_t P, int64_t T, double *fitness, void *stream);

/* ---------------------------------------------------------------- S4: SHIA 3-tank (wmf_py semantics)
 * Replaces wmf_py/models_py/core.py:64-295 shia_v1, fp64.  rain/et double [nsteps][ncell]
 * (et nullable, already multiplied by nothing: et_scale = dt/3600 or dt/86400 is applied
 * inside); cap/grav/aqu in/out [ncell]; q_super_last/q_base_last out [ncell];
 * sums double [nsteps+1][6] out = per-step {q_total, storage, P, ET, q_super, q_base} sums (row 0
 * holds the initial storage sum, row t+1 belongs to step t), from which the host layer forms q_outlet and the mass-balance diagnostics (core.py:234-256);
 * series pointers nullable [nsteps][ncell]. */
typedef struct {
    double dt, max_capilar, max_gravita, max_aquifer, drena, retorno_gr;
    double k_perc_cap_to_grav, k_perc_grav_to_aqu, k_baseflow, max_infil_mm_h;
    double et_scale;
    int32_t proportional;
    int32_t pad;
} wmf_shia3_params;
size_t wmf_shia3_workspace(int64_t ncell, int64_t nsteps);
int wmf_shia3_run(const wmf_shia3_params *params_host, int64_t nsteps, int64_t ncell, const double *rain,
                  const double *et, double *cap, double *grav, double *aqu, double *q_super_last,
                  double *q_base_last, double *sums, double *q_super_t, double *q_base_t, double *q_total_t,
                  double *cap_t, double *grav_t, double *aqu_t, void *workspace, size_t ws_bytes, void *stream);

/* ---------------------------------------------------------------- T2 / K1: direction-code reclassification
 * Replaces wmf_py/cu_py/basics.py:73-121 (_reclass, dir_reclass_opentopo / _rwatershed -> internal 0..7) and
 * cuencas_heavy.f90:1086-1119 (the same two conventions -> keypad).  Two passes, the decision in between is the host
 * layer's: wmf_dir_reclass_scan reports WHICH values the raster holds (presence: 256 bits, bit v+128 of the uint32[8]
 * array for v in -128..127; *oor = 1 when a value lies outside that range), from which the host picks "already
 * internal", the table, or the reference's ValueError("unknown D8 codes: ..."); wmf_dir_reclass_apply maps every
 * element through lut[v+128] (int64 [256], device) -- values beyond int8 pass through (keep_oor) or become oor_value.
 * `in` holds n integers of elem_bytes (1, 2, 4, 8) each; out is int64 [n] (numpy's astype(int)). */
int wmf_dir_reclass_scan(const void *in, int elem_bytes, int64_t n, uint32_t *presence, int32_t *oor, void *stream);
int wmf_dir_reclass_apply(const void *in, int elem_bytes, int64_t n, const int64_t *lut, int keep_oor, int64_t oor_value,
                          int64_t *out, void *stream);

/* ---------------------------------------------------------------- N3: basin <-> raster by linear index
 * Replaces wmf_py/cu_py/basics.py:251-301 (basin_2map / basin_map2basin).  wmf_index_scan sets *flag when an index lies
 * outside [0, ncell); wmf_index_repair rewrites idx as floor(idx / 100) * nx + (idx mod 100) mod nx (the reference's
 * "row*100 + col" repair, basics.py:273-276, Python floor semantics); wmf_scatter_last writes out[idx[i]] = vals[i] into
 * a raster the caller has filled with nodata -- with duplicate indices the last i wins, like the reference's sequential
 * assignment (owner: int32 [ncell] scratch); wmf_gather reads out[i] = flat[idx[i]].  Indices must be in range. */
int wmf_index_scan(const int64_t *idx, int64_t n, int64_t ncell, int32_t *flag, void *stream);
int wmf_index_repair(int64_t *idx, int64_t n, int64_t nx, void *stream);
int wmf_scatter_last(const void *vals, int elem_bytes, const int64_t *idx, int64_t n, void *out, int32_t *owner,
                     int64_t ncell, void *stream);
int wmf_gather(const void *flat, int elem_bytes, const int64_t *idx, int64_t n, void *out, void *stream);

/* ---------------------------------------------------------------- synthetic inputs
 * Scheidegger network: per cell a counter-based RNG keyed by (seed, linear index) draws
 * internal code 0 (E) / 1 (SE) / 2 (S) with p = 1/4, 1/2, 1/4 (keypad: 6 / 3 / 2).
 * (row0, col0) place the [ny][nx] window inside a raster nx_total columns wide, so stripes and
 * oracle sub-windows of one large raster agree cell by cell. */
int wmf_synth_scheidegger(int8_t *dir, int64_t ny, int64_t nx, int64_t row0, int64_t col0, int64_t nx_total,
                          uint64_t seed, int encoding, void *stream);
/* same generator on the host, for oracles and CPU baselines (host pointer). */
void wmf_synth_scheidegger_host(int8_t *dir_host, int64_t ny, int64_t nx, int64_t row0, int64_t col0,
                                int64_t nx_total, uint64_t seed, int encoding);

#ifdef __cplusplus
}
#endif
#endif /* WMF_B200_H */
