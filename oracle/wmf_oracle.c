/*
 * wmf_oracle.c -- CPU restatement of the reference's raster-drainage hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product package (wmf_heavy_b200/)
 * may import, link or call this file; only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs use it, as the checker.
 *
 * Two semantic families are restated, each function citing the reference lines
 * (paths relative to the reference checkout, uranzea/WMF_heavy):
 *
 *   (A) wmf_py (pure Python, executable reference)
 *         orc_basin_acum      <- wmf_py/cu_py/basics.py:124-177
 *         orc_basin_cut_mask  <- wmf_py/cu_py/basics.py:206-248
 *         orc_shia3           <- wmf_py/models_py/core.py:64-295
 *       PINNED: against the reference's own golden vectors
 *       (wmf_py/tests/test_cu_basics.py, test_models_core.py) and against
 *       fixtures produced by importing the Python reference
 *       (tests/golden/make_golden.py -> tests/golden/ fixtures).
 *
 *   (B) Fortran modules `cu` / `models` (f2py), NOT compilable in this image
 *       (no gfortran/flang):
 *         orc_f_coord2fil_col <- wmf_heavy/cuencas_heavy.f90:355-365
 *         orc_f_basin_find    <- wmf_heavy/cuencas_heavy.f90:507-576 (find+cut)
 *         orc_f_basin_basics  <- wmf_heavy/cuencas_heavy.f90:581-627
 *         orc_f_basin_coordxy <- wmf_heavy/cuencas_heavy.f90:639-646
 *         orc_f_stream_nod    <- wmf_heavy/cuencas_heavy.f90:780-809
 *         orc_f_calc_speed    <- wmf_heavy/Modelos_heavy.f90:907-919
 *         orc_f_shia_v1       <- wmf_heavy/Modelos_heavy.f90:169-548
 *       PARITY UNPINNED: the reference holds no test, fixture or golden vector
 *       for these routines, and the Fortran cannot be built here.  This
 *       restatement (fp32, Fortran evaluation order, compiled with
 *       -ffp-contract=off) is the arbiter; the hand-derived 3x3 example of
 *       SURVEY.md 8(c) is its only external anchor.
 *
 * Array conventions: rasters are C-order (row, col) for family (A) and
 * Fortran-order DIR(col, fil) == C-order [fil][col] for family (B) (the same
 * bytes as the GIS raster, wmfv2.py:246 passes Mapa.T to Fortran).  Per-cell
 * tables of family (B) with Fortran shape (k, N) are taken here as C-order
 * [k][N] (SoA); Fortran's memory order only matters for the order of the
 * fp32 whole-array sums, which is reproduced explicitly.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ (A) */

/* basics.py:37-46 */
static const int D8_DR[8] = {0, 1, 1, 1, 0, -1, -1, -1};
static const int D8_DC[8] = {1, 1, 0, -1, -1, -1, 0, 1};

/* basics.py:124-177.  flowdir as int32 (any value outside 0..7 has no
 * receiver, basics.py:151-153), mask as bytes (nonzero = True). */
int orc_basin_acum(const int32_t *fd, const uint8_t *mask, double *acc,
                   int64_t ny, int64_t nx)
{
    int64_t n = ny * nx;
    if (n == 0) return 0;
    int64_t *recv = (int64_t *)malloc(sizeof(int64_t) * n);
    int32_t *indeg = (int32_t *)calloc(n, sizeof(int32_t));
    int64_t *queue = (int64_t *)malloc(sizeof(int64_t) * n);
    if (!recv || !indeg || !queue) { free(recv); free(indeg); free(queue); return -1; }
    for (int64_t i = 0; i < n; i++) { acc[i] = 0.0; recv[i] = -1; }
    /* loop 1, basics.py:147-157 */
    for (int64_t r = 0; r < ny; r++)
        for (int64_t c = 0; c < nx; c++) {
            int64_t i = r * nx + c;
            if (!mask[i]) continue;
            int32_t d = fd[i];
            if (d < 0 || d > 7) continue;
            int64_t rr = r + D8_DR[d], cc = c + D8_DC[d];
            if (rr >= 0 && rr < ny && cc >= 0 && cc < nx && mask[rr * nx + cc]) {
                recv[i] = rr * nx + cc;
                indeg[rr * nx + cc] += 1;
            }
        }
    /* loop 2, basics.py:160-164 */
    int64_t head = 0, tail = 0;
    for (int64_t i = 0; i < n; i++)
        if (mask[i] && indeg[i] == 0) { acc[i] = 1.0; queue[tail++] = i; }
    /* loop 3, basics.py:166-175 */
    while (head < tail) {
        int64_t i = queue[head++];
        int64_t j = recv[i];
        if (j < 0) continue;
        acc[j] += acc[i];
        indeg[j] -= 1;
        if (indeg[j] == 0) { acc[j] += 1.0; queue[tail++] = j; }
    }
    free(recv); free(indeg); free(queue);
    return 0;
}

/* basics.py:206-248 (mask part).  Returns 0, or 1 "outlet outside grid",
 * 2 "outlet not in mask" (the two ValueErrors, basics.py:225-228). */
int orc_basin_cut_mask(const int32_t *fd, const uint8_t *mask, int64_t r0, int64_t c0,
                       uint8_t *mask_basin, int64_t ny, int64_t nx)
{
    int64_t n = ny * nx;
    memset(mask_basin, 0, (size_t)n);
    if (!(r0 >= 0 && r0 < ny && c0 >= 0 && c0 < nx)) return 1;
    if (!mask[r0 * nx + c0]) return 2;
    int64_t *queue = (int64_t *)malloc(sizeof(int64_t) * n);
    if (!queue) return -1;
    int64_t head = 0, tail = 0;
    queue[tail++] = r0 * nx + c0;
    mask_basin[r0 * nx + c0] = 1;
    while (head < tail) {
        int64_t i = queue[head++];
        int64_t r = i / nx, c = i % nx;
        for (int d = 0; d < 8; d++) {
            int64_t rr = r + D8_DR[d], cc = c + D8_DC[d];
            if (!(rr >= 0 && rr < ny && cc >= 0 && cc < nx)) continue;
            int64_t j = rr * nx + cc;
            if (!mask[j] || mask_basin[j]) continue;
            if (fd[j] == (d + 4) % 8) { mask_basin[j] = 1; queue[tail++] = j; }
        }
    }
    free(queue);
    return 0;
}

/* core.py:64-295.  Scalar parameters as in ModelParams (core.py:24-45).
 * rain/et: [nsteps][ncell] fp64 (et nullable; et_scale = dt/3600 for et_mm_h,
 * dt/86400 for et_mm_d, core.py:154-158).  cap/grav/aqu in/out [ncell].
 * q_super_last/q_base_last out [ncell].  Optional series (nullable)
 * [nsteps][ncell].  Sums are plain pairwise (numpy's np.sum is pairwise; the
 * parity tolerance on q_outlet, 1e-10 relative, covers the order).
 * Returns 0, or 3 when the cumulative mass error exceeds 1 % (core.py:255-256,
 * RuntimeError) -- in that case *fail_step holds the step index. */
typedef struct {
    double dt, max_capilar, max_gravita, max_aquifer, drena, retorno_gr;
    double k_perc_cap_to_grav, k_perc_grav_to_aqu, k_baseflow, max_infil_mm_h;
    double cell_area; /* dx*dy, or <0 when the grid has no dx,dy (core.py:113-115) */
    int32_t proportional; /* eta_priority == "proportional" */
    int32_t pad;
} orc_shia3_params;

static double pairwise_sum(const double *a, int64_t n)
{
    if (n <= 8) { double s = 0.0; for (int64_t i = 0; i < n; i++) s += a[i]; return s; }
    int64_t h = n / 2;
    return pairwise_sum(a, h) + pairwise_sum(a + h, n - h);
}

int orc_shia3(const orc_shia3_params *p, const double *rain, const double *et, double et_scale,
              int64_t nsteps, int64_t ncell, double *cap, double *grav, double *aqu,
              double *q_super_last, double *q_base_last, double *q_outlet,
              double *mass_err_step, double *mass_err_cum,
              double *q_super_t, double *q_base_t, double *q_total_t,
              double *cap_t, double *grav_t, double *aqu_t, int64_t *fail_step)
{
    double dt_h = p->dt / 3600.0;
    double *tmp = (double *)malloc(sizeof(double) * ncell * 6);
    if (!tmp) return -1;
    double *t_sto = tmp, *t_p = tmp + ncell, *t_et = tmp + 2 * ncell,
           *t_qs = tmp + 3 * ncell, *t_qb = tmp + 4 * ncell, *t_qt = tmp + 5 * ncell;
    for (int64_t i = 0; i < ncell; i++) t_sto[i] = cap[i] + grav[i] + aqu[i];
    double prev_sum = pairwise_sum(t_sto, ncell);          /* core.py:147 */
    double cum_p = 0.0, cum_err = 0.0;
    int rc = 0;
    for (int64_t t = 0; t < nsteps; t++) {
        const double *rn = rain + t * ncell;
        const double *ev = et ? et + t * ncell : NULL;
        for (int64_t i = 0; i < ncell; i++) {
            double rain_step = fmax(rn[i], 0.0) * dt_h;                 /* :152 */
            double et_step = ev ? fmax(ev[i], 0.0) * et_scale : 0.0;    /* :154-158 */
            double infil = fmin(rain_step, p->max_infil_mm_h * dt_h);   /* :160 */
            double exceso = rain_step - infil;
            double c = cap[i] + infil, g = grav[i], a = aqu[i];
            double et_c, et_g, et_a;
            if (p->proportional) {                                      /* :165-176 */
                double w = c + g + a;
                double fc = 0.0, fg = 0.0, fa = 0.0;
                if (w > 0) { fc = c / w; fg = g / w; fa = a / w; }
                et_c = fmin(et_step * fc, c);
                et_g = fmin(et_step * fg, g);
                et_a = fmin(et_step * fa, a);
            } else {                                                    /* :177-182 */
                et_c = fmin(et_step, c);
                double rem = et_step - et_c;
                et_g = fmin(rem, g);
                double rem2 = rem - et_g;
                et_a = fmin(rem2, a);
            }
            c -= et_c; g -= et_g; a -= et_a;
            double et_total = (et_c + et_g) + et_a;                     /* :187 */
            double perc_cg = fmin(p->k_perc_cap_to_grav * c, c);        /* :189-193 */
            c -= perc_cg; g += perc_cg;
            double perc_ga = fmin(p->k_perc_grav_to_aqu * g, g);        /* :195-200 */
            double retorno = p->retorno_gr * g;
            g = (g - perc_ga) - retorno;
            a += perc_ga;
            double q_base = fmin(p->k_baseflow * a, a);                 /* :202-204 */
            double drena = p->drena * a;
            a = (a - q_base) + drena;
            double overflow = fmax(c - p->max_capilar, 0.0);            /* :206-208 */
            c -= overflow; exceso += overflow;
            c = fmin(fmax(c, 0.0), p->max_capilar);                     /* :210-212 */
            g = fmin(fmax(g, 0.0), p->max_gravita);
            a = fmin(fmax(a, 0.0), p->max_aquifer);
            double q_super = exceso + retorno;                          /* :214-215 */
            double q_total = q_super + q_base;
            cap[i] = c; grav[i] = g; aqu[i] = a;
            q_super_last[i] = q_super; q_base_last[i] = q_base;
            t_sto[i] = (c + g) + a; t_p[i] = rain_step; t_et[i] = et_total;
            t_qs[i] = q_super; t_qb[i] = q_base; t_qt[i] = q_total;
            if (q_super_t) { q_super_t[t * ncell + i] = q_super; q_base_t[t * ncell + i] = q_base;
                             q_total_t[t * ncell + i] = q_total; }
            if (cap_t) { cap_t[t * ncell + i] = c; grav_t[t * ncell + i] = g; aqu_t[t * ncell + i] = a; }
        }
        double q_total_sum = pairwise_sum(t_qt, ncell);                 /* :234-238 */
        q_outlet[t] = (p->cell_area < 0) ? q_total_sum
                                         : q_total_sum * 1.0e-3 * p->cell_area / p->dt;
        double new_sum = pairwise_sum(t_sto, ncell);                    /* :240-242 */
        double deltaS = new_sum - prev_sum;
        prev_sum = new_sum;
        double P = pairwise_sum(t_p, ncell), ETs = pairwise_sum(t_et, ncell);
        double Qs = pairwise_sum(t_qs, ncell), Qb = pairwise_sum(t_qb, ncell);
        double err = P - (((deltaS + Qs) + Qb) + ETs);                  /* :248 */
        cum_p += P; cum_err += err;
        double rel_step = P > 0 ? fabs(err) / P : 0.0;
        double rel_cum = cum_p > 0 ? fabs(cum_err) / cum_p : 0.0;
        mass_err_step[t] = rel_step; mass_err_cum[t] = rel_cum;
        if (rel_cum > 0.01) { rc = 3; if (fail_step) *fail_step = t; break; }
    }
    free(tmp);
    return rc;
}

/* ------------------------------------------------------------------ (B) */

/* cuencas_heavy.f90:355-365, all arithmetic in fp32 (`real`); note dx is used
 * for the row as well.  Returns -999 for out-of-range components. */
void orc_f_coord2fil_col(float x, float y, float xll, float yll, float dx,
                         int32_t ncols, int32_t nrows, int32_t *col, int32_t *fil)
{
    float qx = (x - xll) / dx, qy = (y - yll) / dx;
    int32_t c = (int32_t)ceilf(qx);
    if (c < 0) c = -c;
    int32_t f = nrows - (int32_t)floorf(qy);
    if (c > ncols || c <= 0) c = -999;
    if (f > nrows || f <= 0) f = -999;
    *col = c; *fil = f;
}

/* cuencas_heavy.f90:507-576 (basin_find + basin_cut).  DIR is keypad-coded
 * (7 8 9 / 4 5 6 / 1 2 3), dir[(fil-1)*ncols + (col-1)].  Output
 * structure[3][n] C-order == the transpose-free image of basin_f(3, n):
 * structure[0][i] = drain id, [1][i] = col, [2][i] = fil (all 1-based).
 * Defined deviations (SURVEY 8(a) D3 caveats): up to 8 donors per cell are
 * supported (the Fortran scratch res(2,7) would overflow), and a cell coded 5
 * inside the basin (which would re-enqueue itself forever) returns -5.
 * `cap` = capacity of structure in cells; returns n, or <0 on error
 * (-2 outlet out of range, -3 capacity). */
int64_t orc_f_basin_find(const int32_t *dir, int32_t ncols, int32_t nrows,
                         int32_t coli, int32_t fili, int32_t *structure, int64_t cap)
{
    if (coli < 1 || coli > ncols || fili < 1 || fili > nrows) return -2;
    int64_t celTot = (int64_t)ncols * nrows;
    /* queue in discovery order: q[k] = k-th discovered cell (k=0 the outlet);
     * Fortran stores it reversed at basin_temp(:, celTot-k). */
    int32_t *qc = (int32_t *)malloc(sizeof(int32_t) * celTot);
    int32_t *qf = (int32_t *)malloc(sizeof(int32_t) * celTot);
    int32_t *qd = (int32_t *)malloc(sizeof(int32_t) * celTot);
    if (!qc || !qf || !qd) { free(qc); free(qf); free(qd); return -1; }
    int64_t tenia = 1, cont2 = 1;
    qc[0] = coli; qf[0] = fili; qd[0] = 0;
    if (dir[(int64_t)(fili - 1) * ncols + (coli - 1)] == 5) { free(qc); free(qf); free(qd); return -5; }
    while (cont2 <= tenia) {                 /* `do while (col2 > 0)`: stops at the -1 fill */
        int32_t col2 = qc[cont2 - 1], row2 = qf[cont2 - 1];
        for (int kf = 1; kf <= 3; kf++)
            for (int kc = 1; kc <= 3; kc++) {
                int32_t c = col2 + kc - 2, f = row2 + kf - 2;
                if (c >= 1 && c <= ncols && f >= 1 && f <= nrows) {
                    int32_t code = dir[(int64_t)(f - 1) * ncols + (c - 1)];
                    if (code == 3 * kf - kc + 1) {
                        if (kf == 2 && kc == 2) { free(qc); free(qf); free(qd); return -5; }
                        if (tenia >= celTot) { free(qc); free(qf); free(qd); return -4; }
                        qc[tenia] = c; qf[tenia] = f; qd[tenia] = (int32_t)cont2;
                        tenia++;
                    }
                }
            }
        cont2++;
    }
    int64_t n = tenia;
    if (n > cap) { free(qc); free(qf); free(qd); return -3; }
    for (int64_t k = 0; k < n; k++) {        /* basin_f(:, i) with i = n - k (1-based) */
        int64_t i0 = n - 1 - k;
        structure[0 * n + i0] = qd[k];
        structure[1 * n + i0] = qc[k];
        structure[2 * n + i0] = qf[k];
    }
    free(qc); free(qf); free(qd);
    return n;
}

/* cuencas_heavy.f90:639-646, fp32. */
void orc_f_basin_coordxy(const int32_t *structure, int64_t n, float xll, float yll, float dx,
                         int32_t nrows, float *X, float *Y)
{
    for (int64_t i = 0; i < n; i++) {
        X[i] = xll + dx * ((float)structure[1 * n + i] - 0.5f);
        Y[i] = yll + dx * ((float)(nrows - structure[2 * n + i]) + 0.5f);
    }
}

static int cmp_float(const void *a, const void *b)
{
    float x = *(const float *)a, y = *(const float *)b;
    return (x > y) - (x < y);
}

/* cuencas_heavy.f90:581-627.  scalars[6] = area, pend_media, elevacion,
 * centroX, centroY, (unused).  fp32 sequential sums (:613-615); centroid =
 * median of sorted X / Y (:617-626).  pend(outlet) copies pend(i-1) (:607);
 * for n == 1 that reads pend(0), out of bounds in Fortran -- defined here as 0
 * before the `== 0 -> 0.001` rule. */
void orc_f_basin_basics(const int32_t *structure, const float *dem, const int32_t *dirv, int64_t n,
                        float dxp, float xll, float yll, float dx, int32_t nrows,
                        int32_t *acum, float *long_cel, float *pend, float *elev, float *scalars)
{
    for (int64_t i = 0; i < n; i++) { elev[i] = dem[i]; acum[i] = 1; }
    for (int64_t i = 0; i < n; i++) {
        int32_t dr = structure[i];
        int64_t drenaid = n - dr;                       /* 0-based index of n - dr + 1 */
        if (dirv[i] % 2 == 0) long_cel[i] = dxp;
        else long_cel[i] = dxp * sqrtf(2.0f);
        if (dr != 0) {
            acum[drenaid] += acum[i];
            pend[i] = fabsf(dem[i] - dem[drenaid]) / long_cel[i];
        } else {
            pend[i] = (i > 0) ? pend[i - 1] : 0.0f;
        }
        if (pend[i] == 0.0f) pend[i] = 0.001f;
    }
    float sp = 0.0f, se = 0.0f;
    for (int64_t i = 0; i < n; i++) sp += pend[i];
    for (int64_t i = 0; i < n; i++) se += dem[i];
    scalars[0] = (float)n * (dxp * dxp) / 1e6f;
    scalars[1] = sp / (float)n;
    scalars[2] = se / (float)n;
    float *X = (float *)malloc(sizeof(float) * n), *Y = (float *)malloc(sizeof(float) * n);
    orc_f_basin_coordxy(structure, n, xll, yll, dx, nrows, X, Y);
    qsort(X, n, sizeof(float), cmp_float);
    qsort(Y, n, sizeof(float), cmp_float);
    int64_t m = (n % 2 == 0) ? n / 2 : (n + 1) / 2;     /* 1-based */
    scalars[3] = X[m - 1]; scalars[4] = Y[m - 1]; scalars[5] = 0.0f;
    free(X); free(Y);
}

/* acum only (cu.basin_acum(structure, ncells), wmfv2.py:1868 == the acum of
 * basin_basics, cuencas_heavy.f90:593-606). */
void orc_f_basin_acum(const int32_t *structure, int64_t n, int32_t *acum)
{
    for (int64_t i = 0; i < n; i++) acum[i] = 1;
    for (int64_t i = 0; i < n; i++) {
        int32_t dr = structure[i];
        if (dr != 0) acum[n - dr] += acum[i];
    }
}

/* cuencas_heavy.f90:780-809.  counts[0] = n_nodos, counts[1] = n_cauce. */
void orc_f_stream_nod(const int32_t *structure, const int32_t *acum, int64_t n, int32_t umbral,
                      int32_t *cauce, int32_t *nodos, int32_t *trazado, int32_t *counts)
{
    int32_t n_cauce = 0, n_nodos = 0;
    for (int64_t i = 0; i < n; i++) { cauce[i] = acum[i] >= umbral ? 1 : 0; n_cauce += cauce[i]; nodos[i] = 0; }
    for (int64_t i = 0; i < n; i++) {
        int32_t dr = structure[i];
        if (dr != 0 && cauce[i] == 1) nodos[n - dr] += 1;
    }
    for (int64_t i = 0; i < n; i++) if (nodos[i] == 0 && cauce[i] == 1) nodos[i] = 3;
    nodos[n - 1] = 2;
    for (int64_t i = 0; i < n; i++) { if (nodos[i] < 2) nodos[i] = 0; if (nodos[i] > 0) n_nodos++; }
    for (int64_t i = 0; i < n; i++) trazado[i] = 0;
    trazado[n - 1] = 1;
    for (int64_t i = 0; i < n; i++) {
        int32_t dr = structure[i];
        if (dr != 0 && nodos[n - dr] != 0 && cauce[i] == 1) trazado[i] = 1;
    }
    counts[0] = n_nodos; counts[1] = n_cauce;
}

/* Modelos_heavy.f90:907-919, fp32. */
void orc_f_calc_speed(float sm, float coef, float expo, float elem_long, float dt, int niter,
                      float *speed, float *area)
{
    float sp = *speed, ar = *area;
    for (int i = 0; i < niter; i++) {
        ar = sm / (elem_long + sp * dt);
        float nw = coef * powf(ar, expo);
        sp = (2.0f * nw + sp) / 3.0f;
    }
    *speed = sp; *area = ar;
}

/* Modelos_heavy.f90:169-548 as shipped: vertical 5-tank update + lateral
 * outflow of tanks 2..4, outflow discarded, tank 5 untouched, Q == 0.
 * Layout: v_coef/h_coef/h_exp [4][N]; max_capilar/max_gravita/max_aquifer/
 * hill_long/elem_area [N]; sto [5][N] in/out (already initialised by the
 * caller: StoIn or Storage+storage_constant, :229-233); hspeed [4][N] in/out;
 * hspeed_given != 0 mirrors `HspeedIn(1,1) > 0` (:239-240).  vspeed is always
 * computed (the Fortran leaves it uninitialised in the HspeedIn branch; defined
 * here).  speed_type has 3 entries; row 4 of hspeed is initialised as type 1
 * (the Fortran reads speed_type(4), out of bounds; defined here).
 * rain_records [n_rec][N] int32 (record r at row r-1), pos_evento [N_reg]
 * 1-based, ==1 meaning "no rain" without reading (:366-367).
 * Outputs: balance [N_reg], mean_rain [N_reg], acum_rain [N],
 * mean_storage [5][N_reg] (nullable == show_storage off),
 * mean_speed [4][N_reg] (nullable), retorned_mean [N_reg] (nullable).
 * balance64 [N_reg] (nullable) is NOT a reference output: the same balance terms (the very fp32
 * per-cell values) summed in fp64, i.e. the quantity the Fortran's fp32 running sums approximate.
 * `balance` is a difference of two ~N*5-term fp32 sums (catastrophic cancellation), so tests pin the
 * CUDA path against balance64 tightly and against `balance` only within that fp32 summation noise. */
typedef struct {
    float dt, retorno_gr, retorno_aq;
    int32_t calc_niter;
    int32_t speed_type[3];
    int32_t hspeed_given;
} orc_shia5_params;

int orc_f_shia_v1(const orc_shia5_params *p, int64_t N, int64_t N_reg, const float *calib,
                  const float *v_coef, const float *h_coef, const float *h_exp,
                  const float *max_capilar, const float *max_gravita, const float *max_aquifer,
                  const float *hill_long, const float *elem_area,
                  const int32_t *rain_records, const int32_t *pos_evento, const float *evp,
                  float *sto, float *hspeed, float *balance, float *mean_rain, float *acum_rain,
                  float *mean_storage, float *mean_speed, float *mean_retorno, double *balance64)
{
    float dt = p->dt;
    float *vspeed = (float *)malloc(sizeof(float) * 4 * N);
    float *H = (float *)malloc(sizeof(float) * 3 * N);
    float *rain = (float *)malloc(sizeof(float) * N);
    float *retorned = (float *)calloc(N, sizeof(float));
    if (!vspeed || !H || !rain || !retorned) return -1;
    for (int i = 0; i < 4; i++)                                          /* :242-249 */
        for (int64_t c = 0; c < N; c++) {
            vspeed[i * N + c] = v_coef[i * N + c] * calib[i] * dt;
            if (!p->hspeed_given) {
                int st = i < 3 ? p->speed_type[i] : 1;
                hspeed[i * N + c] = (st == 1) ? h_coef[i * N + c] * calib[i + 4] : 0.0f;
            }
        }
    for (int64_t c = 0; c < N; c++) {                                    /* :252-254 */
        H[0 * N + c] = max_capilar[c] * calib[8];
        H[1 * N + c] = max_gravita[c] * calib[9];
        H[2 * N + c] = max_aquifer[c] * calib[10];
        acum_rain[c] = 0.0f;
    }
    float entradas = 0.0f, salidas = 0.0f;
    for (int64_t t = 0; t < N_reg; t++) {
        float sto_atras = 0.0f;                                          /* :364, Fortran memory order */
        double d_atras = 0.0, d_ent = 0.0, d_sal = 0.0, d_now = 0.0;
        for (int64_t c = 0; c < N; c++) for (int k = 0; k < 5; k++) { sto_atras += sto[k * N + c]; d_atras += sto[k * N + c]; }
        if (pos_evento[t] == 1) for (int64_t c = 0; c < N; c++) rain[c] = 0.0f;
        else {
            const int32_t *rec = rain_records + (int64_t)(pos_evento[t] - 1) * N;
            for (int64_t c = 0; c < N; c++) rain[c] = (float)rec[c] / 1000.0f;   /* :370 */
        }
        float rain_sum = 0.0f;
        for (int64_t c = 0; c < N; c++) acum_rain[c] = acum_rain[c] + rain[c];
        for (int64_t c = 0; c < N; c++) {
            float R = rain[c];
            float S1 = sto[0 * N + c], S2 = sto[1 * N + c], S3 = sto[2 * N + c], S4 = sto[3 * N + c];
            float H1 = H[0 * N + c];
            float vflux[4], S[5];
            entradas = entradas + R; rain_sum = rain_sum + R; d_ent += R;
            vflux[0] = fmaxf(0.0f, (R - H1) + S1);                       /* :393 */
            S1 = (S1 + R) - vflux[0];                                    /* :394 */
            float evp_loss = fminf((evp[t] * vspeed[0 * N + c]) * powf(S1 / H1, 0.6f), S1); /* :396-398 */
            S1 = S1 - evp_loss;                                          /* :407 */
            S[0] = S1; S[1] = S2; S[2] = S3; S[3] = S4;
            for (int i = 0; i < 3; i++) {                                /* :409-412 */
                vflux[i + 1] = fminf(vflux[i], vspeed[(i + 1) * N + c]);
                S[i + 1] = (S[i + 1] + vflux[i]) - vflux[i + 1];
            }
            if (p->retorno_aq > 0) {                                     /* :414-418 */
                float ra = fmaxf(0.0f, S[3] - H[2 * N + c]);
                S[2] = S[2] + ra; S[3] = S[3] - ra;
            }
            if (p->retorno_gr > 0) {                                     /* :420-427 */
                float rg = fmaxf(0.0f, S[2] - H[1 * N + c]);
                S[1] = S[1] + rg; S[2] = S[2] - rg;
                retorned[c] = retorned[c] + rg;
                vflux[1] = vflux[1] + rg; vflux[2] = vflux[2] - rg;
            }
            salidas = (salidas + vflux[3]) + evp_loss;                   /* :459 */
            d_sal += (double)(vflux[3] + evp_loss);
            float L = hill_long[c], m3mm = elem_area[c] / 1000.0f;       /* :224 */
            for (int i = 0; i < 3; i++) {                                /* :461-491 */
                float hf;
                if (p->speed_type[i] == 1) {
                    hf = (1.0f - L / (hspeed[i * N + c] * dt + L)) * S[i + 1];
                } else {
                    float sp = hspeed[i * N + c], area = 0.0f;
                    orc_f_calc_speed(S[i + 1] * m3mm, h_coef[i * N + c] * calib[i + 4],
                                     h_exp[i * N + c], L, dt, p->calc_niter, &sp, &area);
                    hspeed[i * N + c] = sp;
                    hf = fminf(((area * sp) * dt) / m3mm, S[i + 1]);
                }
                S[i + 1] = S[i + 1] - hf;
            }
            sto[0 * N + c] = S[0]; sto[1 * N + c] = S[1]; sto[2 * N + c] = S[2]; sto[3 * N + c] = S[3];
        }
        mean_rain[t] = rain_sum / (float)N;                              /* :494 */
        if (mean_storage)                                                /* :517-519 */
            for (int k = 0; k < 5; k++) {
                float s = 0.0f; for (int64_t c = 0; c < N; c++) s += sto[k * N + c];
                mean_storage[k * N_reg + t] = s / (float)N;
            }
        if (mean_speed)                                                  /* :521-523 */
            for (int k = 0; k < 4; k++) {
                float s = 0.0f; for (int64_t c = 0; c < N; c++) s += hspeed[k * N + c];
                mean_speed[k * N_reg + t] = s / (float)N;
            }
        if (mean_retorno) {                                              /* :525-531 */
            float s = 0.0f; for (int64_t c = 0; c < N; c++) s += retorned[c];
            mean_retorno[t] = s / (float)N;
            for (int64_t c = 0; c < N; c++) retorned[c] = 0.0f;
        }
        float sto_now = 0.0f;                                            /* :533 */
        for (int64_t c = 0; c < N; c++) for (int k = 0; k < 5; k++) { sto_now += sto[k * N + c]; d_now += sto[k * N + c]; }
        balance[t] = ((sto_now - sto_atras) - entradas) + salidas;
        if (balance64) balance64[t] = ((d_now - d_atras) - d_ent) + d_sal;
        entradas = 0.0f; salidas = 0.0f;
    }
    free(vspeed); free(H); free(rain); free(retorned);
    return 0;
}

/* ------------------------------------------------------------------ N4 (extension, no reference behaviour)
 * Channel routing behind the shipped 5-tank update.  The reference declares tank 5, `drena` and the control
 * points but never moves water between cells (Q == 0 as shipped); this is the documented extension of
 * per step, after the shipped vertical / lateral update of a cell (speed type 1 only),
 *     lat   = (hflux2 + hflux3) + hflux4                         the lateral outflows the shipped code drops [mm]
 *     S5    = (S5 + lat) + (sum over donors d, ascending index, of qout_m3(d, t)) / m3mm(c)
 *     q5    = (1 - L / (hspeed4 * dt + L)) * S5 ;  S5 = S5 - q5     linear reservoir, like tanks 2-4
 *     qout_m3(c, t) = q5 * m3mm(c)            goes to the receiver n - drain(c) within the same step
 *     Q(j, t) = qout_m3 / dt  at the cells ctrl marks (row j), row 0 = the outlet (the last cell)
 * Cells are in the structure's upstream-first order, so index order is a topological order.  PARITY UNPINNED by
 * construction.  totals[3] (fp64) = rain volume, losses (v4 + evaporation), outflow at the outlet, all m3. */
int orc_f_shia_route(const orc_shia5_params *p, int64_t N, int64_t N_reg, const float *calib,
                     const float *v_coef, const float *h_coef,
                     const float *max_capilar, const float *max_gravita, const float *max_aquifer,
                     const float *hill_long, const float *elem_area,
                     const int32_t *rain_records, const int32_t *pos_evento, const float *evp,
                     const int32_t *drain, const int32_t *ctrl, int32_t n_ctrl,
                     float *sto, float *acum_rain, float *Q, double *totals)
{
    float dt = p->dt;
    float *qout = (float *)calloc(N, sizeof(float));
    float *qin = (float *)calloc(N, sizeof(float));
    if (!qout || !qin) return -1;
    for (int64_t c = 0; c < N; c++) acum_rain[c] = 0.0f;
    for (int64_t i = 0; i < (int64_t)n_ctrl * N_reg; i++) Q[i] = 0.0f;
    totals[0] = totals[1] = totals[2] = 0.0;
    for (int64_t t = 0; t < N_reg; t++) {
        for (int64_t c = 0; c < N; c++) qin[c] = 0.0f;
        for (int64_t c = 0; c < N; c++) {
            float R = 0.0f;
            if (pos_evento[t] != 1) R = (float)rain_records[(int64_t)(pos_evento[t] - 1) * N + c] / 1000.0f;
            acum_rain[c] = acum_rain[c] + R;
            float vs[4], hs[4];
            for (int i = 0; i < 4; i++) { vs[i] = v_coef[i * N + c] * calib[i] * dt; hs[i] = h_coef[i * N + c] * calib[i + 4]; }
            float H1 = max_capilar[c] * calib[8], H2 = max_gravita[c] * calib[9], H3 = max_aquifer[c] * calib[10];
            float S1 = sto[0 * N + c], S2 = sto[1 * N + c], S3 = sto[2 * N + c], S4 = sto[3 * N + c], S5 = sto[4 * N + c];
            float v1 = fmaxf(0.0f, (R - H1) + S1);
            S1 = (S1 + R) - v1;
            float evl = fminf((evp[t] * vs[0]) * powf(S1 / H1, 0.6f), S1);
            S1 = S1 - evl;
            float v2 = fminf(v1, vs[1]); S2 = (S2 + v1) - v2;
            float v3 = fminf(v2, vs[2]); S3 = (S3 + v2) - v3;
            float v4 = fminf(v3, vs[3]); S4 = (S4 + v3) - v4;
            if (p->retorno_aq > 0) { float ra = fmaxf(0.0f, S4 - H3); S3 = S3 + ra; S4 = S4 - ra; }
            if (p->retorno_gr > 0) { float rg = fmaxf(0.0f, S3 - H2); S2 = S2 + rg; S3 = S3 - rg; }
            float L = hill_long[c], m3mm = elem_area[c] / 1000.0f;
            float hf2 = (1.0f - L / (hs[0] * dt + L)) * S2; S2 = S2 - hf2;
            float hf3 = (1.0f - L / (hs[1] * dt + L)) * S3; S3 = S3 - hf3;
            float hf4 = (1.0f - L / (hs[2] * dt + L)) * S4; S4 = S4 - hf4;
            float lat = (hf2 + hf3) + hf4;
            S5 = (S5 + lat) + qin[c] / m3mm;
            float q5 = (1.0f - L / (hs[3] * dt + L)) * S5;
            S5 = S5 - q5;
            float qm3 = q5 * m3mm;
            qout[c] = qm3;
            int32_t dr = drain[c];
            if (dr != 0) qin[N - dr] = qin[N - dr] + qm3;            /* donors arrive in ascending index order */
            else { Q[t] = qm3 / dt; totals[2] += (double)qm3; }      /* row 0: the outlet */
            if (ctrl[c] > 0) Q[(int64_t)ctrl[c] * N_reg + t] = qm3 / dt;
            totals[0] += (double)R * (double)m3mm;
            totals[1] += ((double)v4 + (double)evl) * (double)m3mm;
            sto[0 * N + c] = S1; sto[1 * N + c] = S2; sto[2 * N + c] = S3; sto[3 * N + c] = S4; sto[4 * N + c] = S5;
        }
    }
    free(qout); free(qin);
    return 0;
}

/* ---------------------------------------------------------------------------------------------
 * Synthetic workload of SURVEY 8(d) on the host, so that the CPU-baseline legs of bench.py need nothing but this
 * library: the Scheidegger DIR raster of the headline benchmark.  Per cell a counter-based generator keyed by
 * (seed, global linear index) -- the splitmix64 finaliser, top two bits: 0 -> E, 1 or 2 -> SE, 3 -> S -- exactly the
 * generator the device uses (wmf_synth_scheidegger), so any window of the raster can be produced on either side.
 * (tests/test_oracle.py checks the two against each other.) */
void orc_synth_scheidegger(int8_t *dir, int64_t ny, int64_t nx, int64_t row0, int64_t col0, int64_t nx_total,
                           uint64_t seed)
{
    for (int64_t r = 0; r < ny; r++)
        for (int64_t c = 0; c < nx; c++) {
            uint64_t x = seed ^ ((uint64_t)((row0 + r) * nx_total + col0 + c) * 0x9E3779B97F4A7C15ull);
            x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ull;
            x ^= x >> 27; x *= 0x94D049BB133111EBull;
            x ^= x >> 31;
            const int u = (int)(x >> 62);
            dir[r * nx + c] = (int8_t)(u == 0 ? 0 : (u == 3 ? 2 : 1));
        }
}
