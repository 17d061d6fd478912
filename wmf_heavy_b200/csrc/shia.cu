// shia.cu -- SHIA per-cell tank update, time loop fused in the kernel.
//
//   k_shia5  replaces models.shia_v1 as shipped (wmf_heavy/Modelos_heavy.f90:169-548) with
//            calc_speed (:907-919): fp32, Fortran evaluation order.  This translation unit is
//            compiled with -fmad=false so a*b+c rounds twice like the un-contracted Fortran.
//   k_shia3  replaces wmf_py/models_py/core.py:64-295: fp64, numpy evaluation order.
//
// As shipped, cells do not interact inside shia_v1 (the lateral outflow is computed, subtracted
// and dropped), so one thread owns one cell for all N_reg steps: state stays in
// registers and the only per-step HBM traffic is the 4-byte rain word (8 bytes for fp64).
// The per-step whole-basin sums the reference forms (balance, Mean_Rain, mean_storage,
// mean_speed; q_outlet and the mass-balance terms) are reduced deterministically:
// warp shuffle -> per-block partial every TB steps (fixed order, fp64) -> one pass over blocks.
#include "common.cuh"

namespace {

constexpr int TH = 256;        // threads per block
constexpr int WARPS = TH / 32;
constexpr int TB = 16;         // time steps batched between block-level reductions

template <typename V>
__device__ __forceinline__ V warp_sum(V v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Stages per-warp sums for TB steps in shared memory, then folds warps in fixed order and
// writes part[blk][t][k] (fp64).  Every thread of the block must call add()/flush() uniformly.
template <int K, typename V>
struct StepReducer {
    double (*stage)[TB][K];     // [WARPS][TB][K] in shared memory
    double *part;               // this block's [T1][K] slice in global memory
    int slot;
    int64_t t0;
    __device__ StepReducer(double (*s)[TB][K], double *p) : stage(s), part(p), slot(0), t0(0) {}
    __device__ __forceinline__ void add(const V (&v)[K])
    {
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
        for (int k = 0; k < K; k++) {
            V s = warp_sum<V>(v[k]);
            if (lane == 0) stage[warp][slot][k] = (double)s;
        }
        if (++slot == TB) flush();
    }
    // TB steps at once (K == 2): q[k][u] is this thread's value of quantity k at step u of the batch.  A transposed
    // butterfly leaves the warp sum of step u in lanes 2u and 2u + 1: 16 shuffles per quantity for 16 steps instead of
    // 5 per step.  The caller flushes.
    __device__ __forceinline__ void add_batch(float (&q)[2][TB], int nvalid)
    {
        static_assert(TB == 16, "the butterfly below reduces 16 steps over 32 lanes");
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
        for (int k = 0; k < 2; k++) {
            float (&a)[TB] = q[k];
#pragma unroll
            for (int h = TB / 2, o = 16; h >= 1; h >>= 1, o >>= 1) {
                const bool upper = (lane & o) != 0;          // this lane keeps the upper half of the indices
#pragma unroll
                for (int i = 0; i < h; i++) {
                    const float send = upper ? a[i] : a[i + h], keep = upper ? a[i + h] : a[i];
                    a[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
                }
            }
            const float tot = a[0] + __shfl_xor_sync(0xffffffffu, a[0], 1);
            const int u = lane >> 1;
            if ((lane & 1) == 0 && u < nvalid) stage[warp][u][k] = (double)tot;
        }
        slot = nvalid;
    }
    __device__ __forceinline__ void flush()
    {
        if (slot == 0) return;
        __syncthreads();
        for (int idx = threadIdx.x; idx < slot * K; idx += TH) {
            int s = idx / K, k = idx - s * K;
            double a = 0.0;
#pragma unroll
            for (int w = 0; w < WARPS; w++) a += stage[w][s][k];
            part[(t0 + s) * K + k] = a;
        }
        __syncthreads();
        t0 += slot;
        slot = 0;
    }
};

// totals[g][t][k] = sum over blocks of part[g][blk][t][k]   (g = parameter set), in a fixed order and in two
// steps so that the fold is a streaming pass, not TK threads walking nblk rows each: chunks of FOLD_CS consecutive
// blocks are summed in place into the chunk's first row (grid.z = chunks), then the chunk rows are summed.
constexpr int FOLD_CS = 32;
__global__ void k_fold_chunks(double *__restrict__ part, int nblk, int64_t TK)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= TK) return;
    const int b0 = blockIdx.z * FOLD_CS, b1 = b0 + FOLD_CS < nblk ? b0 + FOLD_CS : nblk;
    double *p = part + (int64_t)blockIdx.y * nblk * TK;
    double a = 0.0;
    for (int b = b0; b < b1; b++) a += p[(int64_t)b * TK + i];
    p[(int64_t)b0 * TK + i] = a;
}
__global__ void k_fold_blocks(const double *__restrict__ part, int nblk, int64_t TK, double *__restrict__ totals)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= TK) return;
    const double *p = part + (int64_t)blockIdx.y * nblk * TK;
    double a = 0.0;
    for (int b = 0; b < nblk; b += FOLD_CS) a += p[(int64_t)b * TK + i];
    totals[(int64_t)blockIdx.y * TK + i] = a;
}
static void fold_blocks(double *part, int nblk, int64_t TK, int P, double *totals, cudaStream_t st)
{
    k_fold_chunks<<<dim3(blocks_for(TK, 256), P, (nblk + FOLD_CS - 1) / FOLD_CS), 256, 0, st>>>(part, nblk, TK);
    k_fold_blocks<<<dim3(blocks_for(TK, 256), P), 256, 0, st>>>(part, nblk, TK, totals);
}

// ------------------------------------------------------------------------------------ SHIA 5-tank
// Modelos_heavy.f90:907-919.  FAST: area ** expo as exp2(expo * log2(area)) on the SFU (see pow06) instead of powf.
template <bool FAST>
__device__ __forceinline__ void calc_speed(float sm, float coef, float expo, float L, float dt, int niter, float &sp,
                                           float &area)
{
    for (int i = 0; i < niter; i++) {
        if (FAST) {
            // the two divisions of the iteration as an SFU reciprocal (1 ulp) and a multiply by 1/3: an IEEE division is
            // ~9 instructions, and the iteration -- a weighted average with the previous speed -- damps a rounding error
            float l, pw, r;
            asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(L + sp * dt));
            area = sm * r;
            asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l) : "f"(area));
            l = l * expo;
            asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(pw) : "f"(l));
            sp = (2.0f * (coef * pw) + sp) * 0.33333334f;
        } else {
            area = sm / (L + sp * dt);
            const float nw = coef * powf(area, expo);
            sp = (2.0f * nw + sp) / 3.0f;
        }
    }
}

struct Shia5Args {
    wmf_shia5_params p;
    int64_t N, N_reg;
    int32_t n_rec;              // records in `rain`: a posEvento outside 2..n_rec reads nothing (no rain)
    const float *cs;            // [17][N]
    const int32_t *rain;        // [n_rec][N]
    const int32_t *pos;         // [N_reg]
    const float *evp;           // [N_reg]
    const float *calib;         // [P][11]
    float *sto;                 // [P][5][N]
    float *hspeed;              // [P][4][N]
    float *acum_rain;           // [N]
    double *part;               // [P][nblk][N_reg+1][K]
    int nblk;
    float *vfluxes;             // [P][4][N] or null: the vertical fluxes of the last step (Modelos:429-433)
    float *rc_coef;             // [P][2][N] or null: runoff-coefficient accumulators (Modelos:435-438, 543-544)
    float *retorned;            // [P][N] or null: return flow accumulated per cell (Modelos:424)
    int retorned_per_step;      // Retorned is cleared after every step (show_mean_retorno / save_retorno, Modelos:525-531)
    const int32_t *snap_slot;   // [N_reg] or null: slot (>= 0) that receives the storages after that step
    float *snap_out;            // [P][n_snap][5][N]
    int n_snap;
};

// x ** 0.6 for x in [0, 1] (soil saturation): exp2(0.6 * log2 x) on the SFU.  __log2f is within 2^-22
// absolute on [0.5, 2] and 2 ulp elsewhere, so the relative error of the power is ~1e-7 -- two orders of
// magnitude inside the 1e-5 parity tolerance -- against ~70 instructions for the correctly rounded powf.
// The .ftz forms need no range fix-ups: lg2(0) = -inf and ex2(-inf) = 0, a denormal saturation counts as 0.
__device__ __forceinline__ float pow06(float x)
{
    float l, r;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l) : "f"(x));
    l = l * 0.6f;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(l));
    return r;
}

// Per-step sums.  K = 2: {R, q} with q = (storage now - storage before) + (v4 + Evp_loss) per cell, so that
// balance(t) = sum(q) - sum(R) (Modelos:533) needs ONE reduced quantity beside the rain; forming the
// difference per cell in fp32 also avoids the catastrophic cancellation of the Fortran's two ~5N-term sums.
// DIAG adds {S1,S2,S3,S4,h1,h2,h3} and {Ret, vflux1..4} -> K = 14 (slot 0 of the series carries the constant sums of S5
// and h4); it also keeps the per-cell accumulators of the optional outputs (rc_coef, Retorned, the last step's vfluxes).
// T2: some tank uses the kinematic-wave speed (speed_type 2); the all-type-1 kernel carries none of that code.
template <bool DIAG, bool T2, bool SNAP, bool STRICT>
__global__ void __launch_bounds__(TH, DIAG ? 2 : (T2 ? 3 : 4)) k_shia5(const Shia5Args a)
{
    constexpr int K = DIAG ? 14 : 2;
    __shared__ double stage[WARPS][TB][K];
    const int64_t N = a.N, T = a.N_reg;
    const int pset = blockIdx.y;
    const int64_t c = (int64_t)blockIdx.x * TH + threadIdx.x;
    const bool live = c < N;
    const int64_t cc = live ? c : 0;
    const float *cal = a.calib + (int64_t)pset * 11;
    const float dt = a.p.dt;
    StepReducer<K, float> red(stage, a.part + ((int64_t)pset * a.nblk + blockIdx.x) * (T + 1) * K);

    // ---- setup (Modelos:229-254)
    float vs[4], hs[4], coef[3], expo[3];
#pragma unroll
    for (int i = 0; i < 4; i++) vs[i] = a.cs[(0 + i) * N + cc] * cal[i] * dt;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        if (a.p.hspeed_given) hs[i] = a.hspeed[((int64_t)pset * 4 + i) * N + cc];
        else {
            int st = i < 3 ? a.p.speed_type[i] : 1;
            hs[i] = st == 1 ? a.cs[(4 + i) * N + cc] * cal[4 + i] : 0.0f;
        }
    }
#pragma unroll
    for (int i = 0; i < 3; i++) {
        coef[i] = a.cs[(4 + i) * N + cc] * cal[4 + i];
        expo[i] = a.cs[(8 + i) * N + cc];
    }
    const float H1 = a.cs[12 * N + cc] * cal[8], H2 = a.cs[13 * N + cc] * cal[9], H3 = a.cs[14 * N + cc] * cal[10];
    const float L = a.cs[15 * N + cc], m3mm = a.cs[16 * N + cc] / 1000.0f;
    float *sto = a.sto + (int64_t)pset * 5 * N;
    float S1 = sto[0 * N + cc], S2 = sto[1 * N + cc], S3 = sto[2 * N + cc], S4 = sto[3 * N + cc];
    const float S5 = sto[4 * N + cc];
    float fl[3];                              // type-1 outflow fraction, constant in time (Modelos:462)
#pragma unroll
    for (int i = 0; i < 3; i++) fl[i] = 1.0f - L / (hs[i] * dt + L);
    const bool t2_0 = T2 && a.p.speed_type[0] != 1, t2_1 = T2 && a.p.speed_type[1] != 1, t2_2 = T2 && a.p.speed_type[2] != 1;
    const float invH1 = 1.0f / H1;
    const bool raq = a.p.retorno_aq > 0.0f, rgr = a.p.retorno_gr > 0.0f;
    const int niter = a.p.calc_niter;
    float acum = 0.0f;

    {   // slot 0: the constant sums {S5, h4} (tank 5 and speed row 4 are never touched), then the initial S1..S4, h1..h3
        float v[K];
        v[0] = live ? S5 : 0.0f;
        v[1] = live ? hs[3] : 0.0f;
        if (DIAG) {
            v[2] = live ? S1 : 0.0f; v[3] = live ? S2 : 0.0f; v[4] = live ? S3 : 0.0f; v[5] = live ? S4 : 0.0f;
            v[6] = live ? hs[0] : 0.0f; v[7] = live ? hs[1] : 0.0f; v[8] = live ? hs[2] : 0.0f;
            v[9] = v[10] = v[11] = v[12] = v[13] = 0.0f;
        }
        red.add(v);
    }
    float sto_prev = (S1 + S2) + (S3 + S4);

    const int32_t *rain_cc = a.rain + cc;                // this cell's column of the rain records
    float ex_ret = 0.0f, ex_v1 = 0.0f, ex_v2 = 0.0f, ex_v3 = 0.0f, ex_v4 = 0.0f;   // (DIAG) this step's return flow and vfluxes
    float rc1 = 0.0f, rc2 = 0.0f, ret_acc = 0.0f;
    // record 1 = no rain (:366-367); a record beyond the file is the Fortran's 'fuera del rango' read error, whose
    // RainInt is undefined: defined here as no rain (the host layer rejects such headers before the launch).
    // Row offset of step t's record in the rain array, -1 = no rain:
    auto rain_row = [&](int64_t t) -> int64_t {
        const int pos = a.pos[t];
        return (unsigned)(pos - 2) < (unsigned)(a.n_rec - 1) ? (int64_t)(pos - 1) * N : (int64_t)-1;
    };
    auto one_step = [&](int64_t t, int64_t row, float evp_t, float &vR, float &vq) {
        float R = 0.0f;
        if (row >= 0) {                                                                     // :370 (RainInt/1000.)
            const float ri = (float)rain_cc[row];
            R = STRICT ? ri / 1000.0f : ri * 0.001f;
        }
        acum = acum + R;                                                                   // :380
        float v1 = fmaxf(0.0f, (R - H1) + S1);                                             // :393
        S1 = (S1 + R) - v1;                                                                // :394
        float evl = fminf((evp_t * vs[0]) * (STRICT ? powf(S1 / H1, 0.6f) : pow06(S1 * invH1)), S1);   // :396-398
        S1 = S1 - evl;                                                                     // :407
        float v2 = fminf(v1, vs[1]); S2 = (S2 + v1) - v2;                                  // :409-412
        float v3 = fminf(v2, vs[2]); S3 = (S3 + v2) - v3;
        float v4 = fminf(v3, vs[3]); S4 = (S4 + v3) - v4;
        if (raq) { float ra = fmaxf(0.0f, S4 - H3); S3 = S3 + ra; S4 = S4 - ra; }          // :414-418
        float rg = 0.0f;
        if (rgr) { rg = fmaxf(0.0f, S3 - H2); S2 = S2 + rg; S3 = S3 - rg; }                // :420-427
        if (DIAG) {                                                                        // :424-438 (optional outputs)
            ex_ret = rg; ex_v1 = v1; ex_v2 = v2 + rg; ex_v3 = v3 - rg; ex_v4 = v4;
            if (a.retorned_per_step) ret_acc = rg; else ret_acc = ret_acc + rg;
            rc1 = (rc1 + ex_v2) - ex_v3;
            rc2 = rc2 + R;
        }
        float hf, ar;
        if (!t2_0) hf = fl[0] * S2;                                                        // :461-491
        else { ar = 0.0f; calc_speed<!STRICT>(S2 * m3mm, coef[0], expo[0], L, dt, niter, hs[0], ar); hf = fminf(((ar * hs[0]) * dt) / m3mm, S2); }
        S2 = S2 - hf;
        if (!t2_1) hf = fl[1] * S3;
        else { ar = 0.0f; calc_speed<!STRICT>(S3 * m3mm, coef[1], expo[1], L, dt, niter, hs[1], ar); hf = fminf(((ar * hs[1]) * dt) / m3mm, S3); }
        S3 = S3 - hf;
        if (!t2_2) hf = fl[2] * S4;
        else { ar = 0.0f; calc_speed<!STRICT>(S4 * m3mm, coef[2], expo[2], L, dt, niter, hs[2], ar); hf = fminf(((ar * hs[2]) * dt) / m3mm, S4); }
        S4 = S4 - hf;

        const float sto_now = (S1 + S2) + (S3 + S4);
        vR = live ? R : 0.0f;
        vq = live ? ((sto_now - sto_prev) + (v4 + evl)) : 0.0f;                            // :459, :533
        sto_prev = sto_now;
        if (SNAP) {                                                                        // :496-500 (save_storage)
            const int slot = a.snap_slot[t];
            if (slot >= 0 && live) {
                float *o = a.snap_out + ((int64_t)pset * a.n_snap + slot) * 5 * N + c;
                o[0] = S1; o[N] = S2; o[2 * N] = S3; o[3 * N] = S4; o[4 * N] = S5;
            }
        }
    };
    if (DIAG) {
        for (int64_t t = 0; t < T; t++) {
            float v[K];
            one_step(t, rain_row(t), a.evp[t], v[0], v[1]);
            if (DIAG) {
                v[2] = live ? S1 : 0.0f; v[3] = live ? S2 : 0.0f; v[4] = live ? S3 : 0.0f; v[5] = live ? S4 : 0.0f;
                v[6] = live ? hs[0] : 0.0f; v[7] = live ? hs[1] : 0.0f; v[8] = live ? hs[2] : 0.0f;
                v[9] = live ? ex_ret : 0.0f;
                v[10] = live ? ex_v1 : 0.0f; v[11] = live ? ex_v2 : 0.0f; v[12] = live ? ex_v3 : 0.0f; v[13] = live ? ex_v4 : 0.0f;
            }
            red.add(v);
        }
    } else {
        red.flush();                                      // slot 0 is out; batches of TB steps follow
        // The batch's record rows and evaporation rates are the same for every cell: one thread per step looks them up
        // (and checks the record index) once per block and batch, the step itself reads them from shared memory.
        __shared__ long long s_row[TB];
        __shared__ float s_evp[TB];
        for (int64_t tb = 0; tb < T; tb += TB) {
            if (threadIdx.x < TB) {
                const int64_t t = tb + threadIdx.x;
                s_row[threadIdx.x] = t < T ? rain_row(t) : -1;
                s_evp[threadIdx.x] = t < T ? a.evp[t] : 0.0f;
            }
            __syncthreads();                              // (the flush at the end of the batch keeps the next batch's writes behind these reads)
            float q[2][TB];
#pragma unroll
            for (int u = 0; u < TB; u++) {
                q[0][u] = 0.0f; q[1][u] = 0.0f;
                if (tb + u < T) one_step(tb + u, (int64_t)s_row[u], s_evp[u], q[0][u], q[1][u]);
            }
            red.add_batch(q, (int)(T - tb < TB ? T - tb : TB));
            red.flush();
        }
    }
    red.flush();
    if (live) {
        sto[0 * N + c] = S1; sto[1 * N + c] = S2; sto[2 * N + c] = S3; sto[3 * N + c] = S4;
        float *hso = a.hspeed + (int64_t)pset * 4 * N;
#pragma unroll
        for (int i = 0; i < 4; i++) hso[i * N + c] = hs[i];
        if (pset == 0 && a.acum_rain) a.acum_rain[c] = acum;
        if (DIAG) {
            if (a.vfluxes) {
                float *o = a.vfluxes + (int64_t)pset * 4 * N + c;
                o[0] = ex_v1; o[N] = ex_v2; o[2 * N] = ex_v3; o[3 * N] = ex_v4;
            }
            if (a.rc_coef) {                                                               // :543-544
                float *o = a.rc_coef + (int64_t)pset * 2 * N + c;
                o[0] = rc1 > rc2 ? rc2 : rc1; o[N] = rc2;
            }
            if (a.retorned) a.retorned[(int64_t)pset * N + c] = a.retorned_per_step ? 0.0f : ret_acc;
        }
    }
}

// totals [P][T+1][K] -> reference outputs.  balance(t) = sum(StoOut) - StoAtras - entradas + salidas
// (Modelos:533); Mean_Rain(t) = rain_sum / N (:494); mean_storage (:518); mean_speed (:522).
__global__ void k_shia5_outputs(const double *__restrict__ totals, int K, int64_t T, int64_t N, int P,
                                float *__restrict__ balance, float *__restrict__ mean_rain,
                                float *__restrict__ mean_storage, float *__restrict__ mean_speed,
                                float *__restrict__ mean_retorno, float *__restrict__ mean_vfluxes)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)P * T) return;
    int p = (int)(i / T);
    int64_t t = i - (int64_t)p * T;
    const double *tot = totals + (int64_t)p * (T + 1) * K;
    const double *now = tot + (t + 1) * K, *prev = tot + t * K;
    balance[i] = (float)(now[1] - now[0]);
    if (p == 0 && mean_rain) mean_rain[t] = (float)(now[0] / (double)N);
    if (K == 14) {
        if (mean_retorno) mean_retorno[i] = (float)(now[9] / (double)N);                    // Modelos:525-527
        if (mean_vfluxes) {                                                                  // Modelos:502-507
            float *mf = mean_vfluxes + (int64_t)p * 4 * T;
            for (int k = 0; k < 4; k++) mf[k * T + t] = (float)(now[10 + k] / (double)N);
        }
        if (mean_storage) {
            float *ms = mean_storage + (int64_t)p * 5 * T;
            for (int k = 0; k < 4; k++) ms[k * T + t] = (float)(now[2 + k] / (double)N);
            ms[4 * T + t] = (float)(tot[0] / (double)N);       // tank 5 is never touched
        }
        if (mean_speed) {
            float *mv = mean_speed + (int64_t)p * 4 * T;
            for (int k = 0; k < 3; k++) mv[k * T + t] = (float)(now[6 + k] / (double)N);
            mv[3 * T + t] = (float)(tot[1] / (double)N);
        }
    }
}

// ------------------------------------------------------------------------------------ SHIA 3-tank
struct Shia3Args {
    wmf_shia3_params p;
    int64_t nsteps, ncell;
    const double *rain, *et;
    double *cap, *grav, *aqu, *qsl, *qbl;
    double *qs_t, *qb_t, *qt_t, *cap_t, *grav_t, *aqu_t;
    double *part;               // [nblk][nsteps+1][6]
};

__global__ void __launch_bounds__(TH, 4) k_shia3(const Shia3Args a)
{
    constexpr int K = 6;        // {q_total, storage, P, ET, q_super, q_base}
    __shared__ double stage[WARPS][TB][K];
    const int64_t n = a.ncell, T = a.nsteps;
    const int64_t i = (int64_t)blockIdx.x * TH + threadIdx.x;
    const bool live = i < n;
    const int64_t ii = live ? i : 0;
    StepReducer<K, double> red(stage, a.part + (int64_t)blockIdx.x * (T + 1) * K);
    const wmf_shia3_params &p = a.p;
    const double dt_h = p.dt / 3600.0;
    const double infil_cap = p.max_infil_mm_h * dt_h;
    double c = a.cap[ii], g = a.grav[ii], q = a.aqu[ii];
    double q_super = 0.0, q_base = 0.0;
    {
        double v[K] = {0.0, live ? (c + g) + q : 0.0, 0.0, 0.0, 0.0, 0.0};       // core.py:147
        red.add(v);
    }
    for (int64_t t = 0; t < T; t++) {
        double rain_step = fmax(a.rain[t * n + ii], 0.0) * dt_h;                  // :152
        double et_step = a.et ? fmax(a.et[t * n + ii], 0.0) * p.et_scale : 0.0;   // :154-158
        double infil = fmin(rain_step, infil_cap);                                // :160
        double exceso = rain_step - infil;
        c += infil;
        double et_c, et_g, et_a;
        if (p.proportional) {                                                     // :165-176
            double w = (c + g) + q, fc = 0.0, fg = 0.0, fa = 0.0;
            if (w > 0) { fc = c / w; fg = g / w; fa = q / w; }
            et_c = fmin(et_step * fc, c); et_g = fmin(et_step * fg, g); et_a = fmin(et_step * fa, q);
        } else {                                                                  // :177-182
            et_c = fmin(et_step, c);
            double rem = et_step - et_c;
            et_g = fmin(rem, g);
            double rem2 = rem - et_g;
            et_a = fmin(rem2, q);
        }
        c -= et_c; g -= et_g; q -= et_a;
        double et_total = (et_c + et_g) + et_a;                                   // :187
        double perc_cg = fmin(p.k_perc_cap_to_grav * c, c);                       // :189-193
        c -= perc_cg; g += perc_cg;
        double perc_ga = fmin(p.k_perc_grav_to_aqu * g, g);                       // :195-200
        double retorno = p.retorno_gr * g;
        g = (g - perc_ga) - retorno;
        q += perc_ga;
        q_base = fmin(p.k_baseflow * q, q);                                       // :202-204
        double drena = p.drena * q;
        q = (q - q_base) + drena;
        double overflow = fmax(c - p.max_capilar, 0.0);                           // :206-208
        c -= overflow; exceso += overflow;
        c = fmin(fmax(c, 0.0), p.max_capilar);                                    // :210-212
        g = fmin(fmax(g, 0.0), p.max_gravita);
        q = fmin(fmax(q, 0.0), p.max_aquifer);
        q_super = exceso + retorno;                                               // :214-215
        double q_total = q_super + q_base;
        if (live) {
            if (a.qs_t) { a.qs_t[t * n + i] = q_super; a.qb_t[t * n + i] = q_base; a.qt_t[t * n + i] = q_total; }
            if (a.cap_t) { a.cap_t[t * n + i] = c; a.grav_t[t * n + i] = g; a.aqu_t[t * n + i] = q; }
        }
        double v[K];
        v[0] = live ? q_total : 0.0; v[1] = live ? (c + g) + q : 0.0; v[2] = live ? rain_step : 0.0;
        v[3] = live ? et_total : 0.0; v[4] = live ? q_super : 0.0; v[5] = live ? q_base : 0.0;
        red.add(v);
    }
    red.flush();
    if (live) {
        a.cap[i] = c; a.grav[i] = g; a.aqu[i] = q;
        if (a.qsl) a.qsl[i] = q_super;
        if (a.qbl) a.qbl[i] = q_base;
    }
}


// ------------------------------------------------------------------------------------ N4: channel routing (extension)
// The reference declares tank 5, `drena` and the control points but never moves water between cells, so
// this has no reference behaviour to match; it is defined in DESIGN.md 4.5 and checked against a CPU restatement of its own.
// Cell c needs (c, t - 1) and (its donors, t).  With rank(c) = Dmax - depth(c) (depth = links to the outlet) every donor
// has rank(c) - 1, so on "diagonal" d = rank + t all cells can work at once, each on its own time step t = d - rank(c):
// N_reg + Dmax launches instead of N_reg x Dmax level sweeps.  A donor's outflow of step t is written on diagonal d - 1
// and read on diagonal d: two buffers, by the parity of d.  Donors are pulled in ascending index order (CSR), so the
// result does not depend on the schedule.
struct RouteArgs {
    int64_t N, N_reg;
    int32_t n_rec;
    const float *ct;             // [13][N]: vs0..3, fl2..4 (lateral fractions), fl5, 1/H1, H1, H2, H3, m3mm
    const int32_t *rain;         // [n_rec][N]
    const int32_t *pos;          // [N_reg]
    const float *evp;            // [N_reg]
    const int32_t *rank;         // [N]
    const int32_t *don_ptr;      // [N + 1]
    const int32_t *don_idx;      // [don_ptr[N]]
    const int32_t *ctrl;         // [N]: row of Q (> 0), 0 none; the outlet (last cell) always writes row 0
    float *sto;                  // [5][N] in / out
    float *acum;                 // [N]
    float *qbuf;                 // [2][N] outflow volumes [m3]
    float *Q;                    // [n_ctrl][N_reg]
    float dt, retorno_gr, retorno_aq;
};

__global__ void k_route_setup(int64_t N, const float *__restrict__ cs, const float *__restrict__ calib, float dt,
                              float *__restrict__ ct, float *__restrict__ acum)
{
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= N) return;
    const float L = cs[15 * N + c];
#pragma unroll
    for (int i = 0; i < 4; i++) ct[i * N + c] = cs[i * N + c] * calib[i] * dt;                    // vspeed (Modelos:242)
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const float hs = cs[(4 + i) * N + c] * calib[4 + i];
        ct[(4 + i) * N + c] = 1.0f - L / (hs * dt + L);                                           // :462, and tank 5 alike
    }
    const float H1 = cs[12 * N + c] * calib[8];
    ct[8 * N + c] = 1.0f / H1;
    ct[9 * N + c] = H1;
    ct[10 * N + c] = cs[13 * N + c] * calib[9];
    ct[11 * N + c] = cs[14 * N + c] * calib[10];
    ct[12 * N + c] = cs[16 * N + c] / 1000.0f;
    acum[c] = 0.0f;
}

// one cell on one diagonal: time step t = d - rank(c)
__device__ __forceinline__ void route_cell(const RouteArgs &a, int64_t c, int64_t d)
{
    const int64_t N = a.N;
    const int64_t t = d - a.rank[c];
    if (t < 0 || t >= a.N_reg) return;
    const float *ct = a.ct;
    const int pos = a.pos[t];
    float R = 0.0f;
    if ((unsigned)(pos - 2) < (unsigned)(a.n_rec - 1)) R = (float)a.rain[(int64_t)(pos - 1) * N + c] * 0.001f;
    a.acum[c] = a.acum[c] + R;
    float S1 = a.sto[c], S2 = a.sto[N + c], S3 = a.sto[2 * N + c], S4 = a.sto[3 * N + c], S5 = a.sto[4 * N + c];
    const float H1 = ct[9 * N + c], H2 = ct[10 * N + c], H3 = ct[11 * N + c], m3mm = ct[12 * N + c];
    float v1 = fmaxf(0.0f, (R - H1) + S1);
    S1 = (S1 + R) - v1;
    float evl = fminf((a.evp[t] * ct[c]) * pow06(S1 * ct[8 * N + c]), S1);
    S1 = S1 - evl;
    float v2 = fminf(v1, ct[N + c]); S2 = (S2 + v1) - v2;
    float v3 = fminf(v2, ct[2 * N + c]); S3 = (S3 + v2) - v3;
    float v4 = fminf(v3, ct[3 * N + c]); S4 = (S4 + v3) - v4;
    if (a.retorno_aq > 0.0f) { float ra = fmaxf(0.0f, S4 - H3); S3 = S3 + ra; S4 = S4 - ra; }
    if (a.retorno_gr > 0.0f) { float rg = fmaxf(0.0f, S3 - H2); S2 = S2 + rg; S3 = S3 - rg; }
    const float hf2 = ct[4 * N + c] * S2; S2 = S2 - hf2;
    const float hf3 = ct[5 * N + c] * S3; S3 = S3 - hf3;
    const float hf4 = ct[6 * N + c] * S4; S4 = S4 - hf4;
    const float lat = (hf2 + hf3) + hf4;
    const float *qprev = a.qbuf + ((d - 1) & 1) * N;
    float qin = 0.0f;
    for (int32_t k = a.don_ptr[c]; k < a.don_ptr[c + 1]; k++) qin = qin + __ldcg(qprev + a.don_idx[k]);   // (written by other blocks)
    S5 = (S5 + lat) + qin / m3mm;
    const float q5 = ct[7 * N + c] * S5;
    S5 = S5 - q5;
    const float qm3 = q5 * m3mm;
    a.qbuf[(d & 1) * N + c] = qm3;
    if (c == N - 1) a.Q[t] = qm3 / a.dt;
    const int32_t j = a.ctrl[c];
    if (j > 0) a.Q[(int64_t)j * a.N_reg + t] = qm3 / a.dt;
    a.sto[c] = S1; a.sto[N + c] = S2; a.sto[2 * N + c] = S3; a.sto[3 * N + c] = S4; a.sto[4 * N + c] = S5;
}

__global__ void __launch_bounds__(256) k_route_diag(const RouteArgs a, int64_t d)
{
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c < a.N) route_cell(a, c, d);
}

// All diagonals in ONE cooperative launch: the blocks meet at a barrier between diagonals instead of the host launching
// N_reg + max_rank kernels.  The structure is in upstream-first BFS order, so the cells of one rank are contiguous
// (level_start[r] = first cell of rank r) and a diagonal touches only the ranks whose time step is in range; a cell is
// always served by the same thread (its index modulo the grid size), so only the outflow buffer crosses threads.
__device__ __forceinline__ void grid_barrier(unsigned int *bar, unsigned int nblocks)
{
    __syncthreads();
    if (threadIdx.x == 0) {
        volatile unsigned int *gen = bar + 1;
        const unsigned int g = *gen;
        __threadfence();
        if (atomicAdd(bar, 1u) == nblocks - 1) {
            *bar = 0;
            __threadfence();
            atomicAdd(bar + 1, 1u);
        } else {
            while (*gen == g) {}
        }
        __threadfence();
    }
    __syncthreads();
}

__global__ void __launch_bounds__(256) k_route_persistent(const RouteArgs a, const int32_t *__restrict__ level_start, int max_rank,
                                                          unsigned int *bar)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x, me = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (int64_t d = 0; d < a.N_reg + max_rank; d++) {
        const int64_t r_lo = d - a.N_reg + 1 > 0 ? d - a.N_reg + 1 : 0, r_hi = d < max_rank ? d : max_rank;
        const int64_t c_lo = level_start[r_lo], c_hi = level_start[r_hi + 1];
        for (int64_t c = c_lo / stride * stride + me; c < c_hi; c += stride)
            if (c >= c_lo) route_cell(a, c, d);
        grid_barrier(bar, gridDim.x);
    }
}

}  // namespace

extern "C" {

size_t wmf_shia5_workspace(int64_t N, int64_t N_reg, int32_t P, int32_t diagnostics)
{
    int64_t nblk = (N + TH - 1) / TH;
    const size_t K = diagnostics ? 14 : 2;                   // per-step sums the run reduces (see k_shia5)
    size_t part = (size_t)P * nblk * (N_reg + 1) * K * sizeof(double);
    size_t tot = (size_t)P * (N_reg + 1) * K * sizeof(double);
    return wmf_align(part) + wmf_align(tot) + 1024;
}

int wmf_shia5_run(const wmf_shia5_params *params_host, int64_t N, int64_t N_reg, int64_t n_rec, int32_t P,
                  const float *cell_static, const int32_t *rain_records, const int32_t *pos_evento, const float *evp,
                  const float *calib, float *sto, float *hspeed, float *balance, float *mean_rain, float *acum_rain,
                  float *mean_storage, float *mean_speed, void *workspace, size_t ws_bytes, void *stream)
{
    return wmf_shia5_run_snap(params_host, N, N_reg, n_rec, P, cell_static, rain_records, pos_evento, evp, calib, sto, hspeed,
                              balance, mean_rain, acum_rain, mean_storage, mean_speed, nullptr, 0, nullptr, workspace, ws_bytes,
                              stream);
}

int wmf_shia5_run_snap(const wmf_shia5_params *params_host, int64_t N, int64_t N_reg, int64_t n_rec, int32_t P,
                       const float *cell_static, const int32_t *rain_records, const int32_t *pos_evento, const float *evp,
                       const float *calib, float *sto, float *hspeed, float *balance, float *mean_rain, float *acum_rain,
                       float *mean_storage, float *mean_speed, const int32_t *snap_slot, int32_t n_snap, float *snap_out,
                       void *workspace, size_t ws_bytes, void *stream)
{
    return wmf_shia5_run_ext(params_host, N, N_reg, n_rec, P, cell_static, rain_records, pos_evento, evp, calib, sto, hspeed, balance,
                             mean_rain, acum_rain, mean_storage, mean_speed, snap_slot, n_snap, snap_out, nullptr, nullptr, nullptr,
                             nullptr, nullptr, 0, workspace, ws_bytes, stream);
}

int wmf_shia5_run_ext(const wmf_shia5_params *params_host, int64_t N, int64_t N_reg, int64_t n_rec, int32_t P,
                      const float *cell_static, const int32_t *rain_records, const int32_t *pos_evento, const float *evp,
                      const float *calib, float *sto, float *hspeed, float *balance, float *mean_rain, float *acum_rain,
                      float *mean_storage, float *mean_speed, const int32_t *snap_slot, int32_t n_snap, float *snap_out,
                      float *mean_retorno, float *mean_vfluxes, float *vfluxes, float *rc_coef, float *retorned,
                      int32_t retorned_per_step, void *workspace, size_t ws_bytes, void *stream)
{
    cudaStream_t st = as_stream(stream);
    WMF_REQUIRE((snap_slot == nullptr) == (snap_out == nullptr) && n_snap >= 0 && (snap_slot == nullptr || n_snap > 0), WMF_E_ARG,
                "snap_slot, snap_out and n_snap go together");
    WMF_REQUIRE(params_host && cell_static && rain_records && pos_evento && evp && calib && sto && hspeed && balance,
                WMF_E_ARG, "null pointer");
    WMF_REQUIRE(N > 0 && N_reg > 0 && n_rec > 0 && P > 0 && P <= 65535, WMF_E_ARG, "bad size");
    const bool diag = mean_storage || mean_speed || mean_retorno || mean_vfluxes || vfluxes || rc_coef || retorned;
    const int K = diag ? 14 : 2;
    const int nblk = (int)((N + TH - 1) / TH);
    WsCursor cur(workspace, ws_bytes);
    double *part = cur.take<double>((size_t)P * nblk * (N_reg + 1) * K);
    double *totals = cur.take<double>((size_t)P * (N_reg + 1) * K);
    WMF_REQUIRE(cur.ok, WMF_E_WORKSPACE, "workspace too small");
    Shia5Args a;
    WMF_REQUIRE(n_rec < ((int64_t)1 << 31), WMF_E_TOO_LARGE, "more than 2^31-1 rain records");
    a.p = *params_host; a.N = N; a.N_reg = N_reg; a.n_rec = (int32_t)n_rec; a.cs = cell_static; a.rain = rain_records; a.pos = pos_evento;
    a.evp = evp; a.calib = calib; a.sto = sto; a.hspeed = hspeed; a.acum_rain = acum_rain; a.part = part; a.nblk = nblk;
    a.snap_slot = snap_slot; a.snap_out = snap_out; a.n_snap = n_snap;
    a.vfluxes = vfluxes; a.rc_coef = rc_coef; a.retorned = retorned; a.retorned_per_step = retorned_per_step;
    dim3 grid(nblk, P);
    const bool t2 = a.p.speed_type[0] != 1 || a.p.speed_type[1] != 1 || a.p.speed_type[2] != 1;
    const int variant = (diag ? 4 : 0) | (t2 ? 2 : 0) | (snap_slot ? 1 : 0);
#define WMF_SHIA5(D_, T_, S_) \
    do { if (a.p.strict) k_shia5<D_, T_, S_, true><<<grid, TH, 0, st>>>(a); else k_shia5<D_, T_, S_, false><<<grid, TH, 0, st>>>(a); } while (0)
    switch (variant) {
    case 0: WMF_SHIA5(false, false, false); break;
    case 1: WMF_SHIA5(false, false, true); break;
    case 2: WMF_SHIA5(false, true, false); break;
    case 3: WMF_SHIA5(false, true, true); break;
    case 4: WMF_SHIA5(true, false, false); break;
    case 5: WMF_SHIA5(true, false, true); break;
    case 6: WMF_SHIA5(true, true, false); break;
    default: WMF_SHIA5(true, true, true); break;
    }
#undef WMF_SHIA5
    const int64_t TK = (N_reg + 1) * K;
    fold_blocks(part, nblk, TK, P, totals, st);
    k_shia5_outputs<<<blocks_for((int64_t)P * N_reg, 256), 256, 0, st>>>(totals, K, N_reg, N, P, balance, mean_rain,
                                                                         mean_storage, mean_speed, mean_retorno, mean_vfluxes);
    WMF_LAUNCH_CHECK();
    return 0;
}

size_t wmf_shia3_workspace(int64_t ncell, int64_t nsteps)
{
    int64_t nblk = (ncell + TH - 1) / TH;
    return wmf_align((size_t)nblk * (nsteps + 1) * 6 * sizeof(double)) + 1024;
}

int wmf_shia3_run(const wmf_shia3_params *params_host, int64_t nsteps, int64_t ncell, const double *rain,
                  const double *et, double *cap, double *grav, double *aqu, double *q_super_last, double *q_base_last,
                  double *sums, double *q_super_t, double *q_base_t, double *q_total_t, double *cap_t, double *grav_t,
                  double *aqu_t, void *workspace, size_t ws_bytes, void *stream)
{
    cudaStream_t st = as_stream(stream);
    WMF_REQUIRE(params_host && rain && cap && grav && aqu && sums, WMF_E_ARG, "null pointer");
    WMF_REQUIRE(nsteps > 0 && ncell > 0, WMF_E_ARG, "bad size");
    WMF_REQUIRE((q_super_t == nullptr) == (q_base_t == nullptr) && (q_super_t == nullptr) == (q_total_t == nullptr),
                WMF_E_ARG, "flux series must be all set or all null");
    WMF_REQUIRE((cap_t == nullptr) == (grav_t == nullptr) && (cap_t == nullptr) == (aqu_t == nullptr), WMF_E_ARG,
                "storage series must be all set or all null");
    const int nblk = (int)((ncell + TH - 1) / TH);
    WsCursor cur(workspace, ws_bytes);
    double *part = cur.take<double>((size_t)nblk * (nsteps + 1) * 6);
    WMF_REQUIRE(cur.ok, WMF_E_WORKSPACE, "workspace too small");
    Shia3Args a;
    a.p = *params_host; a.nsteps = nsteps; a.ncell = ncell; a.rain = rain; a.et = et; a.cap = cap; a.grav = grav;
    a.aqu = aqu; a.qsl = q_super_last; a.qbl = q_base_last; a.qs_t = q_super_t; a.qb_t = q_base_t; a.qt_t = q_total_t;
    a.cap_t = cap_t; a.grav_t = grav_t; a.aqu_t = aqu_t; a.part = part;
    k_shia3<<<nblk, TH, 0, st>>>(a);
    const int64_t TK = (nsteps + 1) * 6;
    fold_blocks(part, nblk, TK, 1, sums, st);
    WMF_LAUNCH_CHECK();
    return 0;
}


/* N4 host helper: depth[i] = number of links from cell i to the outlet, for a structure in upstream-first order
 * (receiver of i = n - drain[i], drain 0 = outlet).  Returns the largest depth, or -1 on a malformed structure. */
int32_t wmf_basin_depth_host(const int32_t *drain, int64_t n, int32_t *depth)
{
    int32_t dmax = 0;
    for (int64_t i = n - 1; i >= 0; i--) {
        const int32_t dr = drain[i];
        if (dr == 0) { depth[i] = 0; continue; }
        const int64_t j = n - dr;
        if (j <= i || j >= n) return -1;
        depth[i] = depth[j] + 1;
        if (depth[i] > dmax) dmax = depth[i];
    }
    return dmax;
}

size_t wmf_shia5_route_workspace(int64_t N) { return wmf_align((size_t)N * 13 * 4) + wmf_align((size_t)N * 2 * 4) + 2048; }

int wmf_shia5_route_run(const wmf_shia5_params *params_host, int64_t N, int64_t N_reg, int64_t n_rec,
                        const float *cell_static, const int32_t *rain_records, const int32_t *pos_evento, const float *evp,
                        const float *calib, const int32_t *rank, int32_t max_rank, const int32_t *don_ptr,
                        const int32_t *don_idx, const int32_t *ctrl, int32_t n_ctrl, const int32_t *level_start, float *sto,
                        float *acum_rain, float *Q, void *workspace, size_t ws_bytes, void *stream)
{
    cudaStream_t st = as_stream(stream);
    WMF_REQUIRE(params_host && cell_static && rain_records && pos_evento && evp && calib && rank && don_ptr && don_idx &&
                ctrl && sto && acum_rain && Q && workspace, WMF_E_ARG, "null pointer");
    WMF_REQUIRE(N > 0 && N_reg > 0 && n_rec > 0 && n_ctrl >= 1 && max_rank >= 0, WMF_E_ARG, "bad size");
    WMF_REQUIRE(params_host->speed_type[0] == 1 && params_host->speed_type[1] == 1 && params_host->speed_type[2] == 1,
                WMF_E_ARG, "the routing extension is defined for speed type 1");
    WsCursor cur(workspace, ws_bytes);
    float *ct = cur.take<float>((size_t)N * 13), *qbuf = cur.take<float>((size_t)N * 2);
    WMF_REQUIRE(cur.ok, WMF_E_WORKSPACE, "workspace too small");
    WMF_CUDA_CHECK(cudaMemsetAsync(qbuf, 0, sizeof(float) * 2 * (size_t)N, st));
    WMF_CUDA_CHECK(cudaMemsetAsync(Q, 0, sizeof(float) * (size_t)n_ctrl * (size_t)N_reg, st));
    k_route_setup<<<blocks_for(N, 256), 256, 0, st>>>(N, cell_static, calib, params_host->dt, ct, acum_rain);
    RouteArgs a;
    a.N = N; a.N_reg = N_reg; a.n_rec = (int32_t)(n_rec < 0x7fffffff ? n_rec : 0x7fffffff); a.ct = ct; a.rain = rain_records; a.pos = pos_evento; a.evp = evp; a.rank = rank;
    a.don_ptr = don_ptr; a.don_idx = don_idx; a.ctrl = ctrl; a.sto = sto; a.acum = acum_rain; a.qbuf = qbuf; a.Q = Q;
    a.dt = params_host->dt; a.retorno_gr = params_host->retorno_gr; a.retorno_aq = params_host->retorno_aq;
    // ranks in ascending cell order (the structure's BFS order): one cooperative launch walks all diagonals; otherwise
    // (a caller's own ordering) one launch per diagonal
    bool persistent = level_start != nullptr;
    int dev = 0, coop = 0, per_sm = 0, sms = 0;
    if (persistent) {
        WMF_CUDA_CHECK(cudaGetDevice(&dev));
        WMF_CUDA_CHECK(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev));
        WMF_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        WMF_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_route_persistent, 256, 0));
        persistent = coop && per_sm > 0;
    }
    if (persistent) {
        unsigned int *bar = cur.take<unsigned int>(64);
        WMF_REQUIRE(cur.ok, WMF_E_WORKSPACE, "workspace too small");
        WMF_CUDA_CHECK(cudaMemsetAsync(bar, 0, 2 * sizeof(unsigned int), st));
        int64_t want = (N + 255) / 256;                      // no more blocks than the basin has work for
        int grid = (int)(want < (int64_t)sms * per_sm ? want : (int64_t)sms * per_sm);
        int mr = (int)max_rank;
        void *args[] = {(void *)&a, (void *)&level_start, (void *)&mr, (void *)&bar};
        WMF_CUDA_CHECK(cudaLaunchCooperativeKernel((const void *)k_route_persistent, dim3(grid), dim3(256), args, 0, st));
    } else {
        for (int64_t d = 0; d < N_reg + max_rank; d++) k_route_diag<<<blocks_for(N, 256), 256, 0, st>>>(a, d);
    }
    WMF_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
