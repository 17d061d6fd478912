// d8_sweep.cu -- full-raster D8 accumulation, tiled algorithm (WMF_ALGO_SWEEP).
// Replaces wmf_py/cu_py/basics.py:124-177 for ANY input, bit-exact, in three device phases:
//
//   A  per tile (TW x TH cells), in isolation: for every boundary cell the boundary cell where flow
//      entering there leaves the tile again (link), and the accumulation every exit cell would have if
//      nothing entered the tile (exit_acc).  One upward sweep gives every cell the label of the exit its
//      path leaves through; exit_acc is the histogram of those labels.
//   G  the tiles' exit cells form a contracted forest (parent(x) = link(receiver(x)), weight =
//      exit_acc(x)).  int32 accumulations: one 64-bit word per node (pending arrivals | weight + finished
//      children), one L2 atomic per hop, every finalised exit pushes its total into the entry cell it
//      drains to (inflow).  int64 accumulations of rasters of 2^32 cells and more: the generic engine of
//      forest.cu.
//   C  per tile again, the downward accumulation sweep with the external inflow injected at the
//      boundary cells -> final ACC.  A raster below 2^32 cells is accumulated in 32 bits whatever the
//      output type and widened (or converted to float64) by the store.
//
// Per-tile solvers.  FAST: one warp owns a tile and streams its rows with the running row held in
// registers -- contributions from the neighbouring row are lane-local or one shuffle away, and East /
// West chains inside a row are segmented warp scans (shuffle).  It is exact for tiles whose in-tile
// flow never points upwards and has no E<->W two-cycles ("simple" tiles; a Scheidegger network is all
// simple).  GENERAL: every other tile (and every tile that receives flow from an unresolved, i.e.
// cyclic, upstream) is solved by one thread block with the in-degree chain walk in shared memory.
// Both produce the same outputs, so the result does not depend on which solver ran: integer adds,
// any schedule is bit-exact.
//
// Cycle semantics (basics.py:166-175): a cell is finalised only when all its donors are; others
// keep the partial sum of finalised inflows without their own +1 and pass nothing on.  The
// contracted forest carries this through `blocked` nodes and the `ext_unres` entry flags.
#include <cstdlib>
#include <mutex>
#include "common.cuh"

namespace {

constexpr int NC = 8;                   // cells per lane in the fast kernels
constexpr int NWD = NC / 4;             // 32-bit DIR words per lane per row
constexpr int TW = 32 * NC;             // tile width  (one warp)
constexpr int TH = 64;                  // tile height
constexpr int NSLOT = 2 * TW + 2 * TH;  // boundary slots: top row, bottom row, left col, right col
constexpr int TCELLS = TW * TH;
constexpr int WPB = 4;                  // warps (= tiles) per block in the fast kernels
constexpr uint32_t LNONE = 1023u;       // "no exit": above every slot, and the last bin of phase A's label histogram
constexpr uint32_t FULL = 0xffffffffu;
// work flags (one 32-bit word each, cleared before every accumulation): do the rarely needed tile solvers have anything to do?
constexpr int F_COMPLEX = 0, F_CHAIN = 1, F_UNRES = 2;   // tiles for the jump solver / for the chain walk / fed by an unresolved upstream
constexpr int F_NEXT_A = 4, F_NEXT_C = 5;                // work counters of the jump kernels
constexpr int NFLAGS = 8;
// raise a work flag: on a terrain raster EVERY tile raises the same one, so only those who still read 0 store (measured:
// no difference -- the 190 us k_sweepA spends on a 16384^2 terrain raster before its tiles give up are not these stores)
__device__ __forceinline__ void raise_flag(uint32_t *flags, int which)
{
    if (__ldcg(&flags[which]) == 0u) flags[which] = 1u;
}

// meta word of a boundary slot: [15:0] link, [19:16] k, [20] locally unresolved, [21] active
__host__ __device__ __forceinline__ uint32_t meta_pack(uint32_t link, int k, bool unres, bool active)
{
    return (link & 0xFFFFu) | ((uint32_t)k << 16) | ((uint32_t)unres << 20) | ((uint32_t)active << 21);
}

struct Geo {
    int64_t ny, nx;          // this raster (or row stripe)
    int tiles_x, tiles_y;
    int64_t gy0, gny;        // the stripe holds rows [gy0, gy0 + ny) of a raster of gny rows (gy0 = 0, gny = ny when whole)
    int extra_edge;          // stripes, 32-bit sums: `extra` holds inflow only for the stripe's first / last tile row
    int tile0;               // band launches: the first tile this launch handles (the `ntiles` argument is the end)
};

// canonical boundary slot of local cell (lr, lc) in a tile of th x tw cells, or -1
__host__ __device__ __forceinline__ int slot_of(int lr, int lc, int th, int tw)
{
    if (lr == 0) return lc;
    if (lr == th - 1) return TW + lc;
    if (lc == 0) return 2 * TW + lr;
    if (lc == tw - 1) return 2 * TW + TH + lr;
    return -1;
}
// inverse: local cell of a slot, false when the slot is unused / non-canonical for this tile
__host__ __device__ __forceinline__ bool cell_of(int s, int th, int tw, int &lr, int &lc)
{
    if (s < TW) { lr = 0; lc = s; return lc < tw; }
    if (s < 2 * TW) { lr = th - 1; lc = s - TW; return th > 1 && lc < tw; }
    if (s < 2 * TW + TH) { lr = s - 2 * TW; lc = 0; return lr > 0 && lr < th - 1; }
    lr = s - 2 * TW - TH; lc = tw - 1;
    return tw > 1 && lr > 0 && lr < th - 1;
}

__device__ __forceinline__ void tile_dims(const Geo &g, int tile, int &ty, int &tx, int &th, int &tw)
{
    ty = tile / g.tiles_x;
    tx = tile - ty * g.tiles_x;
    int64_t rem_r = g.ny - (int64_t)ty * TH, rem_c = g.nx - (int64_t)tx * TW;
    th = rem_r < TH ? (int)rem_r : TH;
    tw = rem_c < TW ? (int)rem_c : TW;
}

// segmented warp scan of an East chain: returns the inflow arriving at this lane's cell 0 from the
// lane on its left.  (A, P) = value leaving the lane to the right with zero carry-in, and whether
// the lane is all-pass.
template <typename T>
__device__ __forceinline__ T chain_carry_from_left(T A, bool P, int lane)
{
    if (__ballot_sync(FULL, P) == 0) {              // no lane passes a carry through: one hop is exact
        T c = __shfl_up_sync(FULL, A, 1);
        return lane == 0 ? (T)0 : c;
    }
    int p = P;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        T a2 = __shfl_up_sync(FULL, A, off);
        int p2 = __shfl_up_sync(FULL, p, off);
        if (lane >= off) { if (p) A += a2; p &= p2; }
    }
    T c = __shfl_up_sync(FULL, A, 1);
    return lane == 0 ? (T)0 : c;
}
template <typename T>
__device__ __forceinline__ T chain_carry_from_right(T A, bool P, int lane)
{
    if (__ballot_sync(FULL, P) == 0) {
        T c = __shfl_down_sync(FULL, A, 1);
        return lane == 31 ? (T)0 : c;
    }
    int p = P;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        T a2 = __shfl_down_sync(FULL, A, off);
        int p2 = __shfl_down_sync(FULL, p, off);
        if (lane + off < 32) { if (p) A += a2; p &= p2; }
    }
    T c = __shfl_down_sync(FULL, A, 1);
    return lane == 31 ? (T)0 : c;
}
// copy-scans for labels: value of the chain end is copied back along the chain
// (phase A keeps a label as the shared-memory ADDRESS of its histogram bin -- see k_sweepA --, so "no exit" is a parameter)
__device__ __forceinline__ uint32_t label_carry_from_right(uint32_t V, bool P, int lane, uint32_t none = LNONE)
{
    if (__ballot_sync(FULL, P) != 0) {
        int p = P;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            uint32_t v2 = __shfl_down_sync(FULL, V, off);
            int p2 = __shfl_down_sync(FULL, p, off);
            if (lane + off < 32) { if (p) V = v2; p &= p2; }
            else if (p) { V = none; }
        }
    }
    uint32_t c = __shfl_down_sync(FULL, V, 1);
    return lane == 31 ? none : c;
}
__device__ __forceinline__ uint32_t label_carry_from_left(uint32_t V, bool P, int lane, uint32_t none = LNONE)
{
    if (__ballot_sync(FULL, P) != 0) {
        int p = P;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            uint32_t v2 = __shfl_up_sync(FULL, V, off);
            int p2 = __shfl_up_sync(FULL, p, off);
            if (lane >= off) { if (p) V = v2; p &= p2; }
            else if (p) { V = none; }
        }
    }
    uint32_t c = __shfl_up_sync(FULL, V, 1);
    return lane == 0 ? none : c;
}

// ---- packed variant for int32 accumulations -----------------------------------------------------
// ONE 64-bit word per boundary slot:
//   [31:0]  sum: own weight + totals of finished children (the node's total once it is finalised)
//   [43:32] pending arrivals: children + 1 for the node's own walker (+ 1 more when the node is blocked)
//   [57:54] the cell's normalised D8 code (phase A writes it beside the weight)
// k_g_build counts the children into the parents' words (atomics) and writes the 4-byte parent pointer of every exit.
// A walker brings its total and its "-1 pending" in ONE atomic; whoever takes pending to zero finalises the node and
// carries on downstream (the parent's parent is fetched before the atomic, so a hop is one L2 round trip).  Nothing is
// pushed into the tiles: phase C PULLS the inflow of its boundary cells from the words of the (at most five) cells
// across the tile edge -- a word whose code points at the cell and whose pending field is zero is a finalised donor
// (slots that are not exits are never walked and keep their pending arrival).
constexpr unsigned long long PEND1 = 1ull << 32;
constexpr int W_K = 54;
constexpr int32_t PAR_ROOT = -1, PAR_NONE = -2, PAR_STRIPE = -3;   // values of the parent array beside a parent (>= 0)
__device__ __forceinline__ uint32_t w_pending(unsigned long long w) { return (uint32_t)(w >> 32) & 0xFFFu; }
__device__ __forceinline__ int w_k(unsigned long long w) { return (int)(w >> W_K) & 0xF; }
static_assert(3 * NSLOT + 2 < 0xFFF, "the children count fits its field");

struct TileGeo { int last_th, last_tw; };
__device__ __forceinline__ TileGeo tile_geo(const Geo &g)
{
    return TileGeo{(int)(g.ny - (int64_t)(g.tiles_y - 1) * TH), (int)(g.nx - (int64_t)(g.tiles_x - 1) * TW)};
}

// where exit slot s of tile (ty, tx) drains: the tile of the entry (and its slot in s2), -2 = into a neighbouring row
// stripe, -1 = not an exit / no receiver.  Tile geometry only; k is the slot's normalised code, active its activity.
__device__ __forceinline__ int exit_target(const Geo &g, const TileGeo &tg, int ty, int tx, int s, int k, bool active, int &s2)
{
    const int th = ty == g.tiles_y - 1 ? tg.last_th : TH, tw = tx == g.tiles_x - 1 ? tg.last_tw : TW;
    int lr, lc;
    if (!cell_of(s, th, tw, lr, lc)) return -1;
    if (!active || k == 8) return -1;
    const int rr = lr + d8_dr(k), cc = lc + d8_dc(k);
    if (rr >= 0 && rr < th && cc >= 0 && cc < tw) return -1;            // stays in the tile: not an exit
    const int ty2 = ty + (rr < 0 ? -1 : (rr >= th ? 1 : 0)), tx2 = tx + (cc < 0 ? -1 : (cc >= tw ? 1 : 0));
    if (tx2 < 0 || tx2 >= g.tiles_x) return -1;                         // off the raster's side: no receiver
    if (ty2 < 0 || ty2 >= g.tiles_y) return -2;                         // into a neighbouring row stripe
    const int th2 = ty2 == g.tiles_y - 1 ? tg.last_th : TH, tw2 = tx2 == g.tiles_x - 1 ? tg.last_tw : TW;
    const int lr2 = rr < 0 ? th2 - 1 : (rr >= th ? 0 : rr), lc2 = cc < 0 ? tw2 - 1 : (cc >= tw ? 0 : cc);
    s2 = slot_of(lr2, lc2, th2, tw2);
    return ty2 * g.tiles_x + tx2;
}
__device__ __forceinline__ int build_target(const Geo &g, int ty, int tx, int last_th, int last_tw, int s, uint32_t m, int &s2)
{
    return exit_target(g, TileGeo{last_th, last_tw}, ty, tx, s, (int)((m >> 16) & 0xF), ((m >> 21) & 1) != 0, s2);
}

// ---- phase C's side of the packed forest: the inflow of a boundary cell, pulled from the donors across the tile edge
// value a neighbour's word contributes when that neighbour is an exit with code kexp: its total once finalised
__device__ __forceinline__ uint32_t donor_val(unsigned long long w, int kexp)
{
    const uint32_t hi = (uint32_t)(w >> 32);                // [11:0] pending, [25:22] k
    return (hi & (0xFFFu | (0xFu << 22))) == ((uint32_t)kexp << 22) ? (uint32_t)w : 0u;
}
// any tile, any slot: the sum over the (up to five) cells outside the tile that drain into slot s; `unres` is set when
// one of them is an exit that never finalised
template <typename WalkedFn>
__device__ __forceinline__ uint32_t pull_slot(const Geo &g, const TileGeo &tg, const unsigned long long *__restrict__ wq, int ty,
                                              int tx, int th, int tw, int s, bool &unres, WalkedFn walked)
{
    int lr, lc;
    if (!cell_of(s, th, tw, lr, lc)) return 0u;
    uint32_t sum = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) {                           // the neighbour that reaches (lr, lc) with code k
        const int rr = lr - d8_dr(k), cc = lc - d8_dc(k);
        if (rr >= 0 && rr < th && cc >= 0 && cc < tw) continue;
        const int ty2 = ty + (rr < 0 ? -1 : (rr >= th ? 1 : 0)), tx2 = tx + (cc < 0 ? -1 : (cc >= tw ? 1 : 0));
        if (ty2 < 0 || ty2 >= g.tiles_y || tx2 < 0 || tx2 >= g.tiles_x) continue;
        const int th2 = ty2 == g.tiles_y - 1 ? tg.last_th : TH, tw2 = tx2 == g.tiles_x - 1 ? tg.last_tw : TW;
        const int lr2 = rr < 0 ? th2 - 1 : (rr >= th ? 0 : rr), lc2 = cc < 0 ? tw2 - 1 : (cc >= tw ? 0 : cc);
        const unsigned long long w = __ldcg(&wq[(int64_t)(ty2 * g.tiles_x + tx2) * NSLOT + slot_of(lr2, lc2, th2, tw2)]);
        const uint32_t hi = (uint32_t)(w >> 32);
        if (((hi >> 22) & 0xFu) == (uint32_t)k) {          // a donor: finalised, never finalising (unresolved), or not an
            if ((hi & 0xFFFu) == 0u) sum += (uint32_t)w;    // exit at all (`walked` tells the last two apart)
            else if (walked((int64_t)(ty2 * g.tiles_x + tx2) * NSLOT + slot_of(lr2, lc2, th2, tw2))) unres = true;
        }
    }
    return sum;
}

// A FULL tile's 640 inflows by one warp, vectorised: ext[] in shared memory (slot order).  Its neighbours above and to
// the left are full tiles too; the ones below / to the right may be ragged, which does not move their first row / column.
__device__ __forceinline__ void pull_full_tile(const Geo &g, const unsigned long long *__restrict__ wq, const uint8_t *__restrict__ tile_top,
                                               int tile, int ty, int tx, int lane, uint32_t *ext)
{
    const bool up = ty > 0, dn = ty + 1 < g.tiles_y, lf = tx > 0, rt = tx + 1 < g.tiles_x;
    auto word = [&](int t, int s) { return __ldcg(&wq[(int64_t)t * NSLOT + s]); };
    // ---- first row (donors: last row of the tile above, codes SE / S / SW) and last row (first row of the tile below,
    // codes NE / N / NW); the corner cells' diagonal donors live in the diagonal tiles
#pragma unroll
    for (int side = 0; side < 2; side++) {
        const bool have = side == 0 ? up : dn;
        const int t = side == 0 ? tile - g.tiles_x : tile + g.tiles_x;
        const int kl = side == 0 ? 1 : 7, km = side == 0 ? 2 : 6, kr = side == 0 ? 3 : 5;   // donor at c-1 / c / c+1
        uint32_t v[NC];
#pragma unroll
        for (int j = 0; j < NC; j++) v[j] = 0;
        uint32_t to_right = 0, to_left = 0;                 // what this lane's last / first donor sends to the next / previous lane
        if (have && (side == 0 || tile_top[t])) {           // (a tile without upward flow in its first row feeds nothing upwards)
            const ulonglong2 *src = reinterpret_cast<const ulonglong2 *>(wq + (int64_t)t * NSLOT + (side == 0 ? TW : 0) + NC * lane);
            unsigned long long w[NC];
#pragma unroll
            for (int q = 0; q < NC / 2; q++) { const ulonglong2 x = __ldcg(src + q); w[2 * q] = x.x; w[2 * q + 1] = x.y; }
#pragma unroll
            for (int j = 0; j < NC; j++) {
                v[j] += donor_val(w[j], km);
                if (j + 1 < NC) v[j + 1 < NC ? j + 1 : 0] += donor_val(w[j], kl);
                if (j > 0) v[j > 0 ? j - 1 : 0] += donor_val(w[j], kr);
            }
            to_right = donor_val(w[NC - 1], kl);
            to_left = donor_val(w[0], kr);
        }
        const uint32_t from_left = __shfl_up_sync(FULL, to_right, 1), from_right = __shfl_down_sync(FULL, to_left, 1);
        if (lane > 0) v[0] += from_left;
        if (lane < 31) v[NC - 1] += from_right;
        if (have && lane == 0 && lf) v[0] += donor_val(word(t - 1, side == 0 ? 2 * TW - 1 : TW - 1), kl);
        if (have && lane == 31 && rt) v[NC - 1] += donor_val(word(t + 1, side == 0 ? TW : 0), kr);
        uint32_t *dst = ext + (side == 0 ? 0 : TW) + NC * lane;
#pragma unroll
        for (int j = 0; j < NC; j++) dst[j] = v[j];
    }
    __syncwarp();
    // ---- first column (donors: last column of the tile to the left, codes SE / E / NE from the rows r-1 / r / r+1) and
    // last column (first column of the tile to the right, codes SW / W / NW).  Rows 0 and TH-1 are corner cells: their
    // slots belong to the first / last row, the side donors are ADDED there.
#pragma unroll
    for (int side = 0; side < 2; side++) {
        const bool have = side == 0 ? lf : rt;
        const int t = side == 0 ? tile - 1 : tile + 1;
        const int ku = side == 0 ? 1 : 3, km = side == 0 ? 0 : 4, kd = side == 0 ? 7 : 5;   // donor in row r-1 / r / r+1
        uint32_t a_mid = 0, a_dn = 0, a_up = 0, b_mid = 0, b_dn = 0, b_up = 0;   // by the DONOR's row: lane (a) and lane + 32 (b)
        if (have) {
            // the donor column's slot in row r: corners sit in the first / last row's slot range
            const int ra = lane, rb = lane + 32;
            const int sa = ra == 0 ? (side == 0 ? TW - 1 : 0) : (side == 0 ? 2 * TW + TH : 2 * TW) + ra;
            const int sb = rb == TH - 1 ? (side == 0 ? 2 * TW - 1 : TW) : (side == 0 ? 2 * TW + TH : 2 * TW) + rb;
            const unsigned long long wa = word(t, sa), wb = word(t, sb);
            a_mid = donor_val(wa, km); a_dn = donor_val(wa, ku); a_up = donor_val(wa, kd);   // to row r / r+1 / r-1
            b_mid = donor_val(wb, km); b_dn = donor_val(wb, ku); b_up = donor_val(wb, kd);
        }
        const uint32_t a_from_above = __shfl_up_sync(FULL, a_dn, 1), a_from_below = __shfl_down_sync(FULL, a_up, 1);
        const uint32_t b_from_above = __shfl_up_sync(FULL, b_dn, 1), b_from_below = __shfl_down_sync(FULL, b_up, 1);
        const uint32_t a31_dn = __shfl_sync(FULL, a_dn, 31), b0_up = __shfl_sync(FULL, b_up, 0);
        const uint32_t va = a_mid + (lane > 0 ? a_from_above : 0u) + (lane < 31 ? a_from_below : b0_up);
        const uint32_t vb = b_mid + (lane > 0 ? b_from_above : a31_dn) + (lane < 31 ? b_from_below : 0u);
        // rows lane and lane + 32 of this tile's first / last column
        const int col_base = side == 0 ? 2 * TW : 2 * TW + TH;
        if (lane == 0) ext[side == 0 ? 0 : TW - 1] += va;                       // row 0: the first row's corner slot
        else ext[col_base + lane] = va;
        if (lane == 31) ext[side == 0 ? TW : 2 * TW - 1] += vb;                 // row TH-1: the last row's corner slot
        else ext[col_base + 32 + lane] = vb;
    }
    __syncwarp();
}

// One row of the downward sweep, any codes (generic path).  b[] = base (own cell + inflow from above
// + external inflow) for active cells; returns tot[] and d[] = contributions to the row below.
// Sets `twocycle` when the row holds an E -> W adjacent pair (the tile is then not "simple").
template <typename T>
__device__ __forceinline__ void sweep_row_down(const int (&k)[NC], int am, const T (&b)[NC], T (&tot)[NC], T (&d)[NC],
                                               int lane, bool &twocycle)
{
    T e_in[NC], w_in[NC];
    {   // East chains: e_in[j] = inflow from the west neighbour
        T S = 0;
        bool P = true;
#pragma unroll
        for (int j = 0; j < NC; j++) { bool ps = k[j] == 0; S = ps ? b[j] + S : (T)0; P = P && ps; }
        e_in[0] = chain_carry_from_left<T>(S, P, lane);
#pragma unroll
        for (int j = 1; j < NC; j++) e_in[j] = (k[j - 1] == 0) ? b[j - 1] + e_in[j - 1] : (T)0;
    }
    bool hasW = false;
#pragma unroll
    for (int j = 0; j < NC; j++) hasW = hasW || k[j] == 4;
    if (__any_sync(FULL, hasW)) {
        T S = 0;
        bool P = true;
#pragma unroll
        for (int j = NC - 1; j >= 0; j--) { bool ps = k[j] == 4; S = ps ? b[j] + S : (T)0; P = P && ps; }
        w_in[NC - 1] = chain_carry_from_right<T>(S, P, lane);
#pragma unroll
        for (int j = NC - 2; j >= 0; j--) w_in[j] = (k[j + 1] == 4) ? b[j + 1] + w_in[j + 1] : (T)0;
        // E -> W adjacent pairs are two-cycles: not handled by the scans
        int k0_right = __shfl_down_sync(FULL, k[0], 1);
        bool tc = lane < 31 && k[NC - 1] == 0 && k0_right == 4;
#pragma unroll
        for (int j = 0; j + 1 < NC; j++) tc = tc || (k[j] == 0 && k[j + 1] == 4);
        twocycle = twocycle || __any_sync(FULL, tc);
    } else {
#pragma unroll
        for (int j = 0; j < NC; j++) w_in[j] = 0;
    }
#pragma unroll
    for (int j = 0; j < NC; j++) tot[j] = ((am >> j) & 1) ? b[j] + e_in[j] + w_in[j] : (T)0;
    // contributions to the row below: S from the same column, SE from the left neighbour, SW from the right one
    T se_out = (k[NC - 1] == 1) ? tot[NC - 1] : (T)0, sw_out = (k[0] == 3) ? tot[0] : (T)0;
    T from_left = __shfl_up_sync(FULL, se_out, 1), from_right = __shfl_down_sync(FULL, sw_out, 1);
    if (lane == 0) from_left = 0;
    if (lane == 31) from_right = 0;
#pragma unroll
    for (int j = 0; j < NC; j++) {
        T v = (k[j] == 2) ? tot[j] : (T)0;
        v += (j > 0) ? ((k[j > 0 ? j - 1 : 0] == 1) ? tot[j > 0 ? j - 1 : 0] : (T)0) : from_left;
        v += (j < NC - 1) ? ((k[j < NC - 1 ? j + 1 : NC - 1] == 3) ? tot[j < NC - 1 ? j + 1 : NC - 1] : (T)0) : from_right;
        d[j] = v;
    }
}

// ---------------------------------------------------------------------------------------------
// fast solver: staging, row classification and the lean "plain row" step
// ---------------------------------------------------------------------------------------------
// The tile's raw DIR words (NC codes = NWD words per lane per row) -- and the mask words when there is
// a mask -- stream through shared memory in chunks of CH rows, double buffered with cp.async (LDGSTS):
// while a chunk is swept the next one is in flight, so the sweeps never wait on HBM row by row and a
// warp holds only 2 * CH rows of shared memory, which keeps occupancy register-bound.
constexpr int CH = 8;
constexpr int ROWW = 32 * NWD;          // 32-bit words of one staged row of one warp
constexpr int HBINS = 1024;            // label histogram bins of phase A: NSLOT slots ... LNONE
static_assert(NSLOT <= (int)LNONE && (int)LNONE == HBINS - 1, "every label is a histogram bin");
template <int BYTES>
__device__ __forceinline__ void cp_async(void *smem_dst, const void *gsrc)
{
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"(d), "l"(gsrc), "n"(BYTES) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// TM variant (rasters whose width is a multiple of the tile width, 16-byte aligned): the chunk is fetched by the TMA
// unit instead -- one bulk copy (cp.async.bulk, SASS UBLKCP) of a whole 256-byte tile row per lane 0..CH-1, completion
// counted in bytes on an mbarrier the warp waits on.  No per-lane address arithmetic, nothing in the L1TEX queue.
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_init_fence() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    asm volatile("{\n\t.reg .pred p;\n\tWMF_WAIT:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t@p bra WMF_DONE;\n\tbra WMF_WAIT;\n\tWMF_DONE:\n\t}"
                 ::"r"(bar), "r"(parity), "r"(0x989680u) : "memory");      // (the hint lets the warp sleep instead of spinning)
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

template <bool AL, bool TM = false>
struct TileRows {
    uint32_t *wbase, *mbase;    // this lane's NWD words in the two DIR / mask chunk buffers (ROWW per row, CH*ROWW per buffer)
    const int8_t *dir0;         // &dir[gr0][gc0]
    const uint8_t *mask0;       // &mask[gr0][gc0] or nullptr
    int64_t nx, gc0;
    int th;

    // TM: `bar` = shared address of the warp's two mbarriers (8 bytes each), handed in by the kernel at every call so
    // that it does not occupy a register of its own; the parity a wait expects follows from the chunk's place `seq` in
    // the order the chunks are fetched (the buffers alternate, so a buffer's n-th use is chunk 2n or 2n + 1 of that order).
    __device__ __forceinline__ void init(uint32_t *smem_words, int warp, int lane, const int8_t *dir, const uint8_t *mask,
                                         int64_t nx_, int64_t gr0, int64_t gc0_, int th_, uint32_t bar = 0u)
    {
        wbase = smem_words + (size_t)warp * 2 * CH * ROWW + lane * NWD;
        mbase = smem_words + (size_t)(WPB + warp) * 2 * CH * ROWW + lane * NWD;
        dir0 = dir + gr0 * nx_ + gc0_;
        mask0 = mask ? mask + gr0 * nx_ + gc0_ : nullptr;
        nx = nx_; gc0 = gc0_; th = th_;
        if (TM) {
            if (lane == 0) { mbar_init(bar, 1u); mbar_init(bar + 8u, 1u); mbar_init_fence(); }
            __syncwarp();
        }
    }
    __device__ __forceinline__ void prefetch(int c, uint32_t bar = 0u) const      // chunk c -> buffer c & 1
    {
        const int lane_ = threadIdx.x & 31;
        uint32_t *dw = wbase + (c & 1) * CH * ROWW, *dm = mbase + (c & 1) * CH * ROWW;
        const int r0 = c * CH;
        if (TM) {
            // (the warp has finished reading this buffer: the __syncwarp of arrive(), or nothing has used it yet)
            const int nrows = th - r0 < CH ? th - r0 : CH;
            const uint32_t b = bar + 8u * (uint32_t)(c & 1);
            if (lane_ == 0) mbar_expect_tx(b, (uint32_t)nrows * (uint32_t)TW * (mask0 ? 2u : 1u));
            __syncwarp();
            if (lane_ < nrows) {
                const int64_t off = (int64_t)(r0 + lane_) * nx - NC * lane_;             // row r0 + lane, the tile's first column
                bulk_g2s((uint32_t)__cvta_generic_to_shared(dw - lane_ * NWD + lane_ * ROWW), dir0 + off, (uint32_t)TW, b);
                if (mask0) bulk_g2s((uint32_t)__cvta_generic_to_shared(dm - lane_ * NWD + lane_ * ROWW), mask0 + off, (uint32_t)TW, b);
            }
            return;
        }
        if (AL) {
            if (gc0 < nx) {
                const int8_t *sd = dir0 + (int64_t)r0 * nx;
                const uint8_t *sm = mask0 ? mask0 + (int64_t)r0 * nx : nullptr;
                if (r0 + CH <= th) {
#pragma unroll
                    for (int r = 0; r < CH; r++, sd += nx) cp_async<4 * NWD>(dw + r * ROWW, sd);
                    if (sm) {
#pragma unroll
                        for (int r = 0; r < CH; r++, sm += nx) cp_async<4 * NWD>(dm + r * ROWW, sm);
                    }
                } else {
                    for (int r = 0; r0 + r < th; r++, sd += nx) {
                        cp_async<4 * NWD>(dw + r * ROWW, sd);
                        if (sm) { cp_async<4 * NWD>(dm + r * ROWW, sm); sm += nx; }
                    }
                }
            } else {
#pragma unroll
                for (int r = 0; r < CH; r++)
#pragma unroll
                    for (int q = 0; q < NWD; q++) { dw[r * ROWW + q] = 0; if (mask0) dm[r * ROWW + q] = 0; }
            }
        } else {
            for (int r = 0; r < CH && r0 + r < th; r++) {
#pragma unroll
                for (int q = 0; q < NWD; q++) {
                    uint32_t ww = 0, mm = 0;
#pragma unroll
                    for (int j = 0; j < 4; j++)
                        if (gc0 + 4 * q + j < nx) {
                            const int64_t o = (int64_t)(r0 + r) * nx + 4 * q + j;
                            ww |= (uint32_t)(uint8_t)dir0[o] << (8 * j);
                            if (mask0) mm |= (uint32_t)(mask0[o] != 0) << (8 * j);
                        }
                    dw[r * ROWW + q] = ww;
                    if (mask0) dm[r * ROWW + q] = mm;
                }
            }
        }
        cp_async_commit();
    }
    // wait for chunk c; start fetching chunk `next` (or nothing when out of range) into the other buffer
    __device__ __forceinline__ void arrive(int c, int seq, int next, int nchunks, uint32_t bar = 0u) const
    {
        if (TM) mbar_wait(bar + 8u * (uint32_t)(c & 1), (uint32_t)(seq >> 1) & 1u);
        else cp_async_wait_all();
        __syncwarp();
        if (next >= 0 && next < nchunks) prefetch(next, bar);
    }
    // nothing may be in flight into the warp's buffers when it leaves (`pending`: a chunk was prefetched and not waited for)
    __device__ __forceinline__ void drain(int pending, int seq, uint32_t bar = 0u) const
    {
        if (TM) { if (pending >= 0) mbar_wait(bar + 8u * (uint32_t)(pending & 1), (uint32_t)(seq >> 1) & 1u); }
        else cp_async_wait_all();
        __syncwarp();
    }
    // this lane's words of chunk c: word q of row r of the chunk at [r * ROWW + q]
    __device__ __forceinline__ const uint32_t *chunk(int c) const { return wbase + (c & 1) * CH * ROWW; }
    __device__ __forceinline__ const uint32_t *mchunk(int c) const { return mbase + (c & 1) * CH * ROWW; }   // ... of the mask
    // active bits of the lane's NC cells in row lr
    __device__ __forceinline__ int active(int lr) const
    {
        int am = 0;
        const uint32_t *mp = mbase + (((lr / CH) & 1) * CH + (lr % CH)) * ROWW;
#pragma unroll
        for (int j = 0; j < NC; j++) {
            bool a = gc0 + j < nx;
            if (mask0) a = a && ((mp[j / 4] >> (8 * (j % 4))) & 0xFFu) != 0;
            am |= (int)a << j;
        }
        return am;
    }
};

// decode staged words: k[NC] (8 = none; inactive cells are none); on border tiles a code whose
// receiver lies outside the raster becomes none.
template <int ENC>
__device__ __forceinline__ void decode_row(const uint32_t *wp, int am, const Geo &g, int64_t gr, int64_t gc0, bool border,
                                           int (&k)[NC])
{
#pragma unroll
    for (int j = 0; j < NC; j++) {
        int kk = ((am >> j) & 1) ? d8_norm<ENC>((int)(int8_t)(wp[j / 4] >> (8 * (j % 4)))) : 8;
        if (border && kk != 8) {
            int64_t rr = g.gy0 + gr + d8_dr(kk), cc = gc0 + j + d8_dc(kk);
            if (rr < 0 || rr >= g.gny || cc < 0 || cc >= g.nx) kk = 8;
        }
        k[j] = kk;
    }
}

// ---- lean "plain row" machinery ---------------------------------------------------------------
// A plain row holds only E / SE / S / SW codes (every cell active, tile interior).
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel)
{
    uint32_t r;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(sel));
    return r;
}

// internal-coded word of 4 cells (E 0, SE 1, S 2, SW 3), or "needs the generic row" when any byte is
// something else.  Keypad words are translated with a PRMT table lookup.
template <int ENC>
__device__ __forceinline__ uint32_t plain_word(uint32_t w, bool &generic)
{
    if (ENC == WMF_ENC_INTERNAL) {
        generic = (w & 0xFCFCFCFCu) != 0;
        return w;
    }
    // keypad 1 SW, 2 S, 3 SE, 6 E -> 3, 2, 1, 0 ; anything else -> 4 (generic)
    bool junk = (w & 0xF8F8F8F8u) != 0;                       // codes >= 8 (N, NE) and negatives
    uint32_t t = (w & 0x00070007u) | ((w & 0x07000700u) >> 4); // two nibbles per half word
    uint32_t sel = (t & 0xFFu) | ((t >> 8) & 0xFF00u);        // selector nibbles = the four codes (0..7)
    uint32_t k = prmt(0x01020304u, 0x04000404u, sel);         // table[code]: 0:4 1:3 2:2 3:1 | 4:4 5:4 6:0 7:4
    generic = junk || (k & 0x04040404u) != 0;
    return k;
}
template <int ENC>
__device__ __forceinline__ bool plain_words(const uint32_t *wp, uint32_t (&kw)[NWD])
{
    bool generic = false;
#pragma unroll
    for (int q = 0; q < NWD; q++) { bool g1; kw[q] = plain_word<ENC>(wp[q], g1); generic = generic || g1; }
    return generic;
}

// exact per-byte tests of a word -> 0x01 per byte
__device__ __forceinline__ uint32_t byte_nonzero(uint32_t w) { return ((((w & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | w) >> 7) & 0x01010101u; }
__device__ __forceinline__ uint32_t byte_zero(uint32_t w) { return byte_nonzero(w) ^ 0x01010101u; }

// Lean rows of a raster WITH a mask (MK): a row qualifies when no ACTIVE cell holds a W / NW / N / NE code.  Inactive
// cells and active cells without a receiver code (pits, nodata codes) ride along: av = 0x01 per active cell, dv = 0x01
// per active cell with a lean direction (E / SE / S / SW).  Flow into an inactive cell needs no test of its own -- in
// the label sweep an inactive cell carries "no exit", which is what its donors must copy; in the accumulation sweep
// whatever arrives at a cell is multiplied by the cell's own activity.
template <int ENC>
__device__ __forceinline__ bool lean_words_mk(const uint32_t *wp, const uint32_t *mp, uint32_t (&kw)[NWD], uint32_t (&dv)[NWD],
                                              uint32_t (&av)[NWD])
{
    bool reject = false;
#pragma unroll
    for (int q = 0; q < NWD; q++) {
        const uint32_t w = wp[q];
        uint32_t ki;                                  // per byte: 0..3 lean direction, 4..7 another direction, else none
        if (ENC == WMF_ENC_INTERNAL) ki = w;
        else {
            // keypad 1 SW, 2 S, 3 SE, 6 E -> 3, 2, 1, 0; 4 W, 7 NW, 8 N, 9 NE -> 4; 0, 5, negatives, > 9 -> 8 (none)
            const uint32_t t = (w & 0x00070007u) | ((w & 0x07000700u) >> 4);
            const uint32_t sel = (t & 0xFFu) | ((t >> 8) & 0xFF00u);
            const uint32_t low = prmt(0x01020308u, 0x04000804u, sel);     // table[code & 7]
            const uint32_t hb = byte_nonzero(w & 0xF8F8F8F8u), is89 = byte_zero((w & 0xFEFEFEFEu) ^ 0x08080808u);
            ki = (low & ~(hb * 0xFFu)) | ((hb << 3) - (is89 << 2));
        }
        av[q] = byte_nonzero(mp[q]);
        const uint32_t t = ki & 0xFCFCFCFCu;
        dv[q] = byte_zero(t) & av[q];
        reject = reject || (byte_zero(t ^ 0x04040404u) & av[q]) != 0;
        kw[q] = ki;
    }
    return reject;
}

// Direction tests of the NC cells as 0/1 integers.  Every select below is a multiply by 0/1 fused into
// an IMAD, which runs on the FMA pipe: the kernels are ALU-pipe bound (LOP3 / PRMT / SEL), so moving the
// selects off that pipe matters more than the instruction count.
struct RowBits { uint32_t e[NC], se[NC], s[NC], sw[NC]; bool all_e; };
__device__ __forceinline__ RowBits row_bits(const uint32_t (&kw)[NWD])
{
    RowBits m;
    m.all_e = true;
#pragma unroll
    for (int q = 0; q < NWD; q++) {
        const uint32_t x0 = kw[q] & 0x01010101u, x1 = (kw[q] >> 1) & 0x01010101u;
        const uint32_t sw = x0 & x1, se = x0 ^ sw, s = x1 ^ sw, e = (x0 | x1) ^ 0x01010101u;
        m.all_e = m.all_e && e == 0x01010101u;
#pragma unroll
        for (int j = 0; j < 4; j++) {             // byte j as a 0/1 integer
            m.e[4 * q + j] = prmt(e, 0u, 0x4440u + j);
            m.se[4 * q + j] = prmt(se, 0u, 0x4440u + j);
            m.s[4 * q + j] = prmt(s, 0u, 0x4440u + j);
            m.sw[4 * q + j] = prmt(sw, 0u, 0x4440u + j);
        }
    }
    return m;
}
// the same with a mask: a cell has a direction only where dv is set; egate (nullable) also cuts an East link whose
// receiver is inactive (accumulation sweep).  act[] = the cells' activity as 0/1 integers.
__device__ __forceinline__ RowBits row_bits_mk(const uint32_t (&kw)[NWD], const uint32_t (&dv)[NWD], const uint32_t *egate)
{
    RowBits m;
    m.all_e = true;
#pragma unroll
    for (int q = 0; q < NWD; q++) {
        const uint32_t x0 = kw[q] & dv[q], x1 = (kw[q] >> 1) & dv[q];
        const uint32_t sw = x0 & x1, se = x0 ^ sw, s = x1 ^ sw;
        uint32_t e = (x0 | x1) ^ dv[q];
        if (egate) e &= egate[q];
        m.all_e = m.all_e && e == 0x01010101u;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            m.e[4 * q + j] = prmt(e, 0u, 0x4440u + j);
            m.se[4 * q + j] = prmt(se, 0u, 0x4440u + j);
            m.s[4 * q + j] = prmt(s, 0u, 0x4440u + j);
            m.sw[4 * q + j] = prmt(sw, 0u, 0x4440u + j);
        }
    }
    return m;
}
__device__ __forceinline__ void byte_flags(const uint32_t (&f)[NWD], uint32_t (&out)[NC])
{
#pragma unroll
    for (int j = 0; j < NC; j++) out[j] = prmt(f[j / 4], 0u, 0x4440u + (j % 4));
}
// activity of the cell to the right of each cell (the next lane's first cell for the last one; none beyond lane 31)
__device__ __forceinline__ void right_neighbour_flags(const uint32_t (&av)[NWD], int lane, uint32_t (&out)[NWD])
{
    uint32_t nxt = __shfl_down_sync(FULL, av[0], 1);
    if (lane == 31) nxt = 0;
#pragma unroll
    for (int q = 0; q < NWD; q++) out[q] = __funnelshift_r(av[q], q + 1 < NWD ? av[q + 1 < NWD ? q + 1 : 0] : nxt, 8);
}

// Downward step of a plain row.  d[] holds, for each cell, 1 + the inflow from the row above (the
// cell's own +1 is folded into the accumulate chain of the row above); tot[] out; d[] out for the
// row below.  extL / extR: external inflow of the left- / right-column cell (phase C), else zero.
template <typename T>
__device__ __forceinline__ void plain_row_down_m(const RowBits &m, T (&d)[NC], T (&tot)[NC], int lane, T extL, T extR)
{
    T t[NC];
    t[0] = d[0] + extL;
#pragma unroll
    for (int j = 1; j < NC; j++) t[j] = t[j - 1] * (T)m.e[j - 1] + d[j];
    t[NC - 1] += extR;
    T carry = chain_carry_from_left<T>(t[NC - 1] * (T)m.e[NC - 1], m.all_e, lane);
#pragma unroll
    for (int j = 0; j < NC; j++) { tot[j] = t[j] + carry; carry *= (T)m.e[j]; }
    T from_left = __shfl_up_sync(FULL, tot[NC - 1] * (T)m.se[NC - 1], 1);
    T from_right = __shfl_down_sync(FULL, tot[0] * (T)m.sw[0], 1);
    from_left = lane == 0 ? (T)1 : from_left + 1;            // the "+1" of the cell below rides along
    if (lane == 31) from_right = 0;
#pragma unroll
    for (int j = 0; j < NC; j++) {
        T v = tot[j] * (T)m.s[j] + (j == 0 ? from_left : (T)1);
        if (j > 0) v = tot[j > 0 ? j - 1 : 0] * (T)m.se[j > 0 ? j - 1 : 0] + v;
        if (j < NC - 1) v = tot[j < NC - 1 ? j + 1 : 0] * (T)m.sw[j < NC - 1 ? j + 1 : 0] + v;
        else v += from_right;
        d[j] = v;
    }
}
template <typename T>
__device__ __forceinline__ void plain_row_down(const uint32_t (&kw)[NWD], T (&d)[NC], T (&tot)[NC], int lane, T extL, T extR)
{
    plain_row_down_m<T>(row_bits(kw), d, tot, lane, extL, extR);
}
// With a mask: what has arrived at a cell (from above, and the external inflow the caller has added to d) counts only
// if the cell is active, and an East link into an inactive cell is cut; everything else is the plain step.
template <typename T>
__device__ __forceinline__ void plain_row_down_mk(const uint32_t (&kw)[NWD], const uint32_t (&dv)[NWD], const uint32_t (&av)[NWD],
                                                  T (&d)[NC], T (&tot)[NC], int lane, T extL, T extR)
{
    uint32_t act[NC], rgt[NWD];
    byte_flags(av, act);
    right_neighbour_flags(av, lane, rgt);
    d[0] += extL; d[NC - 1] += extR;
#pragma unroll
    for (int j = 0; j < NC; j++) d[j] *= (T)act[j];
    plain_row_down_m<T>(row_bits_mk(kw, dv, rgt), d, tot, lane, (T)0, (T)0);
}

// Upward (label) step of a plain row of a full interior tile.  A plain row has no pits, so an East
// cell's own term is 0 and "copy the chain end" is own + next * E.  slotL / slotR: boundary slots of
// the left- / right-column cell of this row.  BOTTOM: the tile's last row, where S / SE / SW leave the tile.
// Labels are kept as base + 4 * slot -- the shared-memory address of the label's histogram bin, so that counting a label
// is one `red.shared` with no address arithmetic; every step below only ever selects ONE label, so the form survives.
template <bool BOTTOM, bool MK>
__device__ __forceinline__ void plain_row_up(const uint32_t (&kw)[NWD], const uint32_t (&dv)[NWD], uint32_t (&lb)[NC], int lane,
                                             uint32_t slotL, uint32_t slotR, uint32_t base)
{
    const uint32_t none = base + 4u * LNONE;
    slotL = base + 4u * slotL; slotR = base + 4u * slotR;
    RowBits m = MK ? row_bits_mk(kw, dv, nullptr) : row_bits(kw);
    uint32_t own[NC];
    if (BOTTOM) {
        const uint32_t first = base + 4u * (uint32_t)(TW + NC * lane);
#pragma unroll
        for (int j = 0; j < NC; j++)
            own[j] = MK ? (first + 4u * j) * (m.se[j] + m.s[j] + m.sw[j]) : (first + 4u * j) * (1u - m.e[j]);
    } else {
        const uint32_t lbR0 = __shfl_down_sync(FULL, lb[0], 1), lbL = __shfl_up_sync(FULL, lb[NC - 1], 1);
#pragma unroll
        for (int j = 0; j < NC; j++) {
            const uint32_t right = j < NC - 1 ? lb[j < NC - 1 ? j + 1 : 0] : lbR0, left = j > 0 ? lb[j > 0 ? j - 1 : 0] : lbL;
            own[j] = lb[j] * m.s[j] + right * m.se[j] + left * m.sw[j];
        }
    }
    if (MK) {                                                  // no direction (inactive, pit, nodata code): no exit
        uint32_t has[NC];
        byte_flags(dv, has);
#pragma unroll
        for (int j = 0; j < NC; j++) own[j] += none * (1u - has[j]);
    }
    const uint32_t kR = (kw[NWD - 1] >> 24) & 0xFFu, kL = kw[0] & 0xFFu;
    const bool dR = !MK || (dv[NWD - 1] >> 24) != 0, dL = !MK || (dv[0] & 0xFFu) != 0;
    if (lane == 31 && dR && kR <= 1u) { own[NC - 1] = slotR; m.e[NC - 1] = 0; }   // E / SE leave through the right column
    if (lane == 0 && dL && kL == 3u) own[0] = slotL;                              // SW leaves through the left column
    uint32_t V = own[NC - 1] + none * m.e[NC - 1];           // label at cell 0 if nothing comes from the right
    uint32_t allE = m.e[NC - 1];
#pragma unroll
    for (int j = NC - 2; j >= 0; j--) { V = V * m.e[j] + own[j]; allE &= m.e[j]; }
    uint32_t cur = label_carry_from_right(V, allE != 0, lane, none);
#pragma unroll
    for (int j = NC - 1; j >= 0; j--) { cur = cur * m.e[j] + own[j]; lb[j] = cur; }
}

// byte j of the plain (internal-coded) words
__device__ __forceinline__ int plain_k(const uint32_t (&kw)[NWD], int j) { return (int)((kw[j / 4] >> (8 * (j % 4))) & 0xFFu); }
// ... of a masked lean row: the normalised code (8 = none) and the activity of cell j
__device__ __forceinline__ int plain_k_mk(const uint32_t (&kw)[NWD], const uint32_t (&dv)[NWD], int j)
{
    return ((dv[j / 4] >> (8 * (j % 4))) & 1u) ? (int)((kw[j / 4] >> (8 * (j % 4))) & 3u) : 8;
}
__device__ __forceinline__ bool flag_of(const uint32_t (&f)[NWD], int j) { return ((f[j / 4] >> (8 * (j % 4))) & 1u) != 0; }

// ---------------------------------------------------------------------------------------------
// phase A, fast solver
// ---------------------------------------------------------------------------------------------
template <int ENC, bool AL, bool MK, bool TM>
__global__ void __launch_bounds__(WPB * 32, 6) k_sweepA(const int8_t *__restrict__ dir, const uint8_t *__restrict__ mask,
                                                     Geo g, int ntiles, uint32_t *__restrict__ meta,
                                                     int32_t *__restrict__ exit_acc, uint8_t *__restrict__ tile_class,
                                                     unsigned long long *__restrict__ row_plain,
                                                     unsigned long long *__restrict__ wq, uint8_t *__restrict__ tile_top,
                                                     uint32_t *__restrict__ flags, bool has_multi)
{
    extern __shared__ __align__(16) uint32_t s_dyn[];   // [WPB][2][CH][ROWW] DIR words (+ the same again for mask words)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tile = g.tile0 + blockIdx.x * WPB + warp;
    if (tile >= ntiles) return;
    int ty, tx, th, tw;
    tile_dims(g, tile, ty, tx, th, tw);
    const int64_t gr0 = (int64_t)ty * TH, gc0 = (int64_t)tx * TW + NC * lane;
    uint32_t *m_t = meta + (int64_t)tile * NSLOT;
    int32_t *x_t = exit_acc ? exit_acc + (int64_t)tile * NSLOT : nullptr;
    unsigned long long *w_t = wq + (int64_t)tile * NSLOT;     // packed forest words: pending << 32 | sum
    constexpr unsigned long long BIAS = 1ull << 32;           // every node waits for its own walker
    const int rlane = (tw - 1) / NC, rj = (tw - 1) % NC;      // lane / sub-cell holding the right column
    const bool border = ty == 0 || tx == 0 || ty == g.tiles_y - 1 || tx == g.tiles_x - 1;
    const bool full = th == TH && tw == TW;
    // Lean rows also serve the raster's own edge tiles: a left-column SW cell or a right-column E / SE cell
    // drains nowhere in the row arithmetic and keeps its own slot as exit label; k_g_build / entry_of give
    // such an exit no entry.  Only the raster's last row (S / SE / SW leave the raster) stays generic.
    const bool plain_tile = full;                            // (MK: the raster has a mask)
    const int last_generic = (g.gy0 + gr0 + th == g.gny) ? th - 1 : -1;
    const int nch = (th + CH - 1) / CH;

    // boundary-row stores: NC consecutive slots per lane
    auto store_meta_row = [&](int base, const uint32_t (&mw)[NC]) {
#pragma unroll
        for (int q = 0; q < NC / 4; q++)
            *reinterpret_cast<uint4 *>(m_t + base + NC * lane + 4 * q) = make_uint4(mw[4 * q], mw[4 * q + 1], mw[4 * q + 2], mw[4 * q + 3]);
    };

    // exit_acc without a downward sweep: the local accumulation of an exit cell is the number of tile cells whose
    // path leaves the tile there, i.e. the histogram of the labels of ALL cells, which the label sweep visits anyway.
    // One 1024-bin histogram per warp in shared memory (bin 1023 swallows LNONE).
    int *hist = reinterpret_cast<int *>(s_dyn + (size_t)(mask ? 2 : 1) * WPB * 2 * CH * ROWW) + HBINS * warp;
#pragma unroll
    for (int q = 0; q < HBINS / 128; q++) *reinterpret_cast<int4 *>(hist + 128 * q + 4 * lane) = make_int4(0, 0, 0, 0);
    __syncwarp();
    // (TM: the warp's two mbarriers live in bins no label uses -- labels are slots < NSLOT, or LNONE)
    constexpr int BAR_BIN = 1000;
    static_assert(BAR_BIN >= NSLOT && BAR_BIN + 4 <= (int)LNONE && (BAR_BIN * 4) % 8 == 0, "the mbarriers sit in unused histogram bins");
    const uint32_t hist_sa = (uint32_t)__cvta_generic_to_shared(hist);
#define WMF_BAR_A (hist_sa + 4u * BAR_BIN)
    TileRows<AL, TM> rows;
    rows.init(s_dyn, warp, lane, dir, mask, g.nx, gr0, gc0, th, WMF_BAR_A);
    // A label is kept as the shared-memory address of its bin, hist_sa + 4 * (slot or LNONE): counting it needs no
    // address arithmetic (8 instructions per row of ~195); slot_label() gives the slot back where a meta word is packed.
    const uint32_t lnone = hist_sa + 4u * LNONE;
    auto slot_label = [&](uint32_t lab) { return (lab - hist_sa) >> 2; };
    auto count_labels = [&](const uint32_t (&lab)[NC]) {
#pragma unroll
        for (int j = 0; j < NC; j++) asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(lab[j]) : "memory");
    };
    // side-column slots of a full tile sit in lanes 0 / 31; their forest words are set in one go at the end
    uint32_t *m_side = m_t + 2 * TW + (lane == 31 ? TH : 0);

    // ---- upward sweep: exit labels, bottom row first
    bool complex_tile = false, top_exits = false;
    unsigned long long plain_rows = 0;               // bit lr set: row lr takes the lean path (here and in phase C)
    uint32_t lb[NC];
#pragma unroll
    for (int j = 0; j < NC; j++) lb[j] = lnone;      // labels of the row below
    rows.prefetch(nch - 1, WMF_BAR_A);
    int c_left = nch - 1;                            // the chunk that is in flight when the loop ends early (-1: none)
    for (int c = nch - 1; c >= 0 && !complex_tile; c--) {
        rows.arrive(c, nch - 1 - c, c - 1, nch, WMF_BAR_A);
        c_left = c - 1;
        const int cbeg = c * CH;
        const uint32_t *wc = rows.chunk(c);
        for (int lr = (th < (c + 1) * CH ? th : (c + 1) * CH) - 1; lr >= cbeg; lr--) {
            const uint32_t *wp = wc + (lr - cbeg) * ROWW;
            bool gen = true;
            uint32_t kw[NWD], dv[NWD], av[NWD];
            if (plain_tile && lr != last_generic) {
                if (MK) gen = __any_sync(FULL, lean_words_mk<ENC>(wp, rows.mchunk(c) + (lr - cbeg) * ROWW, kw, dv, av));
                else gen = __any_sync(FULL, plain_words<ENC>(wp, kw));
            }
            if (!gen) {
                plain_rows |= 1ull << lr;
                if (lr == 0 || lr == TH - 1) {        // first / last row of the tile: every cell is a boundary slot
                    if (lr == 0) plain_row_up<false, MK>(kw, dv, lb, lane, 0u, (uint32_t)(TW - 1), hist_sa);
                    else plain_row_up<true, MK>(kw, dv, lb, lane, (uint32_t)TW, (uint32_t)(2 * TW - 1), hist_sa);
                    uint32_t mw[NC];
#pragma unroll
                    for (int j = 0; j < NC; j++)
                        mw[j] = MK ? meta_pack(slot_label(lb[j]), plain_k_mk(kw, dv, j), false, flag_of(av, j))
                                   : meta_pack(slot_label(lb[j]), plain_k(kw, j), false, true);
                    store_meta_row(lr == 0 ? 0 : TW, mw);
                } else {
                    plain_row_up<false, MK>(kw, dv, lb, lane, (uint32_t)(2 * TW + lr), (uint32_t)(2 * TW + TH + lr), hist_sa);
                    if (lane == 0 || lane == 31) {
                        const int js = lane ? NC - 1 : 0;
                        const uint32_t ls = slot_label(lane ? lb[NC - 1] : lb[0]);
                        m_side[lr] = MK ? meta_pack(ls, plain_k_mk(kw, dv, js), false, flag_of(av, js)) : meta_pack(ls, plain_k(kw, js), false, true);
                    }
                }
                count_labels(lb);
                continue;
            }
            int k[NC];
            const int am = rows.active(lr);
            decode_row<ENC>(wp, am, g, gr0 + lr, gc0, border, k);
            {   // tiles the row sweeps cannot solve go to the general solver: upward flow below the first row, and
                // E -> W adjacent pairs (two-cycles)
                bool bad = false;
#pragma unroll
                for (int j = 0; j < NC; j++) bad = bad || (lr > 0 && k[j] >= 5 && k[j] <= 7);
                const int k0_right = __shfl_down_sync(FULL, k[0], 1);
                bad = bad || (lane < 31 && k[NC - 1] == 0 && k0_right == 4);
#pragma unroll
                for (int j = 0; j + 1 < NC; j++) bad = bad || (k[j] == 0 && k[j + 1] == 4);
                if (__any_sync(FULL, bad)) { complex_tile = true; break; }
                if (lr == 0) {
                    bool up0 = false;
#pragma unroll
                    for (int j = 0; j < NC; j++) up0 = up0 || (k[j] >= 5 && k[j] <= 7);
                    top_exits = __any_sync(FULL, up0);
                }
            }
            const uint32_t lb_right0 = __shfl_down_sync(FULL, lb[0], 1), lb_left = __shfl_up_sync(FULL, lb[NC - 1], 1);
            uint32_t own[NC];
            bool pe[NC], pw[NC];
#pragma unroll
            for (int j = 0; j < NC; j++) {
                const int lc = NC * lane + j, kk = k[j];
                const bool top = lr == 0, bot = lr == th - 1, lft = lc == 0, rgt = lc == tw - 1;
                bool exits = (bot && kk >= 1 && kk <= 3) || (top && kk >= 5 && kk <= 7) ||
                             (rgt && (kk == 0 || kk == 1 || kk == 7)) || (lft && kk >= 3 && kk <= 5);
                uint32_t v = lnone;
                pe[j] = pw[j] = false;
                if (kk != 8) {
                    if (exits) v = hist_sa + 4u * (uint32_t)slot_of(lr, lc, th, tw);
                    else if (kk == 2) v = lb[j];
                    else if (kk == 1) v = (j < NC - 1) ? lb[j < NC - 1 ? j + 1 : 0] : lb_right0;
                    else if (kk == 3) v = (j > 0) ? lb[j > 0 ? j - 1 : 0] : lb_left;
                    else if (kk == 0) pe[j] = true;
                    else if (kk == 4) pw[j] = true;
                }
                own[j] = v;
            }
            uint32_t lab[NC];
            {   // East cells copy the label of the chain end on their right
                uint32_t V = lnone;
                bool P = true;
#pragma unroll
                for (int j = NC - 1; j >= 0; j--) { V = pe[j] ? V : own[j]; P = P && pe[j]; }
                uint32_t cur = label_carry_from_right(V, P, lane, lnone);
#pragma unroll
                for (int j = NC - 1; j >= 0; j--) { cur = pe[j] ? cur : own[j]; lab[j] = cur; }
            }
            bool anyw = false;
#pragma unroll
            for (int j = 0; j < NC; j++) anyw = anyw || pw[j];
            if (__any_sync(FULL, anyw)) {
                uint32_t V = lnone;
                bool P = true;
#pragma unroll
                for (int j = 0; j < NC; j++) { V = pw[j] ? V : lab[j]; P = P && pw[j]; }
                uint32_t cur = label_carry_from_left(V, P, lane, lnone);
#pragma unroll
                for (int j = 0; j < NC; j++) { cur = pw[j] ? cur : lab[j]; lab[j] = cur; }
            }
            uint32_t mw[NC];
#pragma unroll
            for (int j = 0; j < NC; j++) {
                lb[j] = lab[j];
                mw[j] = meta_pack(slot_label(lab[j]), k[j], false, (am >> j) & 1);
            }
            if (lr == 0 || lr == th - 1) store_meta_row(lr == 0 ? 0 : TW, mw);
            else {
                if (lane == 0) m_t[2 * TW + lr] = mw[0];
                if (lane == rlane && tw > 1) {
#pragma unroll
                    for (int j = 0; j < NC; j++) if (j == rj) m_t[2 * TW + TH + lr] = mw[j];
                }
            }
            count_labels(lab);
        }
    }
    rows.drain(c_left, nch - 1 - c_left, WMF_BAR_A);
#undef WMF_BAR_A
    if (complex_tile) {                              // the general solver redoes this tile
        if (lane == 0) { tile_class[tile] = 1; tile_top[tile] = 1; raise_flag(flags, has_multi ? F_COMPLEX : F_CHAIN); }
        return;
    }
    // tile_top: can an inner first-row slot of this tile be an exit?  Only through upward flow in row 0 (never in a
    // lean row) or in a tile that is not full size (one row: first row = last row).  The forest kernels skip the
    // first-row slots 1 .. TW-2 of the other tiles without touching their records; the corners can leave sideways.
    if (lane == 0) { tile_class[tile] = 0; row_plain[tile] = plain_rows; tile_top[tile] = (top_exits || !full) ? 1 : 0; }
    // exit_acc of every slot = its bin (slots that are not exits hold whatever drained to them: unused), and the
    // forest word of the slot: one pending arrival (its own walker) | its weight
#pragma unroll
    for (int q = 0; q < NSLOT / 128; q++) {
        const int4 h = *reinterpret_cast<const int4 *>(hist + 128 * q + 4 * lane);
        if (exit_acc) *reinterpret_cast<int4 *>(x_t + 128 * q + 4 * lane) = h;     // (only the 64-bit forest engine reads it)
        // (the slots' meta words, written by this warp a moment ago, come back from L2: their code goes into the word)
        const uint4 mq = __ldcg(reinterpret_cast<const uint4 *>(m_t + 128 * q + 4 * lane));
        auto word = [&](uint32_t m, int hh) { return BIAS | ((unsigned long long)((m >> 16) & 0xFu) << W_K) | (uint32_t)hh; };
        ulonglong2 *wp2 = reinterpret_cast<ulonglong2 *>(w_t + 128 * q + 4 * lane);
        wp2[0] = make_ulonglong2(word(mq.x, h.x), word(mq.y, h.y));
        wp2[1] = make_ulonglong2(word(mq.z, h.z), word(mq.w, h.w));
    }
}

// ---------------------------------------------------------------------------------------------
// phase C, fast solver
// ---------------------------------------------------------------------------------------------
// value of an accumulation computed in T, stored as TO: a 32-bit accumulation of a raster with fewer than 2^32 cells
// never wraps as an unsigned number, so it may be stored into a wider type (int64 / float64) without 64-bit arithmetic
template <typename TO, typename T>
__device__ __forceinline__ TO acc_out(T v)
{
    if (sizeof(T) == 4 && sizeof(TO) == 8) return (TO)(uint32_t)v;
    return (TO)v;
}

// 32-byte stores (STG.256, sm_100): a lane's piece of a row is whole 32-byte sectors, one instruction each
__device__ __forceinline__ void st256(void *p, unsigned long long a, unsigned long long b, unsigned long long c, unsigned long long d)
{
    asm volatile("st.global.v4.b64 [%0], {%1, %2, %3, %4};" ::"l"(p), "l"(a), "l"(b), "l"(c), "l"(d) : "memory");
}

template <bool AL, typename T, typename TO>
__device__ __forceinline__ void store_row(TO *out, const T (&tot)[NC], int64_t gc0, int64_t nx)
{
    static_assert(NC == 8, "a lane stores 8 cells");
    if (AL) {
        if (gc0 < nx) {
            if (sizeof(TO) == 4) {
                unsigned long long w[4];
#pragma unroll
                for (int q = 0; q < 4; q++) w[q] = (unsigned long long)(uint32_t)tot[2 * q] | ((unsigned long long)(uint32_t)tot[2 * q + 1] << 32);
                st256(out, w[0], w[1], w[2], w[3]);
            } else {
                unsigned long long w[8];
#pragma unroll
                for (int j = 0; j < 8; j++) { TO a = acc_out<TO, T>(tot[j]); memcpy(&w[j], &a, 8); }
                st256(out, w[0], w[1], w[2], w[3]);
                st256(out + 4, w[4], w[5], w[6], w[7]);
            }
        }
    } else {
#pragma unroll
        for (int j = 0; j < NC; j++) if (gc0 + j < nx) out[j] = acc_out<TO, T>(tot[j]);
    }
}

// The external inflow of the tile's 640 boundary slots, in shared memory (slot order): pulled from the packed forest
// words of the neighbouring tiles (wq; int32 accumulations), plus an optional per-slot array `extra` (what the other row
// stripes send in; or, for the 64-bit forest engine, everything).
template <typename T>
__device__ __forceinline__ void tile_inflow(const Geo &g, const unsigned long long *__restrict__ wq, const uint8_t *__restrict__ tile_top,
                                            const T *__restrict__ extra, int tile, int ty, int tx, int th, int tw, int lane, T *ext)
{
    if (wq) {
        uint32_t *e32 = reinterpret_cast<uint32_t *>(ext);
        if (th == TH && tw == TW) {
            pull_full_tile(g, wq, tile_top, tile, ty, tx, lane, e32);
        } else {
            const TileGeo tg = tile_geo(g);
            bool un = false;                                // (tiles with an unresolved feeder never come here: class 2)
            for (int s = lane; s < NSLOT; s += 32) e32[s] = pull_slot(g, tg, wq, ty, tx, th, tw, s, un, [](int64_t) { return false; });
            __syncwarp();
        }
        if (sizeof(T) == 8) {                               // widen in place, last slots first (a lane's reads precede its writes
            for (int q = NSLOT / 32 - 1; q >= 0; q--) {     // of higher addresses; lanes are in step)
                const uint32_t v = e32[32 * q + lane];
                __syncwarp();
                ext[32 * q + lane] = (T)v;
                __syncwarp();
            }
        }
    } else {
        for (int s = lane; s < NSLOT; s += 32) ext[s] = 0;
    }
    // (extra_edge: only the stripe's own first / last tile row holds anything -- the rest went into the forest words)
    if (extra && (!g.extra_edge || ty == 0 || ty == g.tiles_y - 1)) {
        const T *x = extra + (int64_t)tile * NSLOT;
        for (int s = lane; s < NSLOT; s += 32) ext[s] += x[s];
    }
    __syncwarp();
}

template <int ENC, bool AL, bool MK, bool TM, typename T, typename TO>
__global__ void __launch_bounds__(WPB * 32, sizeof(T) == 8 ? 4 : 6) k_sweepC(const int8_t *__restrict__ dir, const uint8_t *__restrict__ mask,
                                                     Geo g, int ntiles, const unsigned long long *__restrict__ wq,
                                                     const uint8_t *__restrict__ tile_top, const T *__restrict__ extra,
                                                     const uint8_t *__restrict__ tile_class,
                                                     const unsigned long long *__restrict__ row_plain, TO *__restrict__ acc)
{
    extern __shared__ __align__(16) uint32_t s_dyn[];   // [WPB][2][CH][ROWW] DIR words (+ mask words) + the boundary inflows
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tile = blockIdx.x * WPB + warp;
    if (tile >= ntiles) return;
    if (tile_class[tile] != 0) return;              // general solver
    int ty, tx, th, tw;
    tile_dims(g, tile, ty, tx, th, tw);
    const int64_t gr0 = (int64_t)ty * TH, gc0 = (int64_t)tx * TW + NC * lane;
    const int rlane = (tw - 1) / NC, rj = (tw - 1) % NC;
    const bool border = ty == 0 || tx == 0 || ty == g.tiles_y - 1 || tx == g.tiles_x - 1;
    // (TM: the warp's two mbarriers sit behind the per-chunk lean-row masks, at the end of the block's shared memory)
#define WMF_BAR_C ((uint32_t)__cvta_generic_to_shared(s_dyn + (size_t)(mask ? 2 : 1) * WPB * 2 * CH * ROWW) + (uint32_t)(WPB * NSLOT * sizeof(T) + WPB * 8) + 16u * (threadIdx.x >> 5))
    TileRows<AL, TM> rows;
    rows.init(s_dyn, warp, lane, dir, mask, g.nx, gr0, gc0, th, WMF_BAR_C);
    const int nch = (th + CH - 1) / CH;
    rows.prefetch(0, WMF_BAR_C);
    // external inflow of the boundary cells: first row, last row, first column, last column
    T *ext = reinterpret_cast<T *>(s_dyn + (mask ? 2 : 1) * WPB * 2 * CH * ROWW) + NSLOT * warp;
    const T *extL = ext + 2 * TW, *extR = ext + 2 * TW + TH;
    tile_inflow<T>(g, wq, tile_top, extra, tile, ty, tx, th, tw, lane, ext);
    T d[NC];
#pragma unroll
    for (int j = 0; j < NC; j++) d[j] = 1;           // 1 + inflow from the row above
    bool dummy = false;
    TO *out = acc + gr0 * g.nx + gc0;
    // the per-chunk lean-row masks live in shared memory: an L1TEX load (global or a register spill) issued
    // behind the chunk prefetch would wait for the prefetch's HBM round trip (the L1TEX queue is in order)
    unsigned char *pchunk = reinterpret_cast<unsigned char *>(ext + (size_t)NSLOT * (WPB - warp)) + 8 * warp;
    if (lane == 0) *reinterpret_cast<unsigned long long *>(pchunk) = row_plain[tile];
    const bool is_l = lane == 0, is_r = lane == 31;
    for (int c = 0; c < nch; c++) {
        rows.arrive(c, c, c + 1, nch, WMF_BAR_C);
        const int cend = th < (c + 1) * CH ? th : (c + 1) * CH;
        int lr = c * CH;
        const uint32_t *wp = rows.chunk(c);              // row lr of this chunk at wp
        uint32_t pbits = pchunk[c];
        while (lr < cend) {
            // a run of lean rows (bounded by the chunk): no per-row classification, no divergent control flow
            const uint32_t inv = ~pbits & ((1u << CH) - 1u);
            int run = inv ? __ffs(inv) - 1 : CH;
            run = run < cend - lr ? run : cend - lr;
            pbits >>= run;
            for (const int end = lr + run; lr < end; lr++, out += g.nx, wp += ROWW) {
                T tot[NC];
                uint32_t kw[NWD], dv[NWD], av[NWD];
                if (MK) lean_words_mk<ENC>(wp, rows.mchunk(c) + (lr - c * CH) * ROWW, kw, dv, av);
                else plain_words<ENC>(wp, kw);
                T l = 0, r = 0;
                if (lr == 0 || lr == TH - 1) {            // first / last row: every cell is a boundary slot
                    const T *e4 = ext + (lr == 0 ? 0 : TW) + NC * lane;
#pragma unroll
                    for (int j = 0; j < NC; j++) d[j] += e4[j];
                } else {
                    l = is_l ? extL[lr] : (T)0; r = is_r ? extR[lr] : (T)0;
                }
                if (MK) plain_row_down_mk<T>(kw, dv, av, d, tot, lane, l, r);
                else plain_row_down<T>(kw, d, tot, lane, l, r);
                store_row<AL, T, TO>(out, tot, gc0, g.nx);
            }
            if (lr >= cend) break;
            {   // one generic row
                T tot[NC];
                int k[NC];
                const int am = rows.active(lr);
                decode_row<ENC>(wp, am, g, gr0 + lr, gc0, border, k);
                T ex[NC];
#pragma unroll
                for (int j = 0; j < NC; j++) ex[j] = 0;
                if (lr == 0 || lr == th - 1) {
                    const T *e4 = ext + (lr == 0 ? 0 : TW) + NC * lane;
#pragma unroll
                    for (int j = 0; j < NC; j++) ex[j] = e4[j];
                } else {
                    if (lane == 0) ex[0] = extL[lr];
                    if (lane == rlane && tw > 1) {
#pragma unroll
                        for (int j = 0; j < NC; j++) if (j == rj) ex[j] += extR[lr];
                    }
                }
                T b[NC];
#pragma unroll
                for (int j = 0; j < NC; j++) b[j] = ((am >> j) & 1) ? d[j] + ex[j] : (T)0;
                sweep_row_down<T>(k, am, b, tot, d, lane, dummy);
#pragma unroll
                for (int j = 0; j < NC; j++) d[j] += 1;
                store_row<AL, T, TO>(out, tot, gc0, g.nx);
                lr++; out += g.nx; wp += ROWW; pbits >>= 1;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// general tile solver (one block per tile, chain walk in shared memory), phases A and C
// ---------------------------------------------------------------------------------------------
constexpr int GTH = 256;   // threads per block

template <int ENC>
__device__ __forceinline__ void general_load(const int8_t *__restrict__ dir, const uint8_t *__restrict__ mask, const Geo &g,
                                             int ty, int tx, int th, int tw, uint8_t *kk)
{
    for (int idx = threadIdx.x; idx < TCELLS; idx += GTH) {
        int lr = idx / TW, lc = idx - lr * TW;
        uint8_t v = 8;
        if (lr < th && lc < tw) {
            int64_t gr = (int64_t)ty * TH + lr, gc = (int64_t)tx * TW + lc, gi = gr * g.nx + gc;
            if (cell_active(mask, gi)) {
                int k = d8_norm<ENC>(dir[gi]);
                if (k != 8) {
                    int64_t rr = g.gy0 + gr + d8_dr(k), cc = gc + d8_dc(k);
                    if (rr < 0 || rr >= g.gny || cc < 0 || cc >= g.nx) k = 8;
                }
                v = (uint8_t)(k | 16);
            }
        }
        kk[idx] = v;
    }
}

// in-tile receiver of cell idx: >= 0 index, -1 none / inactive receiver, -2 leaves the tile
__device__ __forceinline__ int local_recv(const uint8_t *kk, int idx, int th, int tw)
{
    int k = kk[idx] & 15;
    if (k == 8) return -1;
    int lr = idx / TW, lc = idx - lr * TW;
    int rr = lr + d8_dr(k), cc = lc + d8_dc(k);
    if (rr < 0 || rr >= th || cc < 0 || cc >= tw) return -2;
    int j = rr * TW + cc;
    return (kk[j] & 16) ? j : -1;
}

__device__ __forceinline__ void general_indegree(const uint8_t *kk, uint8_t *cnt, int th, int tw)
{
    for (int idx = threadIdx.x; idx < TCELLS; idx += GTH) {
        uint8_t c = 0;
        if (kk[idx] & 16) {
            int lr = idx / TW, lc = idx - lr * TW, n = 0;
#pragma unroll
            for (int k = 0; k < 8; k++) {
                int rr = lr - d8_dr(k), cc = lc - d8_dc(k);
                if (rr < 0 || rr >= th || cc < 0 || cc >= tw) continue;
                uint8_t v = kk[rr * TW + cc];
                if ((v & 16) && (v & 15) == k) n++;
            }
            c = n == 0 ? 0xFF : (uint8_t)n;
        }
        cnt[idx] = c;
    }
}

template <typename T>
__device__ __forceinline__ T smem_add(T *p, T v);
template <>
__device__ __forceinline__ int32_t smem_add<int32_t>(int32_t *p, int32_t v) { return atomicAdd(p, v); }
template <>
__device__ __forceinline__ int64_t smem_add<int64_t>(int64_t *p, int64_t v)
{
    return (int64_t)atomicAdd((unsigned long long *)p, (unsigned long long)v);
}

template <typename T>
__device__ __forceinline__ void general_walk(const uint8_t *kk, uint8_t *cnt, T *acc, int th, int tw)
{
    uint32_t *cnt32 = reinterpret_cast<uint32_t *>(cnt);
    for (int idx = threadIdx.x; idx < TCELLS; idx += GTH) {
        if (cnt[idx] != 0xFF) continue;                  // leaves only; 0xFF is never produced by a decrement
        T v = acc[idx] + 1;                              // acc holds the external inflow (0 in phase A)
        acc[idx] = v;
        int cur = idx;
        while (true) {
            int j = local_recv(kk, cur, th, tw);
            if (j < 0) break;
            smem_add<T>(&acc[j], v);
            __threadfence_block();
            unsigned sh = 8u * (unsigned)(j & 3);
            uint32_t old = atomicSub(&cnt32[j >> 2], 1u << sh);
            if (((old >> sh) & 0xFFu) != 1u) break;
            __threadfence_block();
            v = smem_add<T>(&acc[j], (T)1) + 1;
            cur = j;
        }
    }
}

template <int ENC>
__global__ void __launch_bounds__(GTH) k_generalA(const int8_t *__restrict__ dir, const uint8_t *__restrict__ mask, Geo g, int ntiles,
                                                  const uint8_t *__restrict__ tile_class, uint32_t *__restrict__ meta,
                                                  int32_t *__restrict__ exit_acc, unsigned long long *__restrict__ wq,
                                                  const uint32_t *__restrict__ flags)
{
    extern __shared__ __align__(16) unsigned char smem[];
  if (flags[F_CHAIN] == 0u) return;
  // each block owns a contiguous range of tiles and looks at all their classes at once: nearly always
  // nothing is complex and the block leaves after one coalesced load and one barrier
  const int per = (ntiles + gridDim.x - 1) / gridDim.x, t_lo = blockIdx.x * per, t_hi = ntiles < t_lo + per ? ntiles : t_lo + per;
  bool any_mine = false;
  for (int t = t_lo + threadIdx.x; t < t_hi; t += GTH) any_mine = any_mine || (tile_class[t] & 1);
  if (!__syncthreads_or(any_mine)) return;
  for (int tile = t_lo; tile < t_hi; tile++) {
    if (!(tile_class[tile] & 1)) continue;
    __syncthreads();                                   // previous tile of this block is done with smem
    int32_t *acc = reinterpret_cast<int32_t *>(smem);
    uint8_t *kk = smem + sizeof(int32_t) * TCELLS;
    uint8_t *cnt = kk + TCELLS;
    int ty, tx, th, tw;
    tile_dims(g, tile, ty, tx, th, tw);
    general_load<ENC>(dir, mask, g, ty, tx, th, tw, kk);
    for (int idx = threadIdx.x; idx < TCELLS; idx += GTH) acc[idx] = 0;
    __syncthreads();
    general_indegree(kk, cnt, th, tw);
    __syncthreads();
    general_walk<int32_t>(kk, cnt, acc, th, tw);
    __syncthreads();
    for (int s = threadIdx.x; s < NSLOT; s += GTH) {
        int lr, lc;
        uint32_t m = 0;
        int32_t xa = 0;
        if (cell_of(s, th, tw, lr, lc)) {
            int idx = lr * TW + lc;
            uint8_t v = kk[idx];
            if (v & 16) {
                bool unres = cnt[idx] != 0xFF && cnt[idx] != 0;
                uint32_t link = LNONE;
                int cur = idx;
                for (int steps = 0; steps <= th * tw; steps++) {
                    int j = local_recv(kk, cur, th, tw);
                    if (j == -2) { link = (uint32_t)slot_of(cur / TW, cur % TW, th, tw); break; }
                    if (j < 0) break;
                    cur = j;
                }
                m = meta_pack(link, v & 15, unres, true);
                xa = acc[idx];
            }
        }
        meta[(int64_t)tile * NSLOT + s] = m;
        if (exit_acc) exit_acc[(int64_t)tile * NSLOT + s] = xa;
        // pending | code | weight; a locally unresolved exit never finalises
        wq[(int64_t)tile * NSLOT + s] = ((1ull + ((m >> 20) & 1u)) << 32) | ((unsigned long long)((m >> 16) & 0xFu) << W_K) | (uint32_t)xa;
    }
  }
}

template <int ENC, typename T, typename TO>
__global__ void __launch_bounds__(GTH) k_generalC(const int8_t *__restrict__ dir, const uint8_t *__restrict__ mask, Geo g, int ntiles,
                                                  const uint8_t *__restrict__ tile_class, const unsigned long long *__restrict__ wq,
                                                  const int32_t *__restrict__ par, const T *__restrict__ extra,
                                                  const uint8_t *__restrict__ extra_unres, TO *__restrict__ out,
                                                  const uint32_t *__restrict__ flags)
{
    extern __shared__ __align__(16) unsigned char smem[];
  if (flags[F_CHAIN] == 0u && flags[F_UNRES] == 0u) return;
  const int per = (ntiles + gridDim.x - 1) / gridDim.x, t_lo = blockIdx.x * per, t_hi = ntiles < t_lo + per ? ntiles : t_lo + per;
  bool any_mine = false;
  // classes: 0 one-pass sweeps, 4 jump solver; everything else (1 / 3: the sweeps cannot solve it, 2: fed by an unresolved
  // upstream) is solved here
  for (int t = t_lo + threadIdx.x; t < t_hi; t += GTH) any_mine = any_mine || (tile_class[t] != 0 && tile_class[t] != 4);
  if (!__syncthreads_or(any_mine)) return;
  const TileGeo tg = tile_geo(g);
  for (int tile = t_lo; tile < t_hi; tile++) {
    if (tile_class[tile] == 0 || tile_class[tile] == 4) continue;
    __syncthreads();
    T *acc = reinterpret_cast<T *>(smem);
    uint8_t *kk = smem + sizeof(T) * TCELLS;
    uint8_t *cnt = kk + TCELLS;
    int ty, tx, th, tw;
    tile_dims(g, tile, ty, tx, th, tw);
    general_load<ENC>(dir, mask, g, ty, tx, th, tw, kk);
    for (int idx = threadIdx.x; idx < TCELLS; idx += GTH) acc[idx] = 0;
    __syncthreads();
    general_indegree(kk, cnt, th, tw);
    __syncthreads();
    for (int s = threadIdx.x; s < NSLOT; s += GTH) {
        int lr, lc;
        if (!cell_of(s, th, tw, lr, lc)) continue;
        int idx = lr * TW + lc;
        if (!(kk[idx] & 16)) continue;
        bool un = false;
        T v = wq ? (T)pull_slot(g, tg, wq, ty, tx, th, tw, s, un, [&](int64_t node) { return par[node] != PAR_NONE; }) : (T)0;
        if (extra && (!g.extra_edge || ty == 0 || ty == g.tiles_y - 1)) v += extra[(int64_t)tile * NSLOT + s];
        if (extra_unres && extra_unres[(int64_t)tile * NSLOT + s]) un = true;
        acc[idx] = v;
        if (un) cnt[idx] = cnt[idx] == 0xFF ? 1 : cnt[idx] + 1;   // phantom donor: an unresolved upstream never arrives
    }
    __syncthreads();
    general_walk<T>(kk, cnt, acc, th, tw);
    __syncthreads();
    for (int idx = threadIdx.x; idx < TCELLS; idx += GTH) {
        int lr = idx / TW, lc = idx - lr * TW;
        if (lr < th && lc < tw) out[((int64_t)ty * TH + lr) * g.nx + (int64_t)tx * TW + lc] = acc_out<TO, T>(acc[idx]);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// jump solver: tiles with in-tile upward flow (every tile of a real DEM)
// ---------------------------------------------------------------------------------------------
// A row sweep settles a path only as long as it runs with the sweep, so terrain needs a sweep per vertical run of its
// paths (4-5 on average at this tile height, each as dear as the whole one-pass solve).  These tiles are solved by
// pointer jumping over the tile's 16384 cells in shared memory instead, one thread block per tile:
//   phase A  ptr[c] = the receiver of c, or -- resolved -- the exit label of c (an exit cell labels itself, a cell
//            without receiver carries "no exit").  ptr[c] = ptr[ptr[c]] until every pointer is a label; done in place
//            and asynchronously, which only makes the pointers run ahead.  log2(longest path) rounds; exit_acc is the
//            histogram of the labels.  Pointers that are still unresolved after 2^JROUNDS >= 16384 steps sit on a
//            cycle: the tile goes to the chain-walk solver.
//   phase C  subtree sums by doubling: with A_k[c] = what lies upstream of c at a distance below 2^k (A_0 = the cell's
//            own 1 + its external inflow) and ptr_k = the 2^k-th receiver, A_{k+1}[c] = A_k[c] + the sum of A_k[u] over
//            the cells u with ptr_k[u] = c -- a scatter-add of every cell that still has a pointer, then ptr_{k+1} =
//            ptr_k[ptr_k].  Rounds are synchronous here (pointers are double buffered); a cell drops out once its
//            pointer runs off its path.
// The tile's decoded codes (a nibble per cell: 0..7, 8 none, 9 inactive) are handed from phase A to phase C.
constexpr int JTA = 512, JTC = 1024;        // threads per block, phases A / C
constexpr int JROUNDS = 15;
constexpr uint32_t JRES = 0x8000u;          // phase A: the pointer is resolved, the low bits are the label
constexpr uint32_t JNONE = 0xFFFFu;         // phase C: no (further) receiver inside the tile
static_assert(TCELLS <= 0x4000, "a cell index fits below the flag bits of a 16-bit pointer");

// four internal-coded cells (one byte each) and their mask bytes -> four nibbles: the code, 8 for anything that is not
// a direction, 9 for an inactive cell
__device__ __forceinline__ uint32_t jump_nibbles_internal(uint32_t w, uint32_t m)
{
    const uint32_t bad = byte_nonzero(w & 0xF8F8F8F8u), off = byte_zero(m);       // 0x01 per byte
    uint32_t x = (w & 0x07070707u & ~(bad * 0xFFu)) | (bad << 3);                 // code, or 8
    x = (x & ~(off * 0xFFu)) | (off * 9u);                                        // ... or 9
    x = (x | (x >> 4)) & 0x00FF00FFu;
    return (x | (x >> 8)) & 0xFFFFu;
}

// decode the whole tile into nibbles: codes[lr * 32 + l] holds the cells 8 l .. 8 l + 7 of row lr (cells beyond a ragged
// tile's extent are inactive).  vec: rows can be fetched eight bytes at a time.
template <int ENC>
__device__ __forceinline__ void jump_decode(const int8_t *__restrict__ dir, const uint8_t *__restrict__ mask, const Geo &g, int ty, int tx,
                                            int th, bool vec, uint32_t *codes, uint32_t *codes_out)
{
    const int64_t gr0 = (int64_t)ty * TH, gc0 = (int64_t)tx * TW;
    const bool border = ty == 0 || tx == 0 || ty == g.tiles_y - 1 || tx == g.tiles_x - 1;
    for (int q = threadIdx.x; q < TH * 32; q += blockDim.x) {
        const int lr = q >> 5, l = q & 31;
        const int64_t gr = gr0 + lr, gc = gc0 + NC * l;
        uint32_t cw = 0x99999999u;
        if (lr < th && gc < g.nx) {
            uint32_t dw[2] = {0u, 0u}, mw[2] = {0x01010101u, 0x01010101u};
            const int64_t gi = gr * g.nx + gc;
            if (vec) {
                const uint2 v = __ldg(reinterpret_cast<const uint2 *>(dir + gi));
                dw[0] = v.x; dw[1] = v.y;
                if (mask) { const uint2 m = __ldg(reinterpret_cast<const uint2 *>(mask + gi)); mw[0] = m.x; mw[1] = m.y; }
            } else {
#pragma unroll
                for (int j = 0; j < NC; j++) {
                    const bool in = gc + j < g.nx;
                    const uint32_t d = in ? (uint32_t)(uint8_t)dir[gi + j] : 0u, m = in ? (mask ? (uint32_t)(mask[gi + j] != 0) : 1u) : 0u;
                    dw[j / 4] |= d << (8 * (j % 4));
                    mw[j / 4] = (mw[j / 4] & ~(0xFFu << (8 * (j % 4)))) | (m << (8 * (j % 4)));
                }
            }
            if (ENC == WMF_ENC_INTERNAL && !border) {
                cw = jump_nibbles_internal(dw[0], mw[0]) | (jump_nibbles_internal(dw[1], mw[1]) << 16);
            } else {
                cw = 0;
#pragma unroll
                for (int j = 0; j < NC; j++) {
                    int kk = d8_norm<ENC>((int)(int8_t)(dw[j / 4] >> (8 * (j % 4))));
                    if (border && kk != 8) {
                        const int64_t rr = g.gy0 + gr + d8_dr(kk), cc = gc + j + d8_dc(kk);
                        if (rr < 0 || rr >= g.gny || cc < 0 || cc >= g.nx) kk = 8;
                    }
                    if (((mw[j / 4] >> (8 * (j % 4))) & 0xFFu) == 0u) kk = 9;
                    cw |= (uint32_t)kk << (4 * j);
                }
            }
        }
        codes[q] = cw;
        if (codes_out) codes_out[q] = cw;
    }
}

// in-tile receiver of cell c with code kr (0..7), for a cell that is not on the tile's boundary: c + dr * TW + dc
__device__ __forceinline__ uint32_t jump_recv(uint32_t c, uint32_t kr)
{
    const uint32_t dr1 = (0x01A9u >> (2u * kr)) & 3u, dc1 = (0x901Au >> (2u * kr)) & 3u;      // dr + 1, dc + 1 (d8_dr / d8_dc)
    return c - (uint32_t)(TW + 1) + dr1 * (uint32_t)TW + dc1;
}
__device__ __forceinline__ uint32_t jump_code(const uint32_t *codes, int lr, int lc) { return (codes[lr * 32 + lc / NC] >> (4 * (lc % NC))) & 15u; }

// The pointers of a whole tile, eight cells per step.  Cells without a direction get `none`.  On a full tile every chunk
// takes the interior formula, which is wrong for the boundary cells only: the caller puts those right (one thread per
// boundary slot) -- a per-chunk test would split every warp, lanes 0 and 31 hold the side columns.  out_of_tile(lr, lc)
// gives the pointer of a cell whose receiver lies outside the tile.
template <typename OutFn>
__device__ __forceinline__ void jump_pointers(const uint32_t *codes, int th, int tw, uint32_t none, uint32_t *ptr32, OutFn out_of_tile)
{
    const bool full = th == TH && tw == TW;
    for (int q = threadIdx.x; q < TH * 32; q += blockDim.x) {
        const uint32_t cw = codes[q];
        uint32_t pv[NC];
        if (full) {
#pragma unroll
            for (int j = 0; j < NC; j++) {
                const uint32_t kr = (cw >> (4 * j)) & 15u;
                pv[j] = (kr < 8u ? jump_recv((uint32_t)(NC * q + j), kr) : none) & 0xFFFFu;
            }
        } else {
            const int lr = q >> 5, l = q & 31;
#pragma unroll
            for (int j = 0; j < NC; j++) {
                const int kr = (int)((cw >> (4 * j)) & 15u), lc = NC * l + j;
                uint32_t v = none;
                if (kr < 8) {
                    const int rr = lr + d8_dr(kr), cc = lc + d8_dc(kr);
                    v = (rr < 0 || rr >= th || cc < 0 || cc >= tw) ? out_of_tile(lr, lc) : (uint32_t)(rr * TW + cc);
                }
                pv[j] = v;
            }
        }
        *reinterpret_cast<uint4 *>(ptr32 + 4 * q) = make_uint4(pv[0] | (pv[1] << 16), pv[2] | (pv[3] << 16), pv[4] | (pv[5] << 16), pv[6] | (pv[7] << 16));
    }
}
// ... of one boundary cell of a full tile
template <typename OutFn>
__device__ __forceinline__ uint32_t jump_pointer_boundary(uint32_t kr, int lr, int lc, uint32_t none, OutFn out_of_tile)
{
    if (kr >= 8u) return none;
    const int rr = lr + d8_dr((int)kr), cc = lc + d8_dc((int)kr);
    return (rr < 0 || rr >= TH || cc < 0 || cc >= TW) ? out_of_tile(lr, lc) : (uint32_t)(rr * TW + cc);
}

// next tile of a persistent block (thread 0 draws, everybody learns through shared memory)
__device__ __forceinline__ int jump_next_tile(uint32_t *counter, int *slot)
{
    __syncthreads();                                   // (the previous tile is done with shared memory, *slot included)
    if (threadIdx.x == 0) *slot = (int)atomicAdd(counter, 1u);
    __syncthreads();
    return *slot;
}

template <int ENC>
__global__ void __launch_bounds__(JTA, 4) k_jumpA(const int8_t *__restrict__ dir, const uint8_t *__restrict__ mask, Geo g, int ntiles,
                                               uint32_t *__restrict__ meta, int32_t *__restrict__ exit_acc,
                                               uint8_t *__restrict__ tile_class, unsigned long long *__restrict__ wq,
                                               uint8_t *__restrict__ tile_top, uint32_t *__restrict__ codes4,
                                               unsigned long long *__restrict__ rounds_of, uint32_t *__restrict__ flags, bool vec)
{
    extern __shared__ __align__(16) uint32_t s_dyn[];   // [TCELLS / 2] pointers (16 bit) + [TH * 32] codes + [HBINS] histogram + 1
    if (flags[F_COMPLEX] == 0u) return;               // (nearly always: no tile has upward flow, e.g. a Scheidegger network)
    uint32_t *ptr32 = s_dyn;
    uint16_t *ptr16 = reinterpret_cast<uint16_t *>(s_dyn);
    uint32_t *codes = s_dyn + TCELLS / 2;
    int *hist = reinterpret_cast<int *>(codes + TH * 32);
    int *slot = hist + HBINS;
    for (;;) {
        const int tile = jump_next_tile(&flags[F_NEXT_A], slot);
        if (tile >= ntiles) return;
        if (tile_class[tile] != 1) continue;
        int ty, tx, th, tw;
        tile_dims(g, tile, ty, tx, th, tw);
        jump_decode<ENC>(dir, mask, g, ty, tx, th, vec, codes, codes4 + (int64_t)tile * TH * 32);   // (phase C starts from the codes)
        for (int b = threadIdx.x; b < HBINS; b += JTA) hist[b] = 0;
        __syncthreads();
        auto exit_label = [&](int lr, int lc) { return JRES | (uint32_t)slot_of(lr, lc, th, tw); };
        jump_pointers(codes, th, tw, JRES | LNONE, ptr32, exit_label);
        const bool full = th == TH && tw == TW;
        if (full) {
            __syncthreads();
            for (int s = threadIdx.x; s < NSLOT; s += JTA) {
                int lr, lc;
                cell_of(s, TH, TW, lr, lc);
                ptr16[lr * TW + lc] = (uint16_t)jump_pointer_boundary(jump_code(codes, lr, lc), lr, lc, JRES | LNONE, exit_label);
            }
        }
        bool up0 = false;                                // upward flow in the first row?
        for (int l = threadIdx.x; l < 32; l += JTA) {
            const uint32_t cw = codes[l];
#pragma unroll
            for (int j = 0; j < NC; j++) { const uint32_t kr = (cw >> (4 * j)) & 15u; up0 = up0 || (kr >= 5u && kr <= 7u); }
        }
        const bool top_up = __syncthreads_or(up0) != 0;
        int rounds = 0;
        bool open = true;
        for (; open && rounds < JROUNDS; rounds++) {
            // four cells per step, no branches: a resolved cell reads the pointer of cell `label` (< 1024: harmless) and keeps its own
            uint32_t all = 0x80008000u;
            uint2 *ptr64 = reinterpret_cast<uint2 *>(ptr32);
            for (int i = threadIdx.x; i < TCELLS / 4; i += JTA) {
                uint2 w = ptr64[i];
                if ((w.x & w.y & 0x80008000u) == 0x80008000u) continue;
                const uint32_t a0 = w.x & 0xFFFFu, a1 = w.x >> 16, a2 = w.y & 0xFFFFu, a3 = w.y >> 16;
                const uint32_t b0 = ptr16[a0 & 0x3FFFu], b1 = ptr16[a1 & 0x3FFFu], b2 = ptr16[a2 & 0x3FFFu], b3 = ptr16[a3 & 0x3FFFu];
                w.x = ((a0 & JRES) ? a0 : b0) | (((a1 & JRES) ? a1 : b1) << 16);
                w.y = ((a2 & JRES) ? a2 : b2) | (((a3 & JRES) ? a3 : b3) << 16);
                ptr64[i] = w;
                all &= w.x & w.y;
            }
            open = __syncthreads_or(all != 0x80008000u) != 0;
        }
        if (open) {                                      // a cycle: the chain-walk solver redoes this tile
            if (threadIdx.x == 0) { tile_class[tile] = 3; tile_top[tile] = 1; raise_flag(flags, F_CHAIN); }
            continue;
        }
        // exit_acc: the histogram of the labels (runs of equal labels along a row are counted at once)
        for (int q = threadIdx.x; q < TH * 32; q += JTA) {
            const uint4 x = *reinterpret_cast<const uint4 *>(ptr32 + 4 * q);
            const uint32_t lab[NC] = {x.x & 0x3FFu, (x.x >> 16) & 0x3FFu, x.y & 0x3FFu, (x.y >> 16) & 0x3FFu,
                                      x.z & 0x3FFu, (x.z >> 16) & 0x3FFu, x.w & 0x3FFu, (x.w >> 16) & 0x3FFu};
            uint32_t cur = lab[0];
            int n = 1;
#pragma unroll
            for (int j = 1; j < NC; j++) {
                if (lab[j] == cur) n++;
                else { atomicAdd(&hist[cur], n); cur = lab[j]; n = 1; }
            }
            atomicAdd(&hist[cur], n);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            tile_class[tile] = 4;
            tile_top[tile] = (top_up || !full) ? 1 : 0;
            rounds_of[tile] = (unsigned long long)rounds;
        }
        // the boundary slots: link = the cell's label, the forest word = one pending arrival | code | weight
        uint32_t *m_t = meta + (int64_t)tile * NSLOT;
        int32_t *x_t = exit_acc ? exit_acc + (int64_t)tile * NSLOT : nullptr;
        unsigned long long *w_t = wq + (int64_t)tile * NSLOT;
        for (int s = threadIdx.x; s < NSLOT; s += JTA) {
            int lr, lc;
            uint32_t m = 0u;
            if (cell_of(s, th, tw, lr, lc)) {
                const int kr = (int)jump_code(codes, lr, lc);
                m = meta_pack((uint32_t)ptr16[lr * TW + lc] & 0x3FFu, kr < 8 ? kr : 8, false, kr != 9);
            }
            const int h = hist[s];
            m_t[s] = m;
            if (x_t) x_t[s] = h;
            w_t[s] = PEND1 | ((unsigned long long)((m >> 16) & 0xFu) << W_K) | (uint32_t)h;
        }
    }
}

// Phase C.  32-bit sums (ONE = true): every thread keeps the running totals of its 16 cells in registers and the adds of
// round k go into buffer k & 1, which the owners of still-live cells fold into their totals (and clear) behind the round's
// only barrier, while round k + 1 already adds into the other buffer; a cell whose pointer has run out folds nothing any
// more, what still arrives for it stays in the two buffers and is added at the end.  64-bit sums (two such buffers do not
// fit): one array, every value read before the barrier and the adds behind it, a second barrier after the adds.
template <typename T, typename TO>
__global__ void __launch_bounds__(JTC, 1) k_jumpC(Geo g, int ntiles, const uint32_t *__restrict__ codes4,
                                                  const unsigned long long *__restrict__ wq, const uint8_t *__restrict__ tile_top,
                                                  const T *__restrict__ extra, const uint8_t *__restrict__ tile_class,
                                                  TO *__restrict__ acc, unsigned long long *__restrict__ rounds_of,
                                                  uint32_t *__restrict__ flags, bool al)
{
    // [TCELLS] sums (T; twice for 32-bit sums) + 2 x [TCELLS / 2] pointers (16 bit) + [TH * 32] codes + [NSLOT] boundary inflows (T) + 1
    extern __shared__ __align__(16) uint32_t s_dyn[];
    constexpr bool ONE = sizeof(T) == 4;
    if (flags[F_COMPLEX] == 0u) return;
    T *A = reinterpret_cast<T *>(s_dyn);
    T *B = A + TCELLS;                                  // (ONE only)
    uint32_t *pbuf = reinterpret_cast<uint32_t *>(A + (ONE ? 2 : 1) * TCELLS);
    uint32_t *codes = pbuf + TCELLS;
    T *ext = reinterpret_cast<T *>(codes + TH * 32);
    int *slot = reinterpret_cast<int *>(ext + NSLOT);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int QPT = TCELLS / 4 / JTC;               // pointer quads (four cells each) per thread
    for (;;) {
        const int tile = jump_next_tile(&flags[F_NEXT_C], slot);
        if (tile >= ntiles) return;
        if (tile_class[tile] != 4) continue;
        int ty, tx, th, tw;
        tile_dims(g, tile, ty, tx, th, tw);
        const bool full = th == TH && tw == TW;
        const uint32_t *c4 = codes4 + (int64_t)tile * TH * 32;
        for (int q = threadIdx.x; q < TH * 32; q += JTC) codes[q] = __ldcg(c4 + q);
        if (warp == 0) tile_inflow<T>(g, wq, tile_top, extra, tile, ty, tx, th, tw, lane, ext);
        __syncthreads();
        // pointers and the cells' own weight.  A cell draining into an inactive cell simply adds to it: an inactive cell has
        // no pointer, passes nothing on, and the store at the end writes zero for it.
        uint32_t *cur32 = pbuf, *nxt32 = pbuf + TCELLS / 2;
        auto no_receiver = [](int, int) { return JNONE; };
        jump_pointers(codes, th, tw, JNONE, cur32, no_receiver);
        for (int q = threadIdx.x; q < TH * 32; q += JTC) {
            const uint32_t cw = codes[q];
            T av[NC];
#pragma unroll
            for (int j = 0; j < NC; j++) av[j] = ((cw >> (4 * j)) & 15u) != 9u ? (T)1 : (T)0;
#pragma unroll
            for (int j = 0; j < NC; j++) A[NC * q + j] = av[j];
            if (ONE) {
#pragma unroll
                for (int j = 0; j < NC; j++) B[NC * q + j] = 0;
            }
        }
        __syncthreads();
        for (int s = threadIdx.x; s < NSLOT; s += JTC) {   // the boundary cells: pointers of a full tile, external inflow
            int lr, lc;
            if (!cell_of(s, th, tw, lr, lc)) continue;
            const uint32_t kr = jump_code(codes, lr, lc);
            if (full) reinterpret_cast<uint16_t *>(cur32)[lr * TW + lc] = (uint16_t)jump_pointer_boundary(kr, lr, lc, JNONE, no_receiver);
            if (kr != 9u) A[lr * TW + lc] += ext[s];
        }
        __syncthreads();
        int rounds = 0;
        if constexpr (ONE) {
            uint2 pw[QPT];
            T tot[4 * QPT];
#pragma unroll
            for (int k = 0; k < QPT; k++) {
                const int i = threadIdx.x + JTC * k;
                pw[k] = reinterpret_cast<const uint2 *>(cur32)[i];
                const uint4 v = *reinterpret_cast<const uint4 *>(A + 4 * i);
                tot[4 * k] = (T)v.x; tot[4 * k + 1] = (T)v.y; tot[4 * k + 2] = (T)v.z; tot[4 * k + 3] = (T)v.w;
                *reinterpret_cast<uint4 *>(A + 4 * i) = make_uint4(0u, 0u, 0u, 0u);
            }
            __syncthreads();                             // (buffer 0 is clear before anybody adds to it)
            for (; rounds < JROUNDS; rounds++) {
                T *buf = (rounds & 1) ? B : A;
                const uint16_t *cur16 = reinterpret_cast<const uint16_t *>(cur32);
                uint2 *nxt64 = reinterpret_cast<uint2 *>(nxt32);
                bool mine = false;
#pragma unroll
                for (int k = 0; k < QPT; k++) {
                    const int i = threadIdx.x + JTC * k;
                    const uint2 w = pw[k];
                    if ((w.x & w.y) == 0xFFFFFFFFu) { nxt64[i] = w; continue; }   // (others read "none" from either buffer)
                    const uint32_t p0 = w.x & 0xFFFFu, p1 = w.x >> 16, p2 = w.y & 0xFFFFu, p3 = w.y >> 16;
                    if (p0 != JNONE) smem_add<T>(&buf[p0], tot[4 * k]);
                    if (p1 != JNONE) smem_add<T>(&buf[p1], tot[4 * k + 1]);
                    if (p2 != JNONE) smem_add<T>(&buf[p2], tot[4 * k + 2]);
                    if (p3 != JNONE) smem_add<T>(&buf[p3], tot[4 * k + 3]);
                    // (no branches: a cell without pointer reads some cell's pointer and keeps "none")
                    const uint32_t n0 = cur16[p0 & 0x3FFFu], n1 = cur16[p1 & 0x3FFFu], n2 = cur16[p2 & 0x3FFFu], n3 = cur16[p3 & 0x3FFFu];
                    pw[k] = make_uint2((p0 == JNONE ? JNONE : n0) | ((p1 == JNONE ? JNONE : n1) << 16),
                                       (p2 == JNONE ? JNONE : n2) | ((p3 == JNONE ? JNONE : n3) << 16));
                    nxt64[i] = pw[k];
                    mine = true;
                }
                if (!__syncthreads_or(mine)) break;
#pragma unroll
                for (int k = 0; k < QPT; k++) {         // still-live cells fold what has just arrived for them
                    const uint2 w = pw[k];
                    if ((w.x & w.y) == 0xFFFFFFFFu) continue;
                    const int i = threadIdx.x + JTC * k;
                    uint4 v = *reinterpret_cast<const uint4 *>(buf + 4 * i);
                    if ((w.x & 0xFFFFu) != JNONE) { tot[4 * k] += (T)v.x; v.x = 0u; }
                    if ((w.x >> 16) != JNONE) { tot[4 * k + 1] += (T)v.y; v.y = 0u; }
                    if ((w.y & 0xFFFFu) != JNONE) { tot[4 * k + 2] += (T)v.z; v.z = 0u; }
                    if ((w.y >> 16) != JNONE) { tot[4 * k + 3] += (T)v.w; v.w = 0u; }
                    *reinterpret_cast<uint4 *>(buf + 4 * i) = v;
                }
                uint32_t *t = cur32; cur32 = nxt32; nxt32 = t;
            }
            // the sums: own total + what is left in the two buffers.  (Behind the last barrier nobody adds any more; the
            // folds above touch own cells only.)
#pragma unroll
            for (int k = 0; k < QPT; k++) {
                const int i = threadIdx.x + JTC * k;
                const uint4 va = *reinterpret_cast<const uint4 *>(A + 4 * i), vb = *reinterpret_cast<const uint4 *>(B + 4 * i);
                *reinterpret_cast<uint4 *>(A + 4 * i) = make_uint4((uint32_t)tot[4 * k] + va.x + vb.x, (uint32_t)tot[4 * k + 1] + va.y + vb.y,
                                                                  (uint32_t)tot[4 * k + 2] + va.z + vb.z, (uint32_t)tot[4 * k + 3] + va.w + vb.w);
            }
            __syncthreads();
        } else {
            for (; rounds < JROUNDS; rounds++) {
                const uint16_t *cur16 = reinterpret_cast<const uint16_t *>(cur32);
                const uint2 *cur64 = reinterpret_cast<const uint2 *>(cur32);
                uint2 *nxt64 = reinterpret_cast<uint2 *>(nxt32);
                uint2 pw[QPT];
                T a[4 * QPT];
                bool mine = false;
#pragma unroll
                for (int k = 0; k < QPT; k++) {
                    const int i = threadIdx.x + JTC * k;
                    const uint2 w = cur64[i];
                    pw[k] = w;
                    if ((w.x & w.y) == 0xFFFFFFFFu) { nxt64[i] = w; continue; }
                    const uint32_t p0 = w.x & 0xFFFFu, p1 = w.x >> 16, p2 = w.y & 0xFFFFu, p3 = w.y >> 16;
#pragma unroll
                    for (int j = 0; j < 4; j++) a[4 * k + j] = A[4 * i + j];
                    const uint32_t n0 = cur16[p0 & 0x3FFFu], n1 = cur16[p1 & 0x3FFFu], n2 = cur16[p2 & 0x3FFFu], n3 = cur16[p3 & 0x3FFFu];
                    nxt64[i] = make_uint2((p0 == JNONE ? JNONE : n0) | ((p1 == JNONE ? JNONE : n1) << 16),
                                          (p2 == JNONE ? JNONE : n2) | ((p3 == JNONE ? JNONE : n3) << 16));
                    mine = true;
                }
                if (!__syncthreads_or(mine)) break;     // (the barrier also puts every read of A before the adds)
#pragma unroll
                for (int k = 0; k < QPT; k++) {
                    if ((pw[k].x & pw[k].y) == 0xFFFFFFFFu) continue;
                    const uint32_t p0 = pw[k].x & 0xFFFFu, p1 = pw[k].x >> 16, p2 = pw[k].y & 0xFFFFu, p3 = pw[k].y >> 16;
                    if (p0 != JNONE) smem_add<T>(&A[p0], a[4 * k]);
                    if (p1 != JNONE) smem_add<T>(&A[p1], a[4 * k + 1]);
                    if (p2 != JNONE) smem_add<T>(&A[p2], a[4 * k + 2]);
                    if (p3 != JNONE) smem_add<T>(&A[p3], a[4 * k + 3]);
                }
                __syncthreads();
                uint32_t *t = cur32; cur32 = nxt32; nxt32 = t;
            }
        }
        if (threadIdx.x == 0) rounds_of[tile] |= (unsigned long long)rounds << 32;
        const int64_t gr0 = (int64_t)ty * TH, gc0 = (int64_t)tx * TW;
        for (int q = threadIdx.x; q < TH * 32; q += JTC) {
            const int lr = q >> 5, l = q & 31;
            if (lr >= th || gc0 + NC * l >= g.nx) continue;
            const uint32_t cw = codes[q];
            T tot[NC];
#pragma unroll
            for (int j = 0; j < NC; j++) tot[j] = ((cw >> (4 * j)) & 15u) != 9u ? A[NC * q + j] : (T)0;
            TO *out = acc + (gr0 + lr) * g.nx + gc0 + NC * l;
            if (al) store_row<true, T, TO>(out, tot, gc0 + NC * l, g.nx);
            else store_row<false, T, TO>(out, tot, gc0 + NC * l, g.nx);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// basin membership (wmf_py/cu_py/basics.py:206-248, cuencas_heavy.f90:507-576): which cells drain to the outlet?
// ---------------------------------------------------------------------------------------------
// The same contraction as the accumulation, with a boolean instead of a sum.  (Pointer jumping over the WHOLE raster,
// which this replaces, costs log2(cells) rounds of random 4-byte gathers over every cell: 16.8 ms for a 8192^2 raster.)
//   k_membA   per tile, pointer jumping in shared memory exactly like k_jumpA: every cell gets the label of the exit its
//             path leaves the tile through -- or LOUT when the path reaches the outlet inside the tile (the outlet is a
//             sink: it keeps no receiver), or LNONE.  Labels go to global memory (2 bytes per cell), the boundary cells'
//             labels and codes into the meta words.  Cycles never resolve and end as LNONE: not in the basin.
//   k_memb_build  per boundary slot: the exit the flow leaving here leaves the NEXT tile through (state >= 0), or the
//             verdict when the path ends in that tile: MEMB_YES at the outlet, MEMB_NO anywhere else.
//   k_memb_jump   state[x] = state[state[x]] over the exits (3.9 % of the cells) until every state is a verdict;
//             a round that changes nothing switches the remaining launches off (no host round trip).
//   k_membC   per cell: label -> LOUT, or the verdict of that exit.  Writes the mask, counts the cells.
constexpr uint32_t LOUT = 1022u;
constexpr int32_t MEMB_NO = -1, MEMB_YES = -2;
constexpr int MEMB_ROUNDS = 26;
static_assert(LOUT >= (uint32_t)NSLOT && LOUT != LNONE, "the outlet label is no slot");

template <int ENC>
__global__ void __launch_bounds__(JTA, 4) k_membA(const int8_t *__restrict__ dir, const uint8_t *__restrict__ mask, Geo g, int ntiles,
                                               int64_t outlet, uint32_t *__restrict__ meta, uint16_t *__restrict__ labels, bool vec)
{
    extern __shared__ __align__(16) uint32_t s_dyn[];   // [TCELLS / 2] pointers (16 bit) + [TH * 32] codes
    uint32_t *ptr32 = s_dyn;
    uint16_t *ptr16 = reinterpret_cast<uint16_t *>(s_dyn);
    uint32_t *codes = s_dyn + TCELLS / 2;
    const int64_t orow = outlet / g.nx, ocol = outlet - orow * g.nx;
    const int otile = (int)(orow / TH) * g.tiles_x + (int)(ocol / TW);
    const int ol = (int)(orow % TH) * TW + (int)(ocol % TW);             // the outlet's cell index inside its tile
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        __syncthreads();                               // (the previous tile is done with shared memory)
        int ty, tx, th, tw;
        tile_dims(g, tile, ty, tx, th, tw);
        jump_decode<ENC>(dir, mask, g, ty, tx, th, vec, codes, nullptr);
        __syncthreads();
        if (tile == otile && threadIdx.x == 0) {       // the outlet keeps no receiver
            const int q = (ol / TW) * 32 + (ol % TW) / NC, sh = 4 * ((ol % TW) % NC);
            if (((codes[q] >> sh) & 15u) != 9u) codes[q] = (codes[q] & ~(15u << sh)) | (8u << sh);
        }
        __syncthreads();
        auto exit_label = [&](int lr, int lc) { return JRES | (uint32_t)slot_of(lr, lc, th, tw); };
        jump_pointers(codes, th, tw, JRES | LNONE, ptr32, exit_label);
        const bool full = th == TH && tw == TW;
        if (full) {
            __syncthreads();
            for (int s = threadIdx.x; s < NSLOT; s += JTA) {
                int lr, lc;
                cell_of(s, TH, TW, lr, lc);
                ptr16[lr * TW + lc] = (uint16_t)jump_pointer_boundary(jump_code(codes, lr, lc), lr, lc, JRES | LNONE, exit_label);
            }
        }
        __syncthreads();
        if (tile == otile && threadIdx.x == 0 && jump_code(codes, ol / TW, ol % TW) != 9u) ptr16[ol] = (uint16_t)(JRES | LOUT);
        __syncthreads();
        bool open = true;
        for (int rounds = 0; open && rounds < JROUNDS; rounds++) {
            uint32_t all = 0x80008000u;
            uint2 *ptr64 = reinterpret_cast<uint2 *>(ptr32);
            for (int i = threadIdx.x; i < TCELLS / 4; i += JTA) {
                uint2 w = ptr64[i];
                if ((w.x & w.y & 0x80008000u) == 0x80008000u) continue;
                const uint32_t a0 = w.x & 0xFFFFu, a1 = w.x >> 16, a2 = w.y & 0xFFFFu, a3 = w.y >> 16;
                const uint32_t b0 = ptr16[a0 & 0x3FFFu], b1 = ptr16[a1 & 0x3FFFu], b2 = ptr16[a2 & 0x3FFFu], b3 = ptr16[a3 & 0x3FFFu];
                w.x = ((a0 & JRES) ? a0 : b0) | (((a1 & JRES) ? a1 : b1) << 16);
                w.y = ((a2 & JRES) ? a2 : b2) | (((a3 & JRES) ? a3 : b3) << 16);
                ptr64[i] = w;
                all &= w.x & w.y;
            }
            open = __syncthreads_or(all != 0x80008000u) != 0;
        }
        // labels of all cells (an unresolved pointer sits on a cycle: no exit), eight cells per store
        uint16_t *l_t = labels + (int64_t)tile * TCELLS;
        for (int q = threadIdx.x; q < TCELLS / 8; q += JTA) {
            const uint4 x = *reinterpret_cast<const uint4 *>(ptr32 + 4 * q);
            auto fix = [](uint32_t w) {
                const uint32_t lo = (w & JRES) ? (w & 0x3FFu) : LNONE, hi = (w & (JRES << 16)) ? ((w >> 16) & 0x3FFu) : LNONE;
                return lo | (hi << 16);
            };
            *reinterpret_cast<uint4 *>(l_t + 8 * q) = make_uint4(fix(x.x), fix(x.y), fix(x.z), fix(x.w));
        }
        uint32_t *m_t = meta + (int64_t)tile * NSLOT;
        for (int s = threadIdx.x; s < NSLOT; s += JTA) {
            int lr, lc;
            uint32_t m = 0u;
            if (cell_of(s, th, tw, lr, lc)) {
                const int kr = (int)jump_code(codes, lr, lc);
                const uint32_t pw = ptr16[lr * TW + lc];
                m = meta_pack((pw & JRES) ? (pw & 0x3FFu) : LNONE, kr < 8 ? kr : 8, false, kr != 9);
            }
            m_t[s] = m;
        }
    }
}

__global__ void __launch_bounds__(256) k_memb_build(Geo g, int64_t nnodes, const uint32_t *__restrict__ meta, int32_t *__restrict__ state)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nnodes) return;
    const int tile = (int)(i / NSLOT), s = (int)(i - (int64_t)tile * NSLOT);
    const int ty = tile / g.tiles_x, tx = tile - ty * g.tiles_x;
    const int last_th = (int)(g.ny - (int64_t)(g.tiles_y - 1) * TH), last_tw = (int)(g.nx - (int64_t)(g.tiles_x - 1) * TW);
    int s2 = 0;
    const int t2 = build_target(g, ty, tx, last_th, last_tw, s, meta[i], s2);
    int32_t st = MEMB_NO;
    if (t2 >= 0) {
        const uint32_t me = meta[(int64_t)t2 * NSLOT + s2];
        const uint32_t link = me & 0xFFFFu;
        if ((me >> 21) & 1) st = link == LOUT ? MEMB_YES : (link == LNONE ? MEMB_NO : t2 * NSLOT + (int32_t)link);
    }
    state[i] = st;
}

// flags[r] != 0: round r moved something (round 0 always runs)
__global__ void __launch_bounds__(256) k_memb_jump(int64_t nnodes, int32_t *state, int *flags, int round)
{
    if (round > 0 && flags[round - 1] == 0) return;                      // (persistent grid: a switched-off round costs a launch)
    bool moved = false;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nnodes; i += (int64_t)gridDim.x * blockDim.x) {
        const int32_t p = state[i];
        if (p < 0) continue;
        state[i] = __ldcg(&state[p]);                                    // a farther exit, or its verdict
        moved = true;
    }
    if (moved) flags[round] = 1;
}

__global__ void __launch_bounds__(256) k_membC(Geo g, int ntiles, const uint16_t *__restrict__ labels, const int32_t *__restrict__ state,
                                               uint8_t *__restrict__ bm, int32_t *__restrict__ flag32, unsigned long long *count)
{
    __shared__ uint8_t yes[NSLOT];
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        __syncthreads();
        for (int s = threadIdx.x; s < NSLOT; s += blockDim.x) yes[s] = state[(int64_t)tile * NSLOT + s] == MEMB_YES;
        __syncthreads();
        int ty, tx, th, tw;
        tile_dims(g, tile, ty, tx, th, tw);
        const uint16_t *l_t = labels + (int64_t)tile * TCELLS;
        unsigned int n = 0;
        for (int idx = threadIdx.x; idx < TCELLS; idx += blockDim.x) {
            const int lr = idx / TW, lc = idx - lr * TW;
            if (lr >= th || lc >= tw) continue;
            const uint32_t lab = l_t[idx];
            const bool in = lab == LOUT || (lab < (uint32_t)NSLOT && yes[lab]);
            const int64_t gi = ((int64_t)ty * TH + lr) * g.nx + (int64_t)tx * TW + lc;
            if (bm) bm[gi] = in ? 1 : 0;
            if (flag32) flag32[gi] = in ? 1 : 0;
            n += in;
        }
        n = __reduce_add_sync(FULL, n);
        if ((threadIdx.x & 31) == 0 && n) atomicAdd(count, (unsigned long long)n);
    }
}

// ---------------------------------------------------------------------------------------------
// phase G: contracted forest over boundary slots
// ---------------------------------------------------------------------------------------------
// entry slot (node id) that exit slot (tile, s) drains to; -1 when s is not an exit; -2 when the
// receiver lies in a neighbouring row stripe (a stripe exit: it has no parent in this stripe)
__device__ __forceinline__ int64_t entry_of(const Geo &g, const uint32_t *__restrict__ meta, int tile, int s,
                                            uint32_t &m_entry)
{
    int ty, tx, th, tw, lr, lc;
    tile_dims(g, tile, ty, tx, th, tw);
    if (!cell_of(s, th, tw, lr, lc)) return -1;
    uint32_t m = meta[(int64_t)tile * NSLOT + s];
    if (!((m >> 21) & 1)) return -1;
    int k = (m >> 16) & 0xF;
    if (k == 8) return -1;
    int rr = lr + d8_dr(k), cc = lc + d8_dc(k);
    if (rr >= 0 && rr < th && cc >= 0 && cc < tw) return -1;            // stays in the tile: not an exit
    int64_t gr = (int64_t)ty * TH + rr, gc = (int64_t)tx * TW + cc;
    if (gc < 0 || gc >= g.nx) return -1;                                // off the raster's side: no receiver
    if (gr < 0 || gr >= g.ny) return -2;
    int ty2 = (int)(gr / TH), tx2 = (int)(gc / TW);
    int tile2 = ty2 * g.tiles_x + tx2, ty3, tx3, th2, tw2;
    tile_dims(g, tile2, ty3, tx3, th2, tw2);
    int s2 = slot_of((int)(gr - (int64_t)ty2 * TH), (int)(gc - (int64_t)tx2 * TW), th2, tw2);
    int64_t node = (int64_t)tile2 * NSLOT + s2;
    m_entry = meta[node];
    return node;
}

__global__ void k_build_forest(Geo g, int64_t nnodes, const uint32_t *__restrict__ meta, int32_t *__restrict__ parent,
                               uint8_t *__restrict__ blocked)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nnodes) return;
    int tile = (int)(i / NSLOT), s = (int)(i - (int64_t)tile * NSLOT);
    uint32_t me = 0;
    int64_t e = entry_of(g, meta, tile, s, me);
    int32_t p = -1;
    if (e >= 0 && ((me >> 21) & 1)) {
        uint32_t link = me & 0xFFFFu;
        if (link != LNONE) p = (int32_t)((e / NSLOT) * NSLOT + link);
    }
    parent[i] = p;
    blocked[i] = (uint8_t)((meta[i] >> 20) & 1);
}

// One thread per boundary slot, BU consecutive tiles per block: the kernel is two dependent gathers (the slot's meta
// word, then the meta word of the entry it drains to) plus two reductions, so each thread keeps BU of them in flight.
// par: the parent exit of every slot (>= 0), or PAR_ROOT (an exit whose flow ends in the next tile), PAR_STRIPE (an exit
// into a neighbouring row stripe), PAR_NONE (not an exit, or its receiver is inactive: never walked).
// pe (nullable; the stripe kernels chase it): one 8-byte record {parent exit, entry cell} per node.
constexpr int BU = 4;
// Band schedule (whole rasters, 32-bit sums; see sweep_phaseAG_bands): the forest is built and walked band by band of
// tile rows, behind phase A.  gstate[tile]: 0 nothing yet, 1 built, 2 walk started.
//   mode 0  every tile of the launch (stripes, small rasters)
//   mode 1  band: a tile is BUILT when the fast solver has finished its whole 3 x 3 neighbourhood (tile_class 0 -- tiles
//           of the jump / chain-walk solvers get their meta words only after all bands) and STARTED when its whole
//           neighbourhood is built; as long as no tile anywhere went to those solvers (flags) that is every tile
//   mode 2  rest, after the jump / chain-walk solvers: whatever mode 1 left out; nothing when no tile went there
struct BandCtl { const uint8_t *tile_class; uint8_t *gstate; const uint32_t *flags; int mode; };
__device__ __forceinline__ bool band_clean(const BandCtl &bc) { return bc.flags[F_COMPLEX] == 0u && bc.flags[F_CHAIN] == 0u; }
// does every tile around `tile` (itself included) satisfy v[t] <relation>?  built: v = gstate >= 1; done: v = tile_class == 0
__device__ __forceinline__ bool band_around(const Geo &g, int tile, const uint8_t *v, bool want_zero)
{
    const int ty = tile / g.tiles_x, tx = tile - ty * g.tiles_x;
    bool ok = true;
    for (int y = ty - 1; y <= ty + 1; y++)
        for (int x = tx - 1; x <= tx + 1; x++)
            if (y >= 0 && y < g.tiles_y && x >= 0 && x < g.tiles_x) {
                const uint8_t b = v[y * g.tiles_x + x];
                ok = ok && (want_zero ? b == 0 : b != 0);
            }
    return ok;
}

template <int MODE>
__global__ void __launch_bounds__(NSLOT) k_g_build(Geo g, int ntiles, const uint32_t *__restrict__ meta,
                                                   const uint8_t *__restrict__ tile_top, int32_t *__restrict__ par,
                                                   int2 *__restrict__ pe, unsigned long long *wq,
                                                   BandCtl bc = BandCtl{nullptr, nullptr, nullptr, 0})
{
  if (MODE == 2 && band_clean(bc)) return;
  const int nblk = (ntiles - g.tile0 + BU - 1) / BU;
  for (int blk = blockIdx.x; MODE != 2 || blk < nblk; blk += gridDim.x) {   // (mode 2 runs a persistent grid; the others one block per BU tiles)
    // last tiles first: what this kernel leaves in L2 is then weighted towards the FIRST tile rows, where the walk
    // that follows starts its chains
    const int tile0 = g.tile0 + (nblk - 1 - blk) * BU, s = threadIdx.x;
    const int last_th = (int)(g.ny - (int64_t)(g.tiles_y - 1) * TH), last_tw = (int)(g.nx - (int64_t)(g.tiles_x - 1) * TW);
    uint32_t m[BU], me[BU];
    int t2[BU], s2[BU];
    bool skip[BU];                                           // an inner first-row slot of a tile that has no exit there (tile_top)
    bool mine[BU];                                           // band schedule: this launch builds the tile
#pragma unroll
    for (int u = 0; u < BU; u++)                             // (in flight before the band tests below)
        m[u] = (tile0 + u < ntiles && !(s > 0 && s < TW - 1 && !tile_top[tile0 + u])) ? __ldcs(&meta[(int64_t)(tile0 + u) * NSLOT + s]) : 0u;
#pragma unroll
    for (int u = 0; u < BU; u++) {
        mine[u] = tile0 + u < ntiles;
        if (MODE == 1 && mine[u]) mine[u] = band_clean(bc) || band_around(g, tile0 + u, bc.tile_class, true);
        if (MODE == 2 && mine[u]) mine[u] = bc.gstate[tile0 + u] == 0;
    }
    if (MODE != 0) {
        if (MODE == 2) __syncthreads();                      // (mode 2: everybody has read gstate before it changes)
        if (s == 0) {
#pragma unroll
            for (int u = 0; u < BU; u++) if (mine[u]) bc.gstate[tile0 + u] = 1;
        }
    }
#pragma unroll
    for (int u = 0; u < BU; u++) skip[u] = !mine[u] || (s > 0 && s < TW - 1 && !tile_top[tile0 + u]);
#pragma unroll
    for (int u = 0; u < BU; u++) if (skip[u]) m[u] = 0u;
    int ty = tile0 / g.tiles_x, tx = tile0 - ty * g.tiles_x;            // block-uniform; the next tiles follow in row-major order
#pragma unroll
    for (int u = 0; u < BU; u++) {
        s2[u] = 0;
        t2[u] = !skip[u] ? build_target(g, ty, tx, last_th, last_tw, s, m[u], s2[u]) : -1;
        if (++tx == g.tiles_x) { tx = 0; ty++; }
    }
#pragma unroll
    for (int u = 0; u < BU; u++) me[u] = t2[u] >= 0 ? __ldcs(&meta[t2[u] * NSLOT + s2[u]]) : 0u;
#pragma unroll
    for (int u = 0; u < BU; u++) {
        if (skip[u]) continue;                               // (its slot is never walked: see node_is_skipped)
        int32_t p = -1, en = t2[u];
        if (en >= 0) {
            en = -1;
            if ((me[u] >> 21) & 1) {                                    // the receiver is active
                en = t2[u] * NSLOT + s2[u];
                const uint32_t link = me[u] & 0xFFFFu;
                if (link != LNONE) {
                    p = t2[u] * NSLOT + (int)link;
                    atomicAdd(&wq[p], PEND1);
                }
            }
        }
        par[(int64_t)(tile0 + u) * NSLOT + s] = p >= 0 ? p : (en >= 0 ? PAR_ROOT : (en == -2 ? PAR_STRIPE : PAR_NONE));
        if (pe) pe[(int64_t)(tile0 + u) * NSLOT + s] = make_int2(p, en);
    }
    if (MODE != 2) break;
  }
}

// "Did every exit finalise?" without a host round trip: GCNT words (one 32-byte sector each, so the
// per-warp adds spread out) hold exits << 32 | finalised; k_g_unres has work only when the halves differ.
constexpr int GCNT = 64;                                     // power of two
constexpr size_t GCNT_BYTES = sizeof(unsigned long long) * 4 * GCNT;
// inner first-row slots of a full tile whose row 0 has no upward flow are never exits: k_g_build left their words alone
__device__ __forceinline__ bool node_is_skipped(const uint8_t *__restrict__ tile_top, int64_t i)
{
    const uint32_t iu = (uint32_t)i, tile = iu / (uint32_t)NSLOT, s = iu - tile * (uint32_t)NSLOT;   // (node ids are below 2^31)
    return s > 0u && s < (uint32_t)(TW - 1) && !tile_top[tile];
}


template <int MODE>
__global__ void __launch_bounds__(256) k_g_walk(int64_t nnodes, const uint8_t *__restrict__ tile_top, const int32_t *__restrict__ par,
                                                unsigned long long *wq, unsigned long long *__restrict__ gcnt, int64_t node0 = 0,
                                                Geo g = Geo{}, BandCtl bc = BandCtl{nullptr, nullptr, nullptr, 0})
{
  if (MODE == 2 && band_clean(bc)) return;
  unsigned int nexit_all = 0, nfin = 0;
  // one node per thread; mode 2 runs a persistent grid over all nodes (whole warps stay together for the ballot)
  for (int64_t i = node0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; MODE != 2 || i - (threadIdx.x & 31) < nnodes;
       i += (int64_t)gridDim.x * blockDim.x) {
    bool mine = i < nnodes;
    int32_t p = (mine && !node_is_skipped(tile_top, i)) ? __ldcs(&par[i]) : PAR_NONE;     // (in flight before the band tests below)
    if (MODE != 0 && mine) {
        const int tile = (int)((uint32_t)i / (uint32_t)NSLOT);
        if (MODE == 1) mine = band_clean(bc) || band_around(g, tile, bc.gstate, false);
        else mine = bc.gstate[tile] != 2;
        // The tile's mark is set by the thread of its slot 0: 2 in a band launch -- others in the same launch only ask
        // "built?", and 1 and 2 both say yes -- and 3 in the rest launch, whose test "!= 2" it does not change.
        if (mine && (uint32_t)i == (uint32_t)tile * (uint32_t)NSLOT) bc.gstate[tile] = MODE == 1 ? 2 : 3;
    }
    if (!mine) p = PAR_NONE;
    const bool walk = p != PAR_NONE;
    const unsigned int nexit = __popc(__ballot_sync(FULL, walk));
    nexit_all += nexit;
    if (walk) {                                             // an exit: carry its total downstream
        unsigned long long old = atomicAdd(&wq[i], (unsigned long long)(-(long long)PEND1));   // own arrival
        uint32_t add = 0;
        while (w_pending(old) == 1u) {                      // we took pending to zero: the node is final, its word holds the total
            const uint32_t total = (uint32_t)old + add;
            nfin++;
            if (p < 0) break;                               // the chain ends here: a root, or a stripe exit
            // the parent's parent does not depend on the atomic: fetch it first so the hop costs ONE L2 round trip
            const int32_t nxt = par[p];
            old = atomicAdd(&wq[p], (unsigned long long)total - PEND1);
            add = total;
            p = nxt;
        }
    }
    __syncwarp();
    if (MODE != 2) break;
  }
    nfin = __reduce_add_sync(FULL, nfin);
    if ((threadIdx.x & 31) == 0 && nexit_all)
        atomicAdd(&gcnt[4 * (blockIdx.x & (GCNT - 1))], ((unsigned long long)nexit_all << 32) | nfin);
}

// Exits that never finalised (cycles, or fed by one) poison the tile they drain into: it goes to the general solver,
// which sees the unresolved donors when it pulls.  Nothing to do when every exit finalised, which the two counters
// tell without a host round trip.
__global__ void k_g_unres(Geo g, int64_t nnodes, const uint8_t *__restrict__ tile_top, const int32_t *__restrict__ par,
                          const unsigned long long *__restrict__ wq, uint8_t *__restrict__ tile_class,
                          const unsigned long long *__restrict__ gcnt, uint32_t *__restrict__ flags)
{
    {
        const int lane = threadIdx.x & 31;
        unsigned long long c = 0;
        for (int q = lane; q < GCNT; q += 32) c += gcnt[4 * q];
        const unsigned int exits = __reduce_add_sync(FULL, (unsigned int)(c >> 32)), fin = __reduce_add_sync(FULL, (unsigned int)c);
        if (exits == fin) return;
    }
    const TileGeo tg = tile_geo(g);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nnodes; i += (int64_t)gridDim.x * blockDim.x) {
        if (node_is_skipped(tile_top, i)) continue;
        const int32_t pp = par[i];
        if (pp == PAR_NONE || pp == PAR_STRIPE) continue;
        const unsigned long long w = wq[i];
        if (w_pending(w) == 0u) continue;
        const int tile = (int)(i / NSLOT), ty = tile / g.tiles_x, tx = tile - ty * g.tiles_x;
        int s2;
        const int te = exit_target(g, tg, ty, tx, (int)(i - (int64_t)tile * NSLOT), w_k(w), true, s2);
        if (te >= 0) { tile_class[te] = 2; raise_flag(flags, F_UNRES); }   // (racing writers all store the same values)
    }
}

template <typename T>
__device__ __forceinline__ void inflow_add(T *p, T v);
template <>
__device__ __forceinline__ void inflow_add<int32_t>(int32_t *p, int32_t v) { atomicAdd(p, v); }
template <>
__device__ __forceinline__ void inflow_add<int64_t>(int64_t *p, int64_t v)
{
    atomicAdd((unsigned long long *)p, (unsigned long long)v);
}

template <typename T>
__global__ void k_push_inflow(Geo g, int64_t nnodes, const uint32_t *__restrict__ meta, const T *__restrict__ A,
                              const int32_t *__restrict__ cnt, T *__restrict__ inflow, uint8_t *__restrict__ ext_unres,
                              uint8_t *__restrict__ tile_class, uint32_t *__restrict__ flags)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nnodes) return;
    int tile = (int)(i / NSLOT), s = (int)(i - (int64_t)tile * NSLOT);
    uint32_t me = 0;
    int64_t e = entry_of(g, meta, tile, s, me);
    if (e < 0 || !((me >> 21) & 1)) return;                           // not an exit, or the receiver is inactive
    if (cnt[i] <= 0) inflow_add<T>(&inflow[e], A[i]);                 // finalised exit: push its total
    else {                                                            // unresolved upstream: the entry never finalises
        ext_unres[e] = 1;
        tile_class[e / NSLOT] = 2;                                    // (racing writers all store non-zero)
        flags[F_UNRES] = 1u;
    }
}

// ---------------------------------------------------------------------------------------------
// row stripes: the stripe is solved like a raster whose top / bottom
// neighbours exist but are elsewhere.  Cells of its first / last row that drain across are "stripe
// exits"; k_stripe_summary condenses the stripe to 2*nx boundary records, k_stripe_inject feeds the
// inflow arriving from the other stripes back in before the second forest walk.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int64_t boundary_node(const Geo &g, int side, int64_t col, int &tile)
{
    const int64_t row = side == 0 ? 0 : g.ny - 1;
    const int ty = (int)(row / TH), tx = (int)(col / TW);
    tile = ty * g.tiles_x + tx;
    int ty2, tx2, th, tw;
    tile_dims(g, tile, ty2, tx2, th, tw);
    return (int64_t)tile * NSLOT + slot_of((int)(row - (int64_t)ty * TH), (int)(col - (int64_t)tx * TW), th, tw);
}

// Follow the contracted path of flow that enters at boundary-slot node e0: exit x1 = link(e0), then
// x_{i+1} = parent[x_i].  Returns the last exit of the path (or -1 when link(e0) is none).
__device__ __forceinline__ int64_t first_exit(const uint32_t *__restrict__ meta, int64_t e0)
{
    const uint32_t l = meta[e0] & 0xFFFFu;
    return l == LNONE ? -1 : (e0 / NSLOT) * NSLOT + l;
}

// Contracted path of every stripe-boundary cell (2*nx of them): the entries flow passes after entering
// the stripe at that cell.  Only the forest (parent / entry) is needed, so the chase runs beside
// k_g_walk on a second stream.  It is a chain of dependent loads, one per tile crossed, so it is cut
// short with macro links: tile rows are grouped MR at a time, k_macro_links chases (in parallel, <= MR-ish
// hops each) from every top-row entry of a group to the exit that leaves the group, and the boundary
// chase then takes one jump per group.  A jump is recorded as one path item (JUMP flag) that the route
// kernel expands again.  path is hop-major ([hop][2*nx]): chase and route both touch it coalesced; a
// path longer than pcap items keeps its continuation (pce, pcx) and the route kernel finishes it serially.
constexpr int MR = 16;
constexpr int32_t JUMP = (int32_t)0x80000000;
struct StripePaths { int32_t *path, *plen, *pce, *pcx, *plink, *mlink; int pcap; };

// macro row of an entry node when it is a top-row cell of the first tile row of a group, else -1
__device__ __forceinline__ int macro_top(const Geo &g, int64_t e)
{
    const int tile = (int)(e / NSLOT), s = (int)(e - (int64_t)tile * NSLOT);
    const int ty = tile / g.tiles_x;
    return (s < TW && ty % MR == 0) ? ty / MR : -1;
}
__device__ __forceinline__ int macro_of(const Geo &g, int64_t e) { return (int)(e / NSLOT) / g.tiles_x / MR; }

// A chase along parent exits must not spin on a drainage cycle of the contracted forest (cycles are legal input): Brent's
// cycle detection, one comparison per hop -- the chase stops at the first exit it meets again.
struct ChaseGuard {
    int64_t mark;
    uint32_t steps, limit;
    __device__ ChaseGuard() : mark(-1), steps(0), limit(2) {}
    __device__ __forceinline__ bool seen(int64_t x)
    {
        if (x == mark) return true;
        if (++steps == limit) { mark = x; limit *= 2; steps = 0; }
        return false;
    }
};

// mlink[macro row][col]: the exit (node id) through which flow entering at that top-row cell leaves the
// group of tile rows (its entry lies in another group, in another stripe, or nowhere); -1: ends inside.
__global__ void k_macro_links(Geo g, int64_t nnodes, const uint32_t *__restrict__ meta, const int2 *__restrict__ pe,
                              int32_t *__restrict__ mlink)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int nmac = (g.tiles_y + MR - 1) / MR;
    if (i >= (int64_t)nmac * g.nx) return;
    const int mr = (int)(i / g.nx);
    const int64_t col = i - (int64_t)mr * g.nx;
    const int tile = mr * MR * g.tiles_x + (int)(col / TW);
    const int64_t e0 = (int64_t)tile * NSLOT + (col % TW);
    int32_t out = -1;
    if ((meta[e0] >> 21) & 1) {
        int64_t x = first_exit(meta, e0);
        ChaseGuard cg;
        for (int64_t hops = 0; x >= 0 && hops <= nnodes; hops++) {
            if (cg.seen(x)) break;                           // a cycle inside the group: the path never leaves it
            const int2 r = pe[x];
            if (r.y < 0 || macro_of(g, r.y) != mr) { out = (int32_t)x; break; }
            x = r.x;
        }
    }
    mlink[i] = out;
}

__global__ void k_stripe_paths(Geo g, int64_t nnodes, const uint32_t *__restrict__ meta, const int2 *__restrict__ pe,
                               StripePaths sp)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, n2 = 2 * g.nx;
    if (i >= n2) return;
    const int side = (int)(i / g.nx);
    const int64_t col = i - (int64_t)side * g.nx;
    int tile;
    const int64_t node = boundary_node(g, side, col, tile);
    int32_t len = 0, link = -1, ce = -1, cx = -1;
    if ((meta[node] >> 21) & 1) {
        int64_t e = node, x = -2, last = -1;                 // x == -2: not looked up yet
        ChaseGuard cg;
        for (int64_t hops = 0; hops <= nnodes; hops++) {
            if (cg.seen(e)) break;                           // a cycle: the path ends inside the stripe
            const int mr = macro_top(g, e);
            if (len >= sp.pcap) {                            // out of room: the serial tail takes over at e (no jumps there)
                if (ce < 0) { ce = (int32_t)e; cx = (int32_t)(x == -2 ? first_exit(meta, e) : x); }
                if (x == -2) x = first_exit(meta, e);
            } else if (mr >= 0) {                            // one jump across the group of tile rows
                const int64_t ecol = (int64_t)((int)(e / NSLOT) % g.tiles_x) * TW + (e % NSLOT);
                sp.path[(int64_t)len * n2 + i] = (int32_t)e | JUMP;
                len++;
                x = sp.mlink[(int64_t)mr * g.nx + ecol];
                if (x < 0) break;                            // the path ends inside the group
                const int32_t en = pe[x].y;
                if (en < 0) { last = x; break; }
                e = en; x = -2;
                continue;
            } else {
                sp.path[(int64_t)len * n2 + i] = (int32_t)e;
                len++;
                if (x == -2) x = first_exit(meta, e);
            }
            if (x < 0) break;                                // the path ends inside the stripe
            const int2 r = pe[x];
            if (r.y < 0) { last = x; break; }                // x leaves the stripe (-2) or has no active receiver (-1)
            e = r.y;
            x = r.x;                                         // = first_exit(meta, en)
        }
        if (last >= 0 && pe[last].y == -2) {                 // the path leaves the stripe at exit `last`
            const int xt = (int)(last / NSLOT);
            int ty, tx, th, tw, lr, lc;
            tile_dims(g, xt, ty, tx, th, tw);
            cell_of((int)(last - (int64_t)xt * NSLOT), th, tw, lr, lc);
            const int64_t row = (int64_t)ty * TH + lr, c2 = (int64_t)tx * TW + lc;
            link = (int32_t)((row == 0 ? 0 : 1) * g.nx + c2);
        }
    }
    sp.plen[i] = len; sp.pce[i] = ce; sp.pcx[i] = cx; sp.plink[i] = link;
}

// Packed boundary records, int64 [2][2][nx] (index side*nx + col inside a plane):
//   plane 0  exit_acc: stripe-local total of a cell draining across the stripe edge (0 elsewhere)
//   plane 1  link (low 32 bits: boundary index where flow ENTERING here leaves the stripe, -1 = ends inside)
//            | k << 32 (normalised code, 8 none, 9 inactive) | unres << 40 (the exit is locally unresolved)
__global__ void k_stripe_summary(Geo g, const uint32_t *__restrict__ meta, const int32_t *__restrict__ par,
                                 const unsigned long long *__restrict__ wq, const int32_t *__restrict__ plink,
                                 int64_t *__restrict__ recs)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 2 * g.nx) return;
    const int side = (int)(i / g.nx);
    const int64_t col = i - (int64_t)side * g.nx;
    int tile;
    const int64_t node = boundary_node(g, side, col, tile);
    const uint32_t m = meta[node];
    const bool active = (m >> 21) & 1;
    const int k = (m >> 16) & 0xF;
    const unsigned long long w = wq[node];
    const bool is_exit = active && par[node] == PAR_STRIPE &&
                         ((side == 0 && k >= 5 && k <= 7) || (side == 1 && k >= 1 && k <= 3));
    const bool resolved = w_pending(w) == 0u;
    const int64_t c_recv = col + (k < 8 ? d8_dc(k) : 0);
    const int64_t rec_k = !active ? 9 : ((c_recv < 0 || c_recv >= g.nx) ? 8 : k);
    const int64_t rec_unres = (is_exit && !resolved) ? 1 : 0;
    recs[i] = (is_exit && resolved) ? (int64_t)(uint32_t)w : 0;                        // plane 0: exit_acc (the finalised total)
    recs[2 * g.nx + i] = (int64_t)(uint32_t)plink[i] | (rec_k << 32) | (rec_unres << 40);   // plane 1: link | k | unres
}

template <typename T>
__global__ void k_widen(const int32_t *__restrict__ a, T *__restrict__ b, int64_t n)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) b[i] = (T)a[i];
}

// Route what arrives from the other stripes along the contracted paths: every entry on the path of a
// boundary cell gets the cell's external inflow added (linearity of the accumulation).  One thread per
// (hop, boundary cell); hop 0 also finishes a path that outgrew the recorded part.
template <typename T>
__global__ void k_stripe_route(Geo g, int64_t nnodes, const uint32_t *__restrict__ meta, const int2 *__restrict__ pe,
                               StripePaths sp, const int64_t *__restrict__ ext_inflow, const uint8_t *__restrict__ ext_un,
                               T *__restrict__ extra, uint8_t *__restrict__ extra_unres, uint8_t *__restrict__ tile_class,
                               uint32_t *__restrict__ flags, unsigned long long *wq)
{
    // 32-bit sums (VIA_WQ): the inflow goes into the forest word of every EXIT on the path -- phase C pulls it from there
    // like everything else, and only the boundary cell itself (the path's first entry) gets it through `extra`.  64-bit
    // sums (the whole raster holds 2^32 cells or more: the 32-bit sum field cannot take it): into `extra` of every entry.
    constexpr bool VIA_WQ = sizeof(T) == 4;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, n2 = 2 * g.nx;
    if (i >= n2) return;
    const int plen = sp.plen[i];
    if ((int)blockIdx.y >= plen) return;
    const int64_t v0 = ext_inflow[i];
    const bool u = ext_un && ext_un[i];
    if (v0 == 0 && !u) return;
    // A boundary cell fed by an exit that never finalises (a drainage cycle through several stripes) never finalises
    // either: it keeps what its finalised donors bring (v0, in `extra`), passes NOTHING on, and neither does any cell
    // further down its path.  The stripe-local pass could not know and finalised those exits: they are marked pending
    // again here, so that phase C's pull sees an unresolved donor instead of their (local) totals.
    const int64_t v = u ? 0 : v0;
    auto at_entry = [&](int64_t e, bool first) {
        if (first) inflow_add<T>(&extra[e], (T)v0);
        else if (!VIA_WQ && v) inflow_add<T>(&extra[e], (T)v);
        if (u) { extra_unres[e] = 1; tile_class[e / NSLOT] = 2; raise_flag(flags, F_UNRES); }
    };
    auto at_exit = [&](int64_t x) {
        if (x < 0) return;
        if (VIA_WQ && v) atomicAdd(&wq[x], (unsigned long long)(uint32_t)v);
        if (u) atomicOr(&wq[x], PEND1);
    };
    for (int h = blockIdx.y; h < plen; h += gridDim.y) {
        const int32_t item = sp.path[(int64_t)h * n2 + i];
        int64_t e = item & ~JUMP;
        at_entry(e, h == 0);
        int64_t x = (VIA_WQ || u || (item & JUMP)) ? first_exit(meta, e) : -1;
        at_exit(x);
        if (item & JUMP) {                                   // the entries the jump skipped inside its group of tile rows
            const int mr = macro_of(g, e);
            ChaseGuard cg;
            for (int64_t hops = 0; x >= 0 && hops <= nnodes; hops++) {
                if (cg.seen(x)) break;
                const int2 r = pe[x];
                const int32_t en = r.y;
                if (en < 0 || macro_of(g, en) != mr) break;
                at_entry(en, false);
                x = r.x;
                at_exit(x);
            }
        }
    }
    if (blockIdx.y == 0 && sp.pce[i] >= 0) {
        int64_t e = sp.pce[i], x = sp.pcx[i];
        ChaseGuard cg;
        for (int64_t hops = 0; hops <= nnodes; hops++) {
            if (cg.seen(e)) break;
            at_entry(e, false);
            at_exit(x);
            if (x < 0) break;
            const int2 r = pe[x];
            if (r.y < 0) break;
            e = r.y;
            x = r.x;
        }
    }
}

// Workspace layout.  Always: meta, the packed forest words, the per-tile arrays.  `wide` adds the arrays of the 64-bit
// forest engine (rasters of 2^32 cells and more), `stripe` the records the stripe kernels chase and the inflow that
// arrives from the other stripes.
struct SweepWs {
    uint32_t *meta; unsigned long long *wq; int32_t *par; uint8_t *tile_class, *tile_top; unsigned long long *row_plain, *gcnt;
    uint32_t *flags; uint8_t *gstate; uint32_t *codes4;                                                         // jump solver: the tiles' decoded codes, phase A -> phase C
    int32_t *exit_acc, *parent, *cnt; int64_t *A, *inflow; uint8_t *blocked, *ext_unres;   // wide
    int2 *pe; void *extra; uint8_t *extra_unres;                              // stripe
    StripePaths sp;
    size_t used;
    bool ok;
};
static int stripe_pcap(const Geo &g) { return g.tiles_y + 32 < 1024 ? g.tiles_y + 32 : 1024; }   // recorded hops per boundary path

static SweepWs carve_ws(void *ws, size_t ws_bytes, const Geo &g, bool wide, bool stripe)
{
    const size_t nt = (size_t)g.tiles_x * (size_t)g.tiles_y, nn = nt * NSLOT;
    WsCursor cur(ws, ws_bytes);
    SweepWs w{};
    w.meta = cur.take<uint32_t>(nn);
    w.wq = cur.take<unsigned long long>(nn);
    w.par = cur.take<int32_t>(nn);
    w.tile_class = cur.take<uint8_t>(nt);
    w.tile_top = cur.take<uint8_t>(nt);
    w.row_plain = cur.take<unsigned long long>(nt);
    w.gcnt = cur.take<unsigned long long>(GCNT_BYTES / sizeof(unsigned long long));
    w.flags = cur.take<uint32_t>(NFLAGS);                     // (right behind the counters: one memset clears both)
    w.gstate = cur.take<uint8_t>(nt);                         // band schedule: built / started marks (cleared by the same memset)
    if (!wide) {                                              // (the 64-bit forest engine keeps the chain-walk solver for such tiles)
        w.codes4 = cur.take<uint32_t>(nt * TH * 32);
    }
    if (wide) {
        w.exit_acc = cur.take<int32_t>(nn);
        w.parent = cur.take<int32_t>(nn);
        w.cnt = cur.take<int32_t>(nn);
        w.A = cur.take<int64_t>(nn);
        w.inflow = cur.take<int64_t>(nn);
        w.blocked = cur.take<uint8_t>(nn);
        w.ext_unres = cur.take<uint8_t>(nn);
    }
    w.sp = StripePaths{nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, 0};
    if (stripe) {
        w.pe = cur.take<int2>(nn);
        w.extra = cur.take<int64_t>(nn);
        w.extra_unres = cur.take<uint8_t>(nn);
        const size_t n2 = (size_t)(2 * g.nx);
        w.sp.pcap = stripe_pcap(g);
        w.sp.path = cur.take<int32_t>(n2 * (size_t)w.sp.pcap);
        w.sp.plen = cur.take<int32_t>(n2);
        w.sp.pce = cur.take<int32_t>(n2);
        w.sp.pcx = cur.take<int32_t>(n2);
        w.sp.plink = cur.take<int32_t>(n2);
        w.sp.mlink = cur.take<int32_t>((size_t)((g.tiles_y + MR - 1) / MR) * (size_t)g.nx);
    }
    w.used = cur.off;
    w.ok = cur.ok;
    return w;
}

static Geo make_geo(int64_t ny, int64_t nx, int64_t gy0, int64_t gny)
{
    Geo g;
    g.ny = ny; g.nx = nx; g.gy0 = gy0; g.gny = gny; g.extra_edge = 0; g.tile0 = 0;
    g.tiles_x = (int)((nx + TW - 1) / TW);
    g.tiles_y = (int)((ny + TH - 1) / TH);
    return g;
}

static int sm_count()
{
    static int cached[64];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (!cached[dev]) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}

template <int ENC>
int sweep_phaseA(const int8_t *dir, const uint8_t *mask, const Geo &g, int ntiles, const SweepWs &w, int al, cudaStream_t st,
                 bool fast = true, bool rare = true, size_t smem_pad = 0)
{   // al: 0 byte-wise staging (any width / alignment), 1 cp.async staging, 2 TMA bulk copies (see stage_mode)
    // fast: the one-pass kernel over the tiles [g.tile0, ntiles); rare: the jump / chain-walk solvers (all tiles)
    const unsigned fast_blocks = (unsigned)((ntiles - g.tile0 + WPB - 1) / WPB);
    const unsigned gen_blocks = (unsigned)(ntiles < sm_count() * 4 ? ntiles : sm_count() * 4);   // persistent: most tiles are skipped
    const size_t smemA = sizeof(int32_t) * TCELLS + 2 * TCELLS;
    {   // once per device and instantiation
        static bool done[64];
        int dev = 0;
        WMF_CUDA_CHECK(cudaGetDevice(&dev));
        if (dev < 0 || dev >= 64 || !done[dev]) {
            WMF_CUDA_CHECK(cudaFuncSetAttribute(k_generalA<ENC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemA));
            if (dev >= 0 && dev < 64) done[dev] = true;
        }
    }
    size_t smemF = (size_t)(mask ? 2 : 1) * WPB * 2 * CH * ROWW * sizeof(uint32_t) + (size_t)WPB * HBINS * sizeof(int);
    static_assert(2 * WPB * 2 * CH * ROWW * sizeof(uint32_t) + WPB * HBINS * sizeof(int) <= 48 * 1024, "phase A fits the default dynamic shared memory");
    if (smemF + smem_pad <= 48 * 1024) smemF += smem_pad;    // (unused shared memory: fewer resident blocks, see the band schedule)
  if (fast) {
#define WMF_LAUNCH_A(AL_, MK_, TM_)                                                                                              \
    k_sweepA<ENC, AL_, MK_, TM_><<<fast_blocks, WPB * 32, smemF, st>>>(dir, mask, g, ntiles, w.meta, w.exit_acc, w.tile_class, w.row_plain, \
                                                                       w.wq, w.tile_top, w.flags, w.codes4 != nullptr)
    if (mask) { if (al == 2) WMF_LAUNCH_A(true, true, true); else if (al) WMF_LAUNCH_A(true, true, false); else WMF_LAUNCH_A(false, true, false); }
    else { if (al == 2) WMF_LAUNCH_A(true, false, true); else if (al) WMF_LAUNCH_A(true, false, false); else WMF_LAUNCH_A(false, false, false); }
#undef WMF_LAUNCH_A
  }
  if (!rare) { WMF_LAUNCH_CHECK(); return 0; }
    if (w.codes4) {                                          // tiles with upward flow: pointer jumping, one block per tile
        const size_t smemJ = sizeof(uint32_t) * (TCELLS / 2 + TH * 32 + HBINS + 4);
        static_assert(sizeof(uint32_t) * (TCELLS / 2 + TH * 32 + HBINS + 4) <= 48 * 1024, "k_jumpA fits the default dynamic shared memory");
        const int cap = sm_count() * 4;
        k_jumpA<ENC><<<ntiles < cap ? ntiles : cap, JTA, smemJ, st>>>(dir, mask, g, ntiles, w.meta, w.exit_acc, w.tile_class, w.wq,
                                                                       w.tile_top, w.codes4, w.row_plain, w.flags, al != 0);
    }
    k_generalA<ENC><<<gen_blocks, GTH, smemA, st>>>(dir, mask, g, ntiles, w.tile_class, w.meta, w.exit_acc, w.wq, w.flags);
    WMF_LAUNCH_CHECK();
    return 0;
}

// wq: the packed forest words phase C pulls from (nullptr: everything comes from `extra`)
template <int ENC, typename T, typename TO>
int sweep_phaseC(const int8_t *dir, const uint8_t *mask, const Geo &g, int ntiles, const SweepWs &w, const unsigned long long *wq,
                 const T *extra, const uint8_t *extra_unres, int al, TO *acc, cudaStream_t st)
{
    const unsigned fast_blocks = (unsigned)((ntiles + WPB - 1) / WPB);
    const unsigned gen_blocks = (unsigned)(ntiles < sm_count() * 4 ? ntiles : sm_count() * 4);
    const size_t smemC = sizeof(T) * TCELLS + 2 * TCELLS;
    const size_t smemF = (size_t)(mask ? 2 : 1) * WPB * 2 * CH * ROWW * sizeof(uint32_t) + (size_t)WPB * NSLOT * sizeof(T) + WPB * 8 + 16 * WPB;
    const size_t smemJ = sizeof(T) * TCELLS * (sizeof(T) == 4 ? 2 : 1) + sizeof(uint32_t) * (TCELLS + TH * 32) + sizeof(T) * NSLOT + 16;   // k_jumpC
    {   // once per device and instantiation
        static bool done[64];
        int dev = 0;
        WMF_CUDA_CHECK(cudaGetDevice(&dev));
        if (dev < 0 || dev >= 64 || !done[dev]) {
            WMF_CUDA_CHECK(cudaFuncSetAttribute(k_generalC<ENC, T, TO>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemC));
            WMF_CUDA_CHECK(cudaFuncSetAttribute(k_jumpC<T, TO>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemJ));
            WMF_CUDA_CHECK(cudaFuncSetAttribute(k_sweepC<ENC, true, true, false, T, TO>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
            WMF_CUDA_CHECK(cudaFuncSetAttribute(k_sweepC<ENC, false, true, false, T, TO>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
            WMF_CUDA_CHECK(cudaFuncSetAttribute(k_sweepC<ENC, true, false, false, T, TO>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
            WMF_CUDA_CHECK(cudaFuncSetAttribute(k_sweepC<ENC, false, false, false, T, TO>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
            WMF_CUDA_CHECK(cudaFuncSetAttribute(k_sweepC<ENC, true, true, true, T, TO>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
            WMF_CUDA_CHECK(cudaFuncSetAttribute(k_sweepC<ENC, true, false, true, T, TO>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
            if (dev >= 0 && dev < 64) done[dev] = true;
        }
    }
#define WMF_LAUNCH_C(AL_, MK_, TM_) \
    k_sweepC<ENC, AL_, MK_, TM_, T, TO><<<fast_blocks, WPB * 32, smemF, st>>>(dir, mask, g, ntiles, wq, w.tile_top, extra, w.tile_class, w.row_plain, acc)
    if (mask) { if (al == 2) WMF_LAUNCH_C(true, true, true); else if (al) WMF_LAUNCH_C(true, true, false); else WMF_LAUNCH_C(false, true, false); }
    else { if (al == 2) WMF_LAUNCH_C(true, false, true); else if (al) WMF_LAUNCH_C(true, false, false); else WMF_LAUNCH_C(false, false, false); }
#undef WMF_LAUNCH_C
    if (w.codes4 && wq) {
        const int cap = sm_count();                          // one block per SM fits (shared memory)
        k_jumpC<T, TO><<<ntiles < cap ? ntiles : cap, JTC, smemJ, st>>>(g, ntiles, w.codes4, wq, w.tile_top, extra, w.tile_class, acc,
                                                                         w.row_plain, w.flags, al != 0);
    }
    k_generalC<ENC, T, TO><<<gen_blocks, GTH, smemC, st>>>(dir, mask, g, ntiles, w.tile_class, wq, w.par, extra, extra_unres, acc, w.flags);
    WMF_LAUNCH_CHECK();
    return 0;
}

// the packed forest (int32 sums): build, walk, flag what never finalised
static int sweep_phaseG_packed(const Geo &g, int ntiles, int64_t nnodes, const SweepWs &w, int2 *pe, cudaStream_t st)
{
    k_g_build<0><<<(ntiles + BU - 1) / BU, NSLOT, 0, st>>>(g, ntiles, w.meta, w.tile_top, w.par, pe, w.wq);
    k_g_walk<0><<<blocks_for(nnodes, 256), 256, 0, st>>>(nnodes, w.tile_top, w.par, w.wq, w.gcnt);
    k_g_unres<<<sm_count() * 8, 256, 0, st>>>(g, nnodes, w.tile_top, w.par, w.wq, w.tile_class, w.gcnt, w.flags);
    WMF_LAUNCH_CHECK();
    return 0;
}

static bool aligned4(const void *dir, const void *mask, const void *acc, int64_t nx)
{
    return (nx % NC == 0) && (((uintptr_t)dir & (NC - 1)) == 0) && (mask == nullptr || ((uintptr_t)mask & (NC - 1)) == 0) &&
           (((uintptr_t)acc & 31) == 0);
}
// How the fast kernels stage the DIR (and mask) rows of a tile: 0 byte by byte (any width, any alignment), 1 cp.async
// (rows of whole 8-byte pieces), 2 TMA bulk copies of whole 256-byte tile rows -- every tile is full width and every row
// piece 16-byte aligned.  Mode 2 is opt-in (WMF_D8_BULK=1 in the environment): measured on the 16384^2 raster it is 1 %
// SLOWER than cp.async (0.706 vs 0.698 ms; phase A 266 vs 257 us, 214 M vs 204 M warp instructions) -- these kernels are
// bound by instruction issue, and polling an mbarrier costs issue slots where cp.async.wait_group stalls for free.
static int stage_mode(const void *dir, const void *mask, const void *acc, int64_t nx)
{
    if (!aligned4(dir, mask, acc, nx)) return 0;
    const char *e = getenv("WMF_D8_BULK");
    if (!(e && e[0] == '1')) return 1;
    return (nx % TW == 0 && ((uintptr_t)dir & 15) == 0 && (mask == nullptr || ((uintptr_t)mask & 15) == 0)) ? 2 : 1;
}

// The only state the library keeps between calls: one helper stream and two events per device, created on first use;
// a per-device lock serialises the enqueue of wmf_d8_stripe_local calls from several host threads (they share the events).
struct SideStream { cudaStream_t stream; cudaEvent_t fork, join; std::mutex busy; };
static SideStream *side_stream()
{
    static SideStream tab[64];
    static bool made[64];
    static std::mutex mu;
    std::lock_guard<std::mutex> lock(mu);
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
    if (!made[dev]) {
        int lo = 0, hi = 0;                                  // highest priority: its few blocks go ahead of the walk's many
        if (cudaDeviceGetStreamPriorityRange(&lo, &hi) != cudaSuccess) return nullptr;
        if (cudaStreamCreateWithPriority(&tab[dev].stream, cudaStreamNonBlocking, hi) != cudaSuccess) return nullptr;
        if (cudaEventCreateWithFlags(&tab[dev].fork, cudaEventDisableTiming) != cudaSuccess) return nullptr;
        if (cudaEventCreateWithFlags(&tab[dev].join, cudaEventDisableTiming) != cudaSuccess) return nullptr;
        made[dev] = true;
    }
    return &tab[dev];
}

// ---- band schedule: phase G of one band of tile rows beside phase A of the next -------------------------------------
// Whole rasters, 32-bit sums.  Phases A and G used to be three kernels with the whole raster between them, and the walk
// is a latency chain (the longest river, ~0.5 us per tile crossed) during which the machine idles.  Here the raster is
// cut into B bands of tile rows: phase A of the bands runs on two alternating streams (so that a band fills the SMs the
// previous one's last blocks leave), and behind every band a high-priority stream builds the tiles whose 3 x 3
// neighbourhood is done and starts the walk of the tiles whose neighbourhood is built (k_g_build / k_g_walk, mode 1) --
// walkers stop at nodes of tiles that have not started, and whoever starts those carries on.  Only the last band's
// build and walk remain behind phase A.  Phase A is held to four blocks per SM (padding of unused shared memory) so
// that the small blocks of the forest kernels find registers beside it.  Tiles of the jump / chain-walk solvers get their
// meta words after all bands: they and their neighbours are built and started by the mode-2 launches that follow.
// OPT-IN (WMF_D8_BANDS=B in the environment), because the gain is small and has a narrow sweet spot: a first version for
// rasters of one-pass tiles only measured 0.700 -> 0.634 ms at 16384^2 with 8 bands (5 blocks per SM 0.662, 6 blocks
// 0.674; 4 / 12 / 16 bands 0.637 / 0.670 / 0.702); with the marks (+8 us) and the trailing launches (+16 us) that mixed
// rasters need, 8 bands take 0.678 ms against 0.695 unbanded, and 4 / 6 / 10 bands 0.695 / 0.687 / 0.704 -- how the block
// scheduler interleaves three streams is not something this code controls.  Bit-identical either way (test_band_schedule).
constexpr int MAXB = 16;
struct BandRes { cudaStream_t sA, sG; cudaEvent_t fork, evA[MAXB], evG; std::mutex busy; };
static BandRes *band_res()
{
    static BandRes tab[64];
    static bool made[64];
    static std::mutex mu;
    std::lock_guard<std::mutex> lock(mu);
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
    if (!made[dev]) {
        int lo = 0, hi = 0;
        if (cudaDeviceGetStreamPriorityRange(&lo, &hi) != cudaSuccess) return nullptr;
        if (cudaStreamCreateWithPriority(&tab[dev].sA, cudaStreamNonBlocking, lo) != cudaSuccess) return nullptr;
        if (cudaStreamCreateWithPriority(&tab[dev].sG, cudaStreamNonBlocking, hi) != cudaSuccess) return nullptr;
        if (cudaEventCreateWithFlags(&tab[dev].fork, cudaEventDisableTiming) != cudaSuccess) return nullptr;
        if (cudaEventCreateWithFlags(&tab[dev].evG, cudaEventDisableTiming) != cudaSuccess) return nullptr;
        for (int b = 0; b < MAXB; b++)
            if (cudaEventCreateWithFlags(&tab[dev].evA[b], cudaEventDisableTiming) != cudaSuccess) return nullptr;
        made[dev] = true;
    }
    return &tab[dev];
}
// bands for a raster of tiles_y tile rows: WMF_D8_BANDS in the environment (absent / 0 / 1 = off)
static int band_count(const Geo &g)
{
    const char *e = getenv("WMF_D8_BANDS");
    int b = e ? atoi(e) : 0;
    if (b > MAXB) b = MAXB;
    if (b > g.tiles_y) b = g.tiles_y;                        // (the tests force bands of a single tile row)
    return b < 2 ? 0 : b;
}
static size_t band_smem_pad()
{
    const char *e = getenv("WMF_D8_BAND_SMEM");
    return e ? (size_t)atoi(e) : 14000;
}

template <int ENC>
int sweep_phaseAG_bands(const int8_t *dir, const uint8_t *mask, const Geo &g, int ntiles, int64_t nnodes, const SweepWs &w, int al, int B,
                        cudaStream_t st)
{
    BandRes *br = band_res();
    WMF_REQUIRE(br != nullptr, (int)cudaErrorUnknown, "cannot create the band streams");
    std::lock_guard<std::mutex> lock(br->busy);              // (the streams and events are shared by the device's callers)
    const BandCtl band{w.tile_class, w.gstate, w.flags, 1}, rest{w.tile_class, w.gstate, w.flags, 2};
    WMF_CUDA_CHECK(cudaEventRecord(br->fork, st));
    WMF_CUDA_CHECK(cudaStreamWaitEvent(br->sA, br->fork, 0));
    WMF_CUDA_CHECK(cudaStreamWaitEvent(br->sG, br->fork, 0));
    const int ty_n = g.tiles_y;
    auto row_of = [&](int b) { return (int)((int64_t)ty_n * b / B); };
    for (int b = 0; b < B; b++) {
        const int r0 = row_of(b), r1 = row_of(b + 1);
        cudaStream_t sa = (b & 1) ? br->sA : st;
        Geo gb = g;
        gb.tile0 = r0 * g.tiles_x;
        int rc = sweep_phaseA<ENC>(dir, mask, gb, r1 * g.tiles_x, w, al, sa, true, false, band_smem_pad());
        if (rc) return rc;
        WMF_CUDA_CHECK(cudaEventRecord(br->evA[b], sa));
        WMF_CUDA_CHECK(cudaStreamWaitEvent(br->sG, br->evA[b], 0));
        const bool last = b == B - 1;
        const int b0 = r0 > 0 ? r0 - 1 : 0, b1 = last ? ty_n : r1 - 1;                  // rows to build: their neighbourhoods are done
        const int w0 = r0 > 1 ? r0 - 2 : 0, w1 = last ? ty_n : (r1 > 1 ? r1 - 2 : 0);   // rows to start: their neighbourhoods are built
        if (b1 > b0) {
            Geo gg = g;
            gg.tile0 = b0 * g.tiles_x;
            const int te = b1 * g.tiles_x;
            k_g_build<1><<<(te - gg.tile0 + BU - 1) / BU, NSLOT, 0, br->sG>>>(gg, te, w.meta, w.tile_top, w.par, nullptr, w.wq, band);
        }
        if (w1 > w0) {
            const int64_t n0 = (int64_t)w0 * g.tiles_x * NSLOT, n1 = (int64_t)w1 * g.tiles_x * NSLOT;
            k_g_walk<1><<<blocks_for(n1 - n0, 128), 128, 0, br->sG>>>(n1, w.tile_top, w.par, w.wq, w.gcnt, n0, g, band);
        }
    }
    WMF_CUDA_CHECK(cudaEventRecord(br->evG, br->sG));
    WMF_CUDA_CHECK(cudaStreamWaitEvent(st, br->evG, 0));     // (sG has waited for every band of both phase-A streams)
    // The jump / chain-walk solvers.  (Launched beside the last band's phase G -- they touch their own tiles only -- they
    // cost 10 us instead of saving some: even empty, their large blocks take the SM slots the walk is waiting for.)
    int rc = sweep_phaseA<ENC>(dir, mask, g, ntiles, w, al, st, false, true);
    if (rc) return rc;
    // whatever the bands had to leave out because of those solvers (nothing: the launches return at once)
    const int nblk = (ntiles + BU - 1) / BU, cap = sm_count() * 3;
    k_g_build<2><<<nblk < cap ? nblk : cap, NSLOT, 0, st>>>(g, ntiles, w.meta, w.tile_top, w.par, nullptr, w.wq, rest);
    k_g_walk<2><<<sm_count() * 8, 256, 0, st>>>(nnodes, w.tile_top, w.par, w.wq, w.gcnt, 0, g, rest);
    k_g_unres<<<sm_count() * 8, 256, 0, st>>>(g, nnodes, w.tile_top, w.par, w.wq, w.tile_class, w.gcnt, w.flags);
    WMF_LAUNCH_CHECK();
    return 0;
}

// T: the type the accumulation is computed in; TO: the type it is stored as (acc_out)
template <int ENC, typename T, typename TO>
int run_sweep_t(const int8_t *dir, const uint8_t *mask, TO *acc, int64_t ny, int64_t nx, void *ws, size_t ws_bytes,
                cudaStream_t st)
{
    const Geo g = make_geo(ny, nx, 0, ny);
    const int64_t ntiles64 = (int64_t)g.tiles_x * g.tiles_y;
    WMF_REQUIRE(ntiles64 * NSLOT < ((int64_t)1 << 31), WMF_E_TOO_LARGE, "raster too large for the tiled accumulation");
    const int ntiles = (int)ntiles64;
    const int64_t nnodes = ntiles64 * NSLOT;
    const bool wide = sizeof(T) == 8;
    const SweepWs w = carve_ws(ws, ws_bytes, g, wide, false);
    WMF_REQUIRE(w.ok, WMF_E_WORKSPACE, "workspace too small");
    const int al = stage_mode(dir, mask, acc, nx);
    if (!wide) {
        WMF_CUDA_CHECK(cudaMemsetAsync(w.gcnt, 0, (size_t)((char *)(w.gstate + ntiles) - (char *)w.gcnt), st));   // counters, flags, band marks
        int rc;
        if (const int B = band_count(g)) {                   // large rasters: the forest grows band by band behind phase A
            rc = sweep_phaseAG_bands<ENC>(dir, mask, g, ntiles, nnodes, w, al, B, st);
            if (rc) return rc;
            return sweep_phaseC<ENC, T, TO>(dir, mask, g, ntiles, w, w.wq, (const T *)nullptr, nullptr, al, acc, st);
        }
        rc = sweep_phaseA<ENC>(dir, mask, g, ntiles, w, al, st);
        if (rc) return rc;
        rc = sweep_phaseG_packed(g, ntiles, nnodes, w, nullptr, st);
        if (rc) return rc;
        return sweep_phaseC<ENC, T, TO>(dir, mask, g, ntiles, w, w.wq, (const T *)nullptr, nullptr, al, acc, st);
    }
    WMF_CUDA_CHECK(cudaMemsetAsync(w.inflow, 0, sizeof(int64_t) * nnodes, st));
    WMF_CUDA_CHECK(cudaMemsetAsync(w.ext_unres, 0, (size_t)nnodes, st));
    int rc = sweep_phaseA<ENC>(dir, mask, g, ntiles, w, al, st);
    if (rc) return rc;
    k_build_forest<<<blocks_for(nnodes, 256), 256, 0, st>>>(g, nnodes, w.meta, w.parent, w.blocked);
    WMF_LAUNCH_CHECK();
    rc = forest_accum_i64(w.parent, w.exit_acc, w.A, w.cnt, nnodes, st, w.blocked);
    if (rc) return rc;
    k_push_inflow<int64_t><<<blocks_for(nnodes, 256), 256, 0, st>>>(g, nnodes, w.meta, w.A, w.cnt, w.inflow, w.ext_unres, w.tile_class, w.flags);
    WMF_LAUNCH_CHECK();
    return sweep_phaseC<ENC, T, TO>(dir, mask, g, ntiles, w, nullptr, (const T *)w.inflow, w.ext_unres, al, acc, st);
}

// (WMF_ACC_WIDE=1 in the environment keeps 64-bit arithmetic for 64-bit outputs: the tests use it to cover that path)
static bool force_wide() { const char *e = getenv("WMF_ACC_WIDE"); return e && e[0] == '1'; }

// ---- stripe entry points --------------------------------------------------------------------------
// The stripe-local pass always runs the packed int32 forest walk (a stripe holds < 2^31 cells); only
// the finish pass is typed by the output accumulation.
template <int ENC>
int stripe_local_t(const int8_t *dir, const uint8_t *mask, int64_t ny, int64_t nx, int64_t row0, int64_t ny_total,
                   int64_t *recs, void *ws, size_t ws_bytes, cudaStream_t st)
{
    const Geo g = make_geo(ny, nx, row0, ny_total);
    const int64_t ntiles64 = (int64_t)g.tiles_x * g.tiles_y;
    WMF_REQUIRE(ntiles64 * NSLOT < ((int64_t)1 << 31), WMF_E_TOO_LARGE, "stripe too large for the tiled accumulation");
    const int ntiles = (int)ntiles64;
    const int64_t nnodes = ntiles64 * NSLOT;
    const SweepWs w = carve_ws(ws, ws_bytes, g, false, true);
    WMF_REQUIRE(w.ok, WMF_E_WORKSPACE, "workspace too small");
    const int al = stage_mode(dir, mask, nullptr, nx);
    WMF_CUDA_CHECK(cudaMemsetAsync(w.gcnt, 0, (size_t)((char *)(w.flags + NFLAGS) - (char *)w.gcnt), st));
    int rc = sweep_phaseA<ENC>(dir, mask, g, ntiles, w, al, st);
    if (rc) return rc;
    k_g_build<0><<<(ntiles + BU - 1) / BU, NSLOT, 0, st>>>(g, ntiles, w.meta, w.tile_top, w.par, w.pe, w.wq);
    // the path chase needs the forest only: it runs beside the walk (both are latency-bound chains)
    SideStream *ss = side_stream();
    WMF_REQUIRE(ss != nullptr, (int)cudaErrorUnknown, "cannot create the side stream");
    {
        std::lock_guard<std::mutex> lock(ss->busy);
        WMF_CUDA_CHECK(cudaEventRecord(ss->fork, st));
        WMF_CUDA_CHECK(cudaStreamWaitEvent(ss->stream, ss->fork, 0));
        // (the finish pass adds what the other stripes send into `extra`: cleared here, beside the walk, where DRAM idles)
        if (ny_total * nx <= 0xFFFFFFFFLL && !force_wide()) {
            // 32-bit sums (see wmf_d8_stripe_finish): the routed inflow goes into the forest words, `extra` (int32 there)
            // holds something only for the boundary cells themselves, in the stripe's first and last tile row
            const size_t rown = (size_t)g.tiles_x * NSLOT;
            WMF_CUDA_CHECK(cudaMemsetAsync(w.extra, 0, sizeof(int32_t) * rown, ss->stream));
            WMF_CUDA_CHECK(cudaMemsetAsync((int32_t *)w.extra + ((size_t)nnodes - rown), 0, sizeof(int32_t) * rown, ss->stream));
        } else
            WMF_CUDA_CHECK(cudaMemsetAsync(w.extra, 0, sizeof(int64_t) * (size_t)nnodes, ss->stream));
        WMF_CUDA_CHECK(cudaMemsetAsync(w.extra_unres, 0, (size_t)nnodes, ss->stream));
        k_macro_links<<<blocks_for((int64_t)((g.tiles_y + MR - 1) / MR) * nx, 128), 128, 0, ss->stream>>>(g, nnodes, w.meta, w.pe, w.sp.mlink);
        k_stripe_paths<<<blocks_for(2 * nx, 64), 64, 0, ss->stream>>>(g, nnodes, w.meta, w.pe, w.sp);
        WMF_CUDA_CHECK(cudaEventRecord(ss->join, ss->stream));
        k_g_walk<0><<<blocks_for(nnodes, 256), 256, 0, st>>>(nnodes, w.tile_top, w.par, w.wq, w.gcnt);
        k_g_unres<<<sm_count() * 8, 256, 0, st>>>(g, nnodes, w.tile_top, w.par, w.wq, w.tile_class, w.gcnt, w.flags);
        WMF_CUDA_CHECK(cudaStreamWaitEvent(st, ss->join, 0));
    }
    k_stripe_summary<<<blocks_for(2 * nx, 128), 128, 0, st>>>(g, w.meta, w.par, w.wq, w.sp.plink, recs);
    WMF_LAUNCH_CHECK();
    return 0;
}

template <int ENC, typename T, typename TO>
int stripe_finish_t(const int8_t *dir, const uint8_t *mask, TO *acc, int64_t ny, int64_t nx, int64_t row0,
                    int64_t ny_total, const int64_t *ext_inflow, const uint8_t *ext_un, void *ws, size_t ws_bytes,
                    cudaStream_t st)
{
    Geo g = make_geo(ny, nx, row0, ny_total);
    g.extra_edge = sizeof(T) == 4;                            // (see k_stripe_route)
    const int64_t ntiles64 = (int64_t)g.tiles_x * g.tiles_y;
    const int ntiles = (int)ntiles64;
    const int64_t nnodes = ntiles64 * NSLOT;
    const SweepWs w = carve_ws(ws, ws_bytes, g, false, true);
    WMF_REQUIRE(w.ok, WMF_E_WORKSPACE, "workspace too small");
    const int al = stage_mode(dir, mask, acc, nx);
    // phase-A state and the finalised forest words of the stripe-local pass are still in the workspace: what the other
    // stripes send is routed along the contracted paths into `extra`, and phase C adds it to what it pulls.
    T *extra = (T *)w.extra;                                 // (cleared by the local pass)
    k_stripe_route<T><<<dim3(blocks_for(2 * nx, 128), (unsigned)(w.sp.pcap < 32 ? w.sp.pcap : 32)), 128, 0, st>>>(g, nnodes, w.meta, w.pe, w.sp, ext_inflow, ext_un,
                                                                                          extra, w.extra_unres, w.tile_class, w.flags, w.wq);
    WMF_LAUNCH_CHECK();
    return sweep_phaseC<ENC, T, TO>(dir, mask, g, ntiles, w, w.wq, (const T *)extra, w.extra_unres, al, acc, st);
}

}  // namespace

// ---- basin membership: entry points for basin.cu ------------------------------------------------------
size_t d8_basin_membership_workspace(int64_t ny, int64_t nx)
{
    const Geo g = make_geo(ny, nx, 0, ny);
    const size_t nt = (size_t)g.tiles_x * (size_t)g.tiles_y, nn = nt * NSLOT;
    return 2 * wmf_align(nn * 4) + wmf_align(nt * TCELLS * 2) + wmf_align(MEMB_ROUNDS * sizeof(int)) + 1024;
}

// bm (nullable): 1 byte per cell; flag32 (nullable): the same as int32 (for a prefix sum); *count (device) += basin cells
int d8_basin_membership(const int8_t *dir, const uint8_t *mask, int64_t ny, int64_t nx, int encoding, int64_t outlet, uint8_t *bm,
                        int32_t *flag32, unsigned long long *count, void *ws, size_t ws_bytes, cudaStream_t st)
{
    const Geo g = make_geo(ny, nx, 0, ny);
    const int64_t ntiles64 = (int64_t)g.tiles_x * g.tiles_y;
    WMF_REQUIRE(ntiles64 * NSLOT < ((int64_t)1 << 31), WMF_E_TOO_LARGE, "raster too large for the tiled basin delineation");
    const int ntiles = (int)ntiles64;
    const int64_t nnodes = ntiles64 * NSLOT;
    WsCursor cur(ws, ws_bytes);
    uint32_t *meta = cur.take<uint32_t>((size_t)nnodes);
    int32_t *state = cur.take<int32_t>((size_t)nnodes);
    uint16_t *labels = cur.take<uint16_t>((size_t)ntiles * TCELLS);
    int *flags = cur.take<int>(MEMB_ROUNDS);
    WMF_REQUIRE(cur.ok, WMF_E_WORKSPACE, "workspace too small");
    const bool vec = aligned4(dir, mask, nullptr, nx);
    const size_t smemA = sizeof(uint32_t) * (TCELLS / 2 + TH * 32);
    const int cap = sm_count() * 4, blocksA = ntiles < cap ? ntiles : cap;
    WMF_CUDA_CHECK(cudaMemsetAsync(flags, 0, MEMB_ROUNDS * sizeof(int), st));
    if (encoding == WMF_ENC_INTERNAL) k_membA<WMF_ENC_INTERNAL><<<blocksA, JTA, smemA, st>>>(dir, mask, g, ntiles, outlet, meta, labels, vec);
    else k_membA<WMF_ENC_KEYPAD><<<blocksA, JTA, smemA, st>>>(dir, mask, g, ntiles, outlet, meta, labels, vec);
    k_memb_build<<<blocks_for(nnodes, 256), 256, 0, st>>>(g, nnodes, meta, state);
    const int64_t jb = blocks_for(nnodes, 256), jcap = (int64_t)sm_count() * 8;
    for (int r = 0; r < MEMB_ROUNDS; r++) k_memb_jump<<<(unsigned)(jb < jcap ? jb : jcap), 256, 0, st>>>(nnodes, state, flags, r);
    const int capC = sm_count() * 8;
    k_membC<<<ntiles < capC ? ntiles : capC, 256, 0, st>>>(g, ntiles, labels, state, bm, flag32, count);
    WMF_LAUNCH_CHECK();
    return 0;
}

int d8_convert_i64_to_f64(void *buf, int64_t n, cudaStream_t st);

size_t d8_sweep_workspace(int64_t ny, int64_t nx, int acc_dtype)
{
    // (64-bit outputs may take the wide engine: rasters of 2^32 cells and more, or WMF_ACC_WIDE)
    const Geo g = make_geo(ny, nx, 0, ny);
    const size_t narrow = carve_ws(nullptr, (size_t)-1, g, false, false).used;
    const size_t wide = acc_dtype != WMF_ACC_I32 ? carve_ws(nullptr, (size_t)-1, g, true, false).used : 0;
    return (narrow > wide ? narrow : wide) + 4096;
}

// How the tiles of the last accumulation that used this workspace were solved: counts[c] = tiles of class c
// (0 one-pass sweeps, 4 jump solver, 1 / 3 chain walk, 2 chain walk because of an unresolved upstream).  Syncs.
int d8_sweep_tile_classes(const void *ws, size_t ws_bytes, int64_t ny, int64_t nx, int64_t *counts8, cudaStream_t st)
{
    const Geo g = make_geo(ny, nx, 0, ny);
    const SweepWs w = carve_ws(const_cast<void *>(ws), ws_bytes, g, false, false);
    WMF_REQUIRE(w.ok, WMF_E_WORKSPACE, "workspace too small");
    const size_t nt = (size_t)g.tiles_x * (size_t)g.tiles_y;
    uint8_t *h = (uint8_t *)malloc(nt);
    WMF_REQUIRE(h != nullptr, WMF_E_ARG, "out of host memory");
    cudaError_t e = cudaMemcpyAsync(h, w.tile_class, nt, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) { free(h); wmf_set_error("d8_sweep_tile_classes: %s", cudaGetErrorString(e)); return (int)e; }
    for (int c = 0; c < 8; c++) counts8[c] = 0;
    for (size_t t = 0; t < nt; t++) counts8[h[t] & 7]++;
    // jump rounds of the class-4 tiles: [5] phase A total, [6] phase C total, [7] the largest of either
    unsigned long long *sw = (unsigned long long *)malloc(nt * 8);
    if (sw && cudaMemcpy(sw, w.row_plain, nt * 8, cudaMemcpyDeviceToHost) == cudaSuccess) {
        for (size_t t = 0; t < nt; t++)
            if (h[t] == 4) {
                const int64_t a = (int64_t)(sw[t] & 0xFFFFFFFFu), c = (int64_t)(sw[t] >> 32);
                counts8[5] += a; counts8[6] += c;
                if (a > counts8[7]) counts8[7] = a;
                if (c > counts8[7]) counts8[7] = c;
            }
    }
    free(sw);
    free(h);
    return 0;
}

// Dispatch on (encoding, arithmetic type, stored type).  A raster with fewer than 2^32 cells is accumulated in 32 bits
// whatever the output type: no accumulation can reach 2^32, so the unsigned value is exact and is widened (or turned
// into a double) only when it is stored.
// (WMF_ACC_WIDE=1 in the environment keeps 64-bit arithmetic for 64-bit outputs: the tests use it to cover that path)

#define WMF_ENC_DISPATCH(FN, T, TO, ...)                                                                     \
    (encoding == WMF_ENC_INTERNAL ? FN<WMF_ENC_INTERNAL, T, TO>(__VA_ARGS__) : FN<WMF_ENC_KEYPAD, T, TO>(__VA_ARGS__))

int d8_sweep(const int8_t *dir, const uint8_t *mask, void *acc, int acc_dtype, int64_t ny, int64_t nx, int encoding,
             void *ws, size_t ws_bytes, cudaStream_t st)
{
    const bool narrow = ny * nx <= 0xFFFFFFFFLL && !force_wide();
    if (acc_dtype == WMF_ACC_I32) return WMF_ENC_DISPATCH(run_sweep_t, int32_t, int32_t, dir, mask, (int32_t *)acc, ny, nx, ws, ws_bytes, st);
    if (acc_dtype == WMF_ACC_I64) {
        if (narrow) return WMF_ENC_DISPATCH(run_sweep_t, int32_t, int64_t, dir, mask, (int64_t *)acc, ny, nx, ws, ws_bytes, st);
        return WMF_ENC_DISPATCH(run_sweep_t, int64_t, int64_t, dir, mask, (int64_t *)acc, ny, nx, ws, ws_bytes, st);
    }
    if (narrow) return WMF_ENC_DISPATCH(run_sweep_t, int32_t, double, dir, mask, (double *)acc, ny, nx, ws, ws_bytes, st);
    int rc = WMF_ENC_DISPATCH(run_sweep_t, int64_t, int64_t, dir, mask, (int64_t *)acc, ny, nx, ws, ws_bytes, st);
    if (rc) return rc;
    return d8_convert_i64_to_f64(acc, ny * nx, st);
}

extern "C" {

size_t wmf_d8_stripe_workspace(int64_t ny, int64_t nx)
{
    return carve_ws(nullptr, (size_t)-1, make_geo(ny, nx, 0, ny), false, true).used + 4096;
}

int wmf_d8_stripe_local(const int8_t *dir, const uint8_t *mask, int64_t ny, int64_t nx, int64_t row0, int64_t ny_total,
                        int encoding, int64_t *records, void *workspace, size_t ws_bytes, void *stream)
{
    WMF_REQUIRE(dir && records && workspace, WMF_E_ARG, "null pointer");
    WMF_REQUIRE(ny >= 2 && nx > 0 && row0 >= 0 && row0 + ny <= ny_total, WMF_E_ARG, "bad stripe geometry (ny >= 2)");
    WMF_REQUIRE(2 * nx < ((int64_t)1 << 31) && ny * nx < ((int64_t)1 << 31), WMF_E_TOO_LARGE, "a stripe must hold fewer than 2^31 cells");
    WMF_REQUIRE(encoding == WMF_ENC_INTERNAL || encoding == WMF_ENC_KEYPAD, WMF_E_ARG, "unknown encoding");
    cudaStream_t st = (cudaStream_t)stream;
    return encoding == WMF_ENC_INTERNAL
               ? stripe_local_t<WMF_ENC_INTERNAL>(dir, mask, ny, nx, row0, ny_total, records, workspace, ws_bytes, st)
               : stripe_local_t<WMF_ENC_KEYPAD>(dir, mask, ny, nx, row0, ny_total, records, workspace, ws_bytes, st);
}

int wmf_d8_stripe_finish(const int8_t *dir, const uint8_t *mask, void *acc, int acc_dtype, int64_t ny, int64_t nx,
                         int64_t row0, int64_t ny_total, int encoding, const int64_t *ext_inflow,
                         const uint8_t *ext_unres, void *workspace, size_t ws_bytes, void *stream)
{
    WMF_REQUIRE(dir && acc && ext_inflow && workspace, WMF_E_ARG, "null pointer");
    WMF_REQUIRE(acc_dtype == WMF_ACC_I32 || acc_dtype == WMF_ACC_I64 || acc_dtype == WMF_ACC_F64, WMF_E_ARG, "unknown acc_dtype");
    WMF_REQUIRE(!(acc_dtype == WMF_ACC_I32 && ny_total * nx > 0x7fffffffLL), WMF_E_TOO_LARGE,
                "int32 accumulation can overflow beyond 2^31-1 cells: use WMF_ACC_I64");
    WMF_REQUIRE(ny >= 2 && nx > 0 && row0 >= 0 && row0 + ny <= ny_total, WMF_E_ARG, "bad stripe geometry (ny >= 2)");
    WMF_REQUIRE(encoding == WMF_ENC_INTERNAL || encoding == WMF_ENC_KEYPAD, WMF_E_ARG, "unknown encoding");
    cudaStream_t st = (cudaStream_t)stream;
    const bool narrow = ny_total * nx <= 0xFFFFFFFFLL && !force_wide();      // see d8_sweep
#define WMF_FINISH(T, TO) WMF_ENC_DISPATCH(stripe_finish_t, T, TO, dir, mask, (TO *)acc, ny, nx, row0, ny_total, ext_inflow, ext_unres, workspace, ws_bytes, st)
    if (acc_dtype == WMF_ACC_I32) return WMF_FINISH(int32_t, int32_t);
    if (acc_dtype == WMF_ACC_I64) return narrow ? WMF_FINISH(int32_t, int64_t) : WMF_FINISH(int64_t, int64_t);
    if (narrow) return WMF_FINISH(int32_t, double);
    int rc = WMF_FINISH(int64_t, int64_t);
#undef WMF_FINISH
    if (rc) return rc;
    return d8_convert_i64_to_f64(acc, ny * nx, st);
}

}  // extern "C"
