"""Seeded synthetic rasters for the parity tests (numpy, host side)."""
import numpy as np

DR = np.array([0, 1, 1, 1, 0, -1, -1, -1])
DC = np.array([1, 1, 0, -1, -1, -1, 0, 1])
INT2KEY = np.array([6, 3, 2, 1, 4, 7, 8, 9], dtype=np.int8)   # internal k -> keypad code


def steepest(ny, nx, seed, pit_frac=0.005, noise=3.0, tilt=(0.7, 0.4)):
    """D8 steepest descent towards increasing (row, col) of a tilted noisy plane; -1 at pits."""
    rng = np.random.default_rng(seed)
    r, c = np.mgrid[0:ny, 0:nx]
    z = tilt[0] * r + tilt[1] * c + noise * rng.standard_normal((ny, nx))
    zp = np.pad(z, 1, constant_values=-np.inf)
    best = np.full((ny, nx), -1, dtype=np.int8)
    bestdrop = np.zeros((ny, nx))
    for d in range(8):
        nb = zp[1 + DR[d]:1 + DR[d] + ny, 1 + DC[d]:1 + DC[d] + nx]
        drop = (nb - z) / np.hypot(DR[d], DC[d])
        take = drop > bestdrop
        best[take] = d
        bestdrop[take] = drop[take]
    best[rng.random((ny, nx)) < pit_frac] = -1
    return best


def random_codes(ny, nx, seed, lo=-1, hi=9):
    return np.random.default_rng(seed).integers(lo, hi, size=(ny, nx)).astype(np.int8)


def serpentine(ny, nx):
    """One path visiting every cell: rows alternate E / W, turning S at the ends."""
    fd = np.zeros((ny, nx), dtype=np.int8)
    fd[1::2, :] = 4
    fd[0::2, -1] = 2
    fd[1::2, 0] = 2
    last = ny - 1
    fd[last, -1 if last % 2 == 0 else 0] = -1
    return fd


def to_keypad(fd_internal):
    """internal 0..7 -> keypad; anything else -> 0 (no receiver)."""
    out = np.zeros(fd_internal.shape, dtype=np.int8)
    ok = (fd_internal >= 0) & (fd_internal <= 7)
    out[ok] = INT2KEY[fd_internal[ok]]
    return out


def shia5_inputs(oracle_mod, N, T, seed, n_rec=7, speed_type=(1, 1, 1), dt=300.0):
    rng = np.random.default_rng(seed)
    f32 = np.float32
    rec = np.vstack([np.zeros((1, N), np.int32), (1000 * rng.gamma(2.0, 2.0, (n_rec - 1, N))).astype(np.int32)])
    pos = np.where(rng.random(T) < 0.6, 1, rng.integers(2, n_rec + 1, T)).astype(np.int32)
    return oracle_mod.Shia5Inputs(
        v_coef=(rng.uniform(0.5, 1.5, (4, N)) * np.array([[1e-2], [3e-3], [1e-3], [1e-4]])).astype(f32),
        h_coef=(rng.uniform(0.5, 1.5, (4, N)) * np.array([[0.5], [0.05], [0.005], [1.0]])).astype(f32),
        h_exp=rng.uniform(0.8, 1.2, (4, N)).astype(f32),
        max_capilar=rng.uniform(50, 150, N).astype(f32), max_gravita=rng.uniform(10, 50, N).astype(f32),
        max_aquifer=rng.uniform(200, 800, N).astype(f32), hill_long=np.full(N, 30.0, f32),
        elem_area=np.full(N, 900.0, f32), rain_records=rec, pos_evento=pos,
        evp=(0.1 + 0.1 * np.sin(np.arange(T) / 7.0)).astype(f32), dt=dt, speed_type=speed_type)


def cell_static(inp):
    return np.vstack([inp.v_coef, inp.h_coef, inp.h_exp, inp.max_capilar[None], inp.max_gravita[None],
                      inp.max_aquifer[None], inp.hill_long[None], inp.elem_area[None]]).astype(np.float32)


# ---------------------------------------------------------------- stripes: numpy restatement of the records
def stripe_records_numpy(fd, r0, n, acc_local):
    """Boundary records of stripe rows [r0, r0+n) of the all-active raster `fd` (internal codes), as
    wmf_d8_stripe_local defines them: k / exit_acc / unres / link, each [2, nx].  `acc_local` is the
    accumulation of the stripe in isolation (oracle).  Acyclic inputs only (unres = 0)."""
    ny_total, nx = fd.shape
    sub = fd[r0:r0 + n]
    k = np.full((2, nx), 8, np.int8)
    exit_acc = np.zeros((2, nx), np.int64)
    link = np.full((2, nx), -1, np.int32)
    for side, lr in ((0, 0), (1, n - 1)):
        for c in range(nx):
            code = int(sub[lr, c])
            if 0 <= code <= 7:
                rr, cc = r0 + lr + DR[code], c + DC[code]
                if 0 <= rr < ny_total and 0 <= cc < nx:
                    k[side, c] = code
            kk = k[side, c]
            if (side == 0 and 5 <= kk <= 7) or (side == 1 and 1 <= kk <= 3):
                exit_acc[side, c] = int(acc_local[lr, c])
            # follow the path of flow entering here until it leaves the stripe
            r, cc = lr, c
            for _ in range(n * nx + 1):
                code = int(sub[r, cc])
                if not (0 <= code <= 7):
                    break
                r2, c2 = r + DR[code], cc + DC[code]
                if not (0 <= r0 + r2 < ny_total and 0 <= c2 < nx):
                    break
                if r2 < 0:
                    link[side, c] = cc
                    break
                if r2 >= n:
                    link[side, c] = nx + cc
                    break
                r, cc = r2, c2
    return {"k": k, "exit_acc": exit_acc, "link": link, "unres": np.zeros((2, nx), np.uint8)}


def forest_numpy(parent, weight, blocked):
    """acc[i] = weight[i] + sum over children; nodes below a blocked / cyclic node stay unresolved."""
    n = len(parent)
    cnt = np.zeros(n, np.int64)
    for p in parent:
        if p >= 0:
            cnt[p] += 1
    cnt += (np.asarray(blocked) != 0)
    acc = np.zeros(n, np.int64)
    done = np.zeros(n, np.uint8)
    stack = [i for i in range(n) if cnt[i] == 0]
    for i in stack:
        acc[i] += weight[i]
        done[i] = 1
    while stack:
        i = stack.pop()
        p = parent[i]
        if p < 0:
            continue
        acc[p] += acc[i]
        cnt[p] -= 1
        if cnt[p] == 0:
            acc[p] += weight[p]
            done[p] = 1
            stack.append(p)
    return acc, done


def expected_stripe_inflow(fd, acc_full, rows):
    """[G, 2, nx]: what enters each stripe's first / last row from the neighbouring stripes."""
    ny, nx = fd.shape
    out = np.zeros((len(rows), 2, nx), np.int64)
    for g, (r0, n) in enumerate(rows):
        for side, r in ((0, r0), (1, r0 + n - 1)):
            for c in range(nx):
                for kk in range(8):
                    rr, cc = r - DR[kk], c - DC[kk]          # donor position for code kk
                    outside = rr < r0 if side == 0 else rr >= r0 + n
                    if outside and 0 <= rr < ny and 0 <= cc < nx and fd[rr, cc] == kk:
                        out[g, side, c] += int(acc_full[rr, cc])
    return out
