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
