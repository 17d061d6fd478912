"""Randomised soak of the D8 paths against the CPU oracle (more seeds / shapes than the test suite):
    python tools/soak.py [rounds]
accumulation (both algorithms, masks, cycles, keypad codes, the optional schedules), the same in row stripes on one device,
basin_cut at random outlets, the BFS-ordered structure."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import torch
import helpers
import oracle
from wmf_heavy_b200 import _lib, ops, stripes
from wmf_heavy_b200.cu_py import basics

rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 12
rng = np.random.default_rng(2026)
bad = 0
for it in range(rounds):
    ny, nx = int(rng.integers(1, 1400)), int(rng.integers(1, 1400))
    kind = ("steep", "random", "sch", "mix")[it % 4]
    seed = int(rng.integers(1 << 30))
    if kind == "steep":
        fd = helpers.steepest(ny, nx, seed)
    elif kind == "random":
        fd = helpers.random_codes(ny, nx, seed)
    elif kind == "sch":
        fd = ops.synth_scheidegger_host(ny, nx, seed=seed)
    else:
        fd = ops.synth_scheidegger_host(ny, nx, seed=seed).copy()
        r0, c0 = int(rng.integers(ny)), int(rng.integers(nx))
        r1, c1 = min(ny, r0 + int(rng.integers(1, 200))), min(nx, c0 + int(rng.integers(1, 400)))
        fd[r0:r1, c0:c1] = helpers.random_codes(r1 - r0, c1 - c0, seed + 1)
    mask = None if it % 3 == 0 else rng.random((ny, nx)) > float(rng.choice([0.0, 0.02, 0.3]))
    m = np.ones((ny, nx), bool) if mask is None else mask
    exp = oracle.basin_acum(fd, m)
    d = torch.from_numpy(fd).cuda()
    mt = None if mask is None else torch.from_numpy(mask.astype(np.uint8)).cuda()
    for algo in (_lib.ALGO_SWEEP, _lib.ALGO_WALK):
        got = ops.d8_accum(d, mt, acc_dtype=_lib.ACC_I64, algo=algo).cpu().numpy()
        if not np.array_equal(got, exp.astype(np.int64)):
            bad += 1
            print("MISMATCH accum", it, kind, ny, nx, algo)
    for bands in ("2", "5"):                                  # the optional band schedule
        os.environ["WMF_D8_BANDS"] = bands
        got = ops.d8_accum(d, mt, acc_dtype=_lib.ACC_I64, algo=_lib.ALGO_SWEEP).cpu().numpy()
        os.environ["WMF_D8_BANDS"] = "0"
        if not np.array_equal(got, exp.astype(np.int64)):
            bad += 1
            print("MISMATCH bands", it, kind, ny, nx, bands)
    os.environ["WMF_D8_BULK"] = "1"                           # TMA bulk-copy staging
    got = ops.d8_accum(d, mt, acc_dtype=_lib.ACC_I32, algo=_lib.ALGO_SWEEP).cpu().numpy()
    os.environ["WMF_D8_BULK"] = "0"
    if not np.array_equal(got, exp.astype(np.int32)):
        bad += 1
        print("MISMATCH bulk", it, kind, ny, nx)
    key = torch.from_numpy(helpers.to_keypad(fd)).cuda()
    got = ops.d8_accum(key, mt, acc_dtype=_lib.ACC_F64, encoding=_lib.ENC_KEYPAD).cpu().numpy()
    if not np.array_equal(got, exp.astype(np.float64)):
        bad += 1
        print("MISMATCH keypad/f64", it, kind, ny, nx)
    if kind in ("steep", "sch") and mask is None:            # BFS-ordered structure (the Fortran flavour has no mask)
        r0, c0 = np.unravel_index(np.argmax(exp), exp.shape)
        est = oracle.f_basin_find(helpers.to_keypad(fd), int(c0) + 1, int(r0) + 1)
        st = ops.basin_structure(key, int(c0) + 1, int(r0) + 1).cpu().numpy()
        if not np.array_equal(st, est):
            bad += 1
            print("MISMATCH structure", it, kind, ny, nx)
    for G in (2, 3, 5):
        if ny >= 2 * G:
            got = stripes.accumulate_striped_single_device(d, G, mask_t=mt, acc_dtype=_lib.ACC_I64).cpu().numpy()
            if not np.array_equal(got, exp.astype(np.int64)):
                bad += 1
                print("MISMATCH stripes", it, kind, ny, nx, G)
    dem = rng.random((ny, nx))
    cand = [np.unravel_index(np.argmax(exp), exp.shape)] + [(int(rng.integers(ny)), int(rng.integers(nx))) for _ in range(3)]
    for r0, c0 in cand:
        if not m[r0, c0]:
            continue
        e = oracle.basin_cut(dem, (int(r0), int(c0)), fd.astype(np.int64), m)["mask"]
        g = basics.basin_cut(dem, (int(r0), int(c0)), fd, m)["mask"]
        if not np.array_equal(e, g):
            bad += 1
            print("MISMATCH basin_cut", it, kind, ny, nx, r0, c0, int(e.sum()), int(g.sum()))
    print(f"round {it}: {kind} {ny}x{nx} mask={'none' if mask is None else round(float(m.mean()), 2)} ok", flush=True)
print("soak:", "FAILED" if bad else "OK", f"({bad} mismatches)")
sys.exit(1 if bad else 0)
