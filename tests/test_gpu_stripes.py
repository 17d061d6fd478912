"""Row-striped accumulation on ONE device (stripes processed one after the other through the same two
CUDA phases + boundary forest the multi-GPU path uses): bit-identical to the unstriped accumulation."""
import numpy as np
import pytest
import torch

import helpers
import oracle
from wmf_heavy_b200 import _lib, ops, stripes

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case,G", [("sch", 2), ("sch", 5), ("steep", 3), ("random", 4), ("mask", 3), ("keypad", 2),
                                    ("serpentine", 2), ("serpentine", 3)])
def test_striped_equals_oracle(case, G):
    mask = None
    enc = _lib.ENC_INTERNAL
    if case == "sch":
        fd = ops.synth_scheidegger_host(700, 515, seed=4)
    elif case == "steep":
        fd = helpers.steepest(533, 390, 5)
    elif case == "random":
        fd = helpers.random_codes(300, 260, 6)
    elif case == "serpentine":            # one path through every cell: boundary paths far longer than the recorded part
        fd = helpers.serpentine(260, 600)
    elif case == "mask":
        fd, mask = helpers.steepest(420, 333, 7), np.random.default_rng(7).random((420, 333)) > 0.06
    else:
        fd = helpers.steepest(400, 300, 8)
    exp = oracle.basin_acum(fd, np.ones(fd.shape, bool) if mask is None else mask).astype(np.int64)
    src = helpers.to_keypad(fd) if case == "keypad" else fd
    if case == "keypad":
        enc = _lib.ENC_KEYPAD
    d = torch.from_numpy(np.ascontiguousarray(src, dtype=np.int8)).cuda()
    m = None if mask is None else torch.from_numpy(mask.astype(np.uint8)).cuda()
    got = stripes.accumulate_striped_single_device(d, G, encoding=enc, mask_t=m)
    assert got.dtype == torch.int64
    assert np.array_equal(got.cpu().numpy(), exp), f"{(got.cpu().numpy() != exp).sum()} cells differ"


def test_cycle_through_two_stripes_and_two_tile_columns():
    """A drainage cycle that crosses a stripe boundary twice, through different tile columns: its cells never finalise
    (they keep what their finalised donors bring, without their own +1, and pass nothing on -- basics.py:166-175).  Each
    stripe alone sees an open chain and finalises it; the finish pass has to take that back along the whole path, also
    across tile edges.  (Found by tools/soak.py: random codes 1216 x 1320 in three stripes.)"""
    ny, nx = 12, 520
    fd = np.full((ny, nx), 2, np.int8)                 # everything runs south ...
    fd[-1, :] = -1                                     # ... into pits in the last row
    # the cycle: (5,256) -SW-> (6,255) -E-> (6,256) -N-> (5,256); rows 5 | 6 is the stripe boundary, columns 255 | 256 a tile edge
    fd[5, 256], fd[6, 255], fd[6, 256] = 3, 0, 6
    fd[5, 257] = 4                                     # a finalised donor of (5,256) from the east (column 257 above it runs south into it)
    fd[7, 255] = 2
    mask = np.ones((ny, nx), bool)
    exp = oracle.basin_acum(fd, mask).astype(np.int64)
    assert exp[5, 256] == exp[:5, 256].size + 6 and exp[6, 256] == 0     # (5 cells above + 6 through (5,257); no own +1)
    d = torch.from_numpy(fd).cuda()
    for G in (2, 3, 4):
        got = stripes.accumulate_striped_single_device(d, G).cpu().numpy()
        assert np.array_equal(got, exp), (G, np.argwhere(got != exp)[:5].tolist())
    # the case the soak found
    fd = helpers.random_codes(1216, 1320, 566697308)
    exp = oracle.basin_acum(fd, np.ones(fd.shape, bool)).astype(np.int64)
    got = stripes.accumulate_striped_single_device(torch.from_numpy(fd).cuda(), 3).cpu().numpy()
    assert np.array_equal(got, exp), np.argwhere(got != exp)[:5].tolist()


def test_striped_records_match_numpy_restatement():
    fd = ops.synth_scheidegger_host(130, 200, seed=11)
    rows = stripes.stripe_rows(130, 2)
    for r0, n in rows:
        packed, _ = ops.d8_stripe_local(torch.from_numpy(fd[r0:r0 + n].copy()).cuda(), r0, 130)
        rec = stripes.unpack_records(packed)
        acc_local = oracle.basin_acum(fd[r0:r0 + n], np.ones((n, 200), bool))
        exp = helpers.stripe_records_numpy(fd, r0, n, acc_local)
        for key in ("k", "exit_acc", "link", "unres"):
            assert np.array_equal(rec[key].cpu().numpy(), exp[key]), key
        assert torch.equal(stripes.pack_records(rec), packed)


def test_boundary_kernels_match_tensor_restatement():
    """The library's boundary forest kernels == the tensor-indexing restatement the CPU/gloo tests cover."""
    fd = helpers.random_codes(211, 157, 12, lo=0, hi=8)           # includes cycles across stripes
    d = torch.from_numpy(fd).cuda()
    rows = stripes.stripe_rows(211, 4)
    recs = torch.stack([ops.d8_stripe_local(d[r0:r0 + n].contiguous(), r0, 211)[0] for r0, n in rows])
    ref_in, ref_un = stripes.boundary_inflow([stripes.unpack_records(recs[g]) for g in range(4)])
    for g in range(4):
        inflow, unres = ops.stripe_boundary_inflow(recs, g)
        assert torch.equal(inflow, ref_in[g]) and torch.equal(unres, ref_un[g])


def test_striped_full_size_int64_and_f64():
    n = 4096
    d = ops.synth_scheidegger(n, n, seed=20260101)
    ref = ops.d8_accum(d, None, acc_dtype=_lib.ACC_I64)
    got = stripes.accumulate_striped_single_device(d, 4)
    assert torch.equal(got, ref)
    gf = stripes.accumulate_striped_single_device(d, 8, acc_dtype=_lib.ACC_F64)
    assert gf.dtype == torch.float64 and torch.equal(gf.to(torch.int64), ref)
    g32 = stripes.accumulate_striped_single_device(d, 3, acc_dtype=_lib.ACC_I32)
    assert g32.dtype == torch.int32 and torch.equal(g32.to(torch.int64), ref)


def test_striped_wide_arithmetic_path(monkeypatch):
    d = ops.synth_scheidegger(1024, 1024, seed=9)
    ref = ops.d8_accum(d, None, acc_dtype=_lib.ACC_I64)
    for wide in ("0", "1"):
        monkeypatch.setenv("WMF_ACC_WIDE", wide)
        assert torch.equal(stripes.accumulate_striped_single_device(d, 3), ref)
        gf = stripes.accumulate_striped_single_device(d, 2, acc_dtype=_lib.ACC_F64)
        assert torch.equal(gf.to(torch.int64), ref)


def test_forest_accum_matches_numpy():
    rng = np.random.default_rng(1)
    n = 5000
    parent = np.array([rng.integers(i + 1, n) if (i < n - 1 and rng.random() < 0.9) else -1 for i in range(n)], np.int32)
    parent[10], parent[20] = 20, 10                       # a two-cycle: both (and their downstream) stay unresolved
    w = rng.integers(0, 100, n).astype(np.int64)
    blocked = (rng.random(n) < 0.01).astype(np.uint8)
    acc, res = ops.forest_accum(torch.from_numpy(parent).cuda(), torch.from_numpy(w).cuda(), torch.from_numpy(blocked).cuda())
    e_acc, e_res = helpers.forest_numpy(parent, w, blocked)
    assert np.array_equal(res.cpu().numpy(), e_res)
    assert np.array_equal(acc.cpu().numpy(), e_acc)


def _band_consistent(d, a):
    """acc == 1 + sum of the donors' acc for the inner rows of a band of a Scheidegger raster (codes E, SE, S)."""
    a = a.to(torch.int64)
    tot = torch.ones_like(a)
    tot[:, 1:] += torch.where(d[:, :-1] == 0, a[:, :-1], torch.zeros_like(a[:, :-1]))          # from the west: E
    tot[1:, 1:] += torch.where(d[:-1, :-1] == 1, a[:-1, :-1], torch.zeros_like(a[:-1, :-1]))  # from the north-west: SE
    tot[1:, :] += torch.where(d[:-1, :] == 2, a[:-1, :], torch.zeros_like(a[:-1, :]))         # from the north: S
    return bool((tot[1:] == a[1:]).all())


def test_continental_65536_striped():
    """BASELINE config 4 at its full size, all stripes on one device: 65536^2 = 2^32 cells, eight row stripes of
    8192 rows, int64 accumulation through the 64-bit arithmetic path (the raster is one cell too large for 32 bits).
    Size-independent checks: a corner window against the oracle (nothing flows into it: the network only runs east and
    south), the donor recurrence on bands across stripe and tile edges, and mass conservation at the raster's exits."""
    free, _ = torch.cuda.mem_get_info()
    if free < 70 * 2**30:
        pytest.skip("needs ~50 GiB of device memory")
    n = 65536
    d = ops.synth_scheidegger(n, n, seed=20260101)
    acc = stripes.accumulate_striped_single_device(d, 8)
    assert acc.dtype == torch.int64 and acc.shape == (n, n)
    w = 1536
    sub = ops.synth_scheidegger_host(w, w, seed=20260101, nx_total=n)
    assert np.array_equal(sub, d[:w, :w].cpu().numpy())
    exp = oracle.basin_acum(sub, np.ones((w, w), bool)).astype(np.int64)
    assert np.array_equal(acc[:w, :w].cpu().numpy(), exp)
    for r0 in (8190, 32766, 57342, 65500):                 # stripe edges, and the raster's last rows
        r1 = min(r0 + 6, n)
        assert _band_consistent(d[r0:r1], acc[r0:r1]), f"rows {r0}..{r1}"
    last_col = (d[:, -1] == 0) | (d[:, -1] == 1)           # E / SE leave through the last column
    last_row = (d[-1, :] == 2) | (d[-1, :] == 1)           # S / SE leave through the last row
    total = int(acc[:, -1][last_col].sum()) + int(acc[-1, :-1][last_row[:-1]].sum())
    if bool(last_row[-1]) and not bool(last_col[-1]):
        total += int(acc[-1, -1])
    assert total == n * n
    assert int(acc.max()) <= n * n
