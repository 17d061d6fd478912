"""D8 accumulation parity: CUDA path (through the C ABI) vs the CPU oracle / golden vectors, bit-exact."""
import numpy as np
import pytest
import torch

import helpers
import oracle
from conftest import golden_names
from wmf_heavy_b200 import _lib, ops
from wmf_heavy_b200.cu_py import basics

pytestmark = pytest.mark.gpu
ALGOS = [_lib.ALGO_WALK, _lib.ALGO_SWEEP]


def run(fd, mask, algo, acc_dtype=_lib.ACC_I32, enc=_lib.ENC_INTERNAL):
    d = torch.from_numpy(np.ascontiguousarray(fd, dtype=np.int8)).cuda()
    m = None if mask is None else torch.from_numpy(np.ascontiguousarray(mask, dtype=np.uint8)).cuda()
    return ops.d8_accum(d, m, acc_dtype=acc_dtype, encoding=enc, algo=algo).cpu().numpy()


def test_reference_fixture_7x7():              # wmf_py/tests/test_cu_basics.py:35-58 through the drop-in API
    flowdir = np.zeros((7, 7), dtype=int)
    flowdir[:, -1] = 2
    flowdir[-1, -1] = -1
    mask = np.ones_like(flowdir, dtype=bool)
    acc = basics.basin_acum(flowdir, mask)
    expected = np.tile(np.arange(1, 8, dtype=float), (7, 1))
    expected[:, -1] = 7 * np.arange(1, 8)
    assert acc.dtype == np.float64 and np.array_equal(acc, expected)
    big = basics.basin_acum(np.zeros((200, 200), dtype=int), np.ones((200, 200), dtype=bool))
    assert np.array_equal(big, np.tile(np.arange(1, 201, dtype=float), (200, 1)))


@pytest.mark.parametrize("algo", ALGOS)
def test_quirks(algo):                         # SURVEY 8(c) known answers measured on the reference
    a = run(np.array([[0, 2, -1], [6, 4, 4], [0, 0, 6]]), None, algo)
    assert np.array_equal(a, [[0, 0, 1], [0, 4, 4], [1, 2, 3]])
    a = run(np.zeros((1, 5)), np.array([[1, 1, 0, 1, 1]]), algo)
    assert np.array_equal(a, [[1, 2, 0, 1, 2]])
    a = run(np.array([[0, 0, 8, 0, 0]]), None, algo)
    assert np.array_equal(a, [[1, 2, 3, 1, 2]])


@pytest.mark.parametrize("algo", ALGOS)
@pytest.mark.parametrize("name", golden_names("acum"))
def test_golden(golden, name, algo):
    fd, mask = golden[f"acum/{name}/flowdir"], golden[f"acum/{name}/mask"]
    assert np.array_equal(run(fd, mask, algo), golden[f"acum/{name}/acc"])
    assert np.array_equal(basics.basin_acum(fd.astype(np.int64), mask), golden[f"acum/{name}/acc"].astype(float))


CASES = {
    "sch_1024": lambda: (ops.synth_scheidegger_host(1024, 1024, seed=5), None),
    "sch_ragged": lambda: (ops.synth_scheidegger_host(517, 1301, seed=6), None),
    "steep_700x513": lambda: (helpers.steepest(700, 513, 7), None),
    "steep_mask": lambda: (helpers.steepest(389, 641, 8), np.random.default_rng(8).random((389, 641)) > 0.05),
    "random_cycles": lambda: (helpers.random_codes(311, 277, 9), np.random.default_rng(9).random((311, 277)) > 0.1),
    "serpentine": lambda: (helpers.serpentine(96, 131), None),
    "all_east": lambda: (np.zeros((130, 4099), np.int8), None),
    "all_north": lambda: (np.full((1500, 70), 6, np.int8), None),
    "all_west_sw": lambda: (np.where(np.random.default_rng(3).random((257, 400)) < 0.5, 4, 3).astype(np.int8), None),
    "thin_col": lambda: (helpers.steepest(3000, 1, 10), None),
    "thin_row": lambda: (helpers.steepest(1, 3000, 11), None),
    "one_cell": lambda: (np.array([[0]], np.int8), None),
    "all_masked_out": lambda: (helpers.steepest(40, 50, 12), np.zeros((40, 50), bool)),
}


@pytest.mark.parametrize("algo", ALGOS)
@pytest.mark.parametrize("case", sorted(CASES))
def test_vs_oracle(case, algo):
    fd, mask = CASES[case]()
    m = np.ones(fd.shape, bool) if mask is None else mask
    exp = oracle.basin_acum(fd, m)
    got = run(fd, mask, algo)
    assert got.dtype == np.int32
    assert np.array_equal(got, exp.astype(np.int32)), f"{(got != exp).sum()} cells differ"


@pytest.mark.parametrize("algo", ALGOS)
def test_dtypes_and_keypad(algo):
    fd = helpers.steepest(300, 420, 13)
    exp = oracle.basin_acum(fd, np.ones(fd.shape, bool))
    g64 = run(fd, None, algo, _lib.ACC_I64)
    assert g64.dtype == np.int64 and np.array_equal(g64, exp.astype(np.int64))
    gf = run(fd, None, algo, _lib.ACC_F64)
    assert gf.dtype == np.float64 and np.array_equal(gf, exp)
    gk = run(helpers.to_keypad(fd), None, algo, _lib.ACC_I32, _lib.ENC_KEYPAD)
    assert np.array_equal(gk, exp.astype(np.int32))


def test_wide_arithmetic_path(monkeypatch):
    """64-bit outputs of rasters below 2^32 cells are accumulated in 32 bits and widened on store; WMF_ACC_WIDE=1 keeps
    the 64-bit arithmetic (the path of larger rasters).  Both must agree with the oracle."""
    fd, mask = CASES["random_cycles"]()
    exp = oracle.basin_acum(fd, mask)
    sch = ops.synth_scheidegger_host(700, 1100, seed=3)
    exp2 = oracle.basin_acum(sch, np.ones(sch.shape, bool))
    for wide in ("0", "1"):
        monkeypatch.setenv("WMF_ACC_WIDE", wide)
        assert np.array_equal(run(fd, mask, _lib.ALGO_SWEEP, _lib.ACC_I64), exp.astype(np.int64))
        assert np.array_equal(run(fd, mask, _lib.ALGO_SWEEP, _lib.ACC_F64), exp)
        assert np.array_equal(run(sch, None, _lib.ALGO_SWEEP, _lib.ACC_I64), exp2.astype(np.int64))
        assert np.array_equal(run(sch, None, _lib.ALGO_SWEEP, _lib.ACC_F64), exp2)


def test_tma_bulk_staging(monkeypatch):
    """WMF_D8_BULK=1 stages the tile rows with TMA bulk copies (cp.async.bulk + mbarrier) instead of cp.async: rasters
    whose width is a multiple of the tile width, with and without a mask, ragged in height, plus a width that falls back."""
    rng = np.random.default_rng(5)
    for ny, nx in ((300, 512), (64, 256), (131, 1024), (200, 520)):
        sch = ops.synth_scheidegger_host(ny, nx, seed=11)
        mask = rng.random((ny, nx)) > 0.05
        exp = oracle.basin_acum(sch, np.ones(sch.shape, bool))
        expm = oracle.basin_acum(sch, mask)
        for bulk in ("1", "0"):
            monkeypatch.setenv("WMF_D8_BULK", bulk)
            assert np.array_equal(run(sch, None, _lib.ALGO_SWEEP, _lib.ACC_I32), exp.astype(np.int32)), (ny, nx, bulk)
            assert np.array_equal(run(sch, mask, _lib.ALGO_SWEEP, _lib.ACC_I64), expm.astype(np.int64)), (ny, nx, bulk)
    fd, m = CASES["random_cycles"]()
    monkeypatch.setenv("WMF_D8_BULK", "1")
    assert np.array_equal(run(fd, m, _lib.ALGO_SWEEP, _lib.ACC_I64), oracle.basin_acum(fd, m).astype(np.int64))


@pytest.mark.parametrize("bands", ["2", "3", "16"])
def test_band_schedule(monkeypatch, bands):
    """WMF_D8_BANDS switches the band schedule on (phase G of a band of tile rows beside phase A of the next): bands of a few -- or single --
    tile rows, every tile solver present (cycles, upward flow, masks), ragged sizes.  Bit-identical to the oracle and to
    the unbanded schedule."""
    rng = np.random.default_rng(7)
    cases = [(name, *CASES[name]()) for name in ("random_cycles", "steep_700x513", "steep_mask", "all_north", "serpentine", "sch_ragged")]
    sch = ops.synth_scheidegger_host(900, 700, seed=5)
    cases.append(("sch", sch, None))
    cases.append(("sch_mask", sch, rng.random(sch.shape) > 0.04))
    mix = ops.synth_scheidegger_host(1100, 600, seed=6).copy()
    mix[300:420, 100:500] = helpers.random_codes(120, 400, 9)            # a block of random codes (cycles) inside a river network
    mix[700:760, :] = helpers.steepest(60, 600, 10)
    cases.append(("mix", mix, None))
    for name, fd, mask in cases:
        m = np.ones(fd.shape, bool) if mask is None else mask
        exp = oracle.basin_acum(fd, m)
        monkeypatch.setenv("WMF_D8_BANDS", bands)
        got = run(fd, mask, _lib.ALGO_SWEEP, _lib.ACC_I64)
        monkeypatch.setenv("WMF_D8_BANDS", "0")
        ref = run(fd, mask, _lib.ALGO_SWEEP, _lib.ACC_I64)
        assert np.array_equal(got, exp.astype(np.int64)), (name, bands)
        assert np.array_equal(ref, got), (name, bands)


def test_accumulations_beyond_int32_in_32_bit_arithmetic():
    """A raster of 2^31 + 65536 cells that drains through ONE cell: every row runs east, the last column south.  Its
    64-bit output is computed in 32-bit (unsigned) arithmetic because the raster has fewer than 2^32 cells; the last
    column holds (r + 1) * nx, up to 2 147 549 184 > INT32_MAX.  Closed-form check of every cell."""
    free, _ = torch.cuda.mem_get_info()
    if free < 40 * 2**30:
        pytest.skip("needs ~25 GiB of device memory")
    ny, nx = 32769, 65536
    d = torch.zeros((ny, nx), dtype=torch.int8, device="cuda")
    d[:, -1] = 2
    with pytest.raises(_lib.WmfError):
        ops.d8_accum(d, None, acc_dtype=_lib.ACC_I32)          # int32 cannot hold it: refused, not wrapped
    acc = ops.d8_accum(d, None, acc_dtype=_lib.ACC_I64)
    cols = torch.arange(1, nx, dtype=torch.int64, device="cuda")
    for r0 in range(0, ny, 4096):                              # (in bands: the expectation is not materialised)
        band = acc[r0:r0 + 4096]
        assert bool((band[:, :-1] == cols).all()), f"rows from {r0}"
        rows = torch.arange(r0 + 1, r0 + band.shape[0] + 1, dtype=torch.int64, device="cuda")
        assert torch.equal(band[:, -1], rows * nx), f"last column, rows from {r0}"
    assert int(acc[-1, -1]) == ny * nx > 2**31
    from wmf_heavy_b200 import stripes                        # and across stripe boundaries (external inflow > INT32_MAX)
    assert torch.equal(stripes.accumulate_striped_single_device(d, 3), acc)


def test_keypad_equals_internal_at_size():
    """The same Scheidegger raster in both encodings (lean rows translate keypad codes with a byte-permute table)."""
    n = 2048
    a = ops.d8_accum(ops.synth_scheidegger(n, n, seed=77), None)
    k = ops.d8_accum(ops.synth_scheidegger(n, n, seed=77, encoding=_lib.ENC_KEYPAD), None, encoding=_lib.ENC_KEYPAD)
    assert torch.equal(a, k)
    m = (torch.rand((n, n), device="cuda") > 0.02).to(torch.uint8)          # and with a mask (no lean rows)
    am = ops.d8_accum(ops.synth_scheidegger(n, n, seed=77), m)
    km = ops.d8_accum(ops.synth_scheidegger(n, n, seed=77, encoding=_lib.ENC_KEYPAD), m, encoding=_lib.ENC_KEYPAD)
    assert torch.equal(am, km) and not torch.equal(am, a)
    assert torch.equal(ops.d8_accum(ops.synth_scheidegger(n, n, seed=77), m, algo=_lib.ALGO_WALK), am)


def test_empty_and_errors():
    assert basics.basin_acum(np.zeros((0, 5), int), np.zeros((0, 5), bool)).shape == (0, 5)
    d = torch.zeros((4, 4), dtype=torch.int8, device="cuda")
    with pytest.raises(_lib.WmfError):
        ops.d8_accum(d, None, encoding=7)


def _local_consistency(dir_t, acc):
    """Size-independent property: acc[c] == 1 + sum of acc over the neighbours draining into c, for an
    acyclic all-active raster; and the accumulation at the terminal cells adds up to the cell count."""
    ny, nx = dir_t.shape
    total = torch.ones_like(acc, dtype=torch.int64)
    for k in range(8):
        dr, dc = int(helpers.DR[k]), int(helpers.DC[k])
        src = torch.where(dir_t == k, acc, torch.zeros_like(acc)).to(torch.int64)
        # cell (r, c) with code k drains into (r+dr, c+dc): shift the source by (dr, dc), dropping wraparound
        sh = torch.zeros_like(src)
        r0, r1 = max(dr, 0), ny + min(dr, 0)
        c0, c1 = max(dc, 0), nx + min(dc, 0)
        sh[r0:r1, c0:c1] = src[r0 - dr:r1 - dr, c0 - dc:c1 - dc]
        total += sh
        del src, sh
    return bool((total == acc.to(torch.int64)).all())


@pytest.mark.parametrize("n", [4096, 16384])
def test_full_size_properties(n):
    """BASELINE config 2 (16384^2 Scheidegger): local consistency everywhere + mass conservation, and
    the two algorithms agree bit for bit."""
    d = ops.synth_scheidegger(n, n, seed=20260101)
    acc = ops.d8_accum(d, None, algo=_lib.ALGO_SWEEP)
    assert _local_consistency(d, acc)
    # terminal cells: E on the last column, S on the last row, SE on either
    term = torch.zeros_like(d, dtype=torch.bool)
    term[:, -1] |= (d[:, -1] == 0) | (d[:, -1] == 1)
    term[-1, :] |= (d[-1, :] == 2) | (d[-1, :] == 1)
    assert int(acc[term].to(torch.int64).sum()) == n * n
    if n <= 4096:
        acc2 = ops.d8_accum(d, None, algo=_lib.ALGO_WALK)
        assert torch.equal(acc, acc2)
        sub = ops.synth_scheidegger_host(n, n, seed=20260101)
        assert np.array_equal(sub, d.cpu().numpy())


@pytest.mark.parametrize("enc", [_lib.ENC_INTERNAL, _lib.ENC_KEYPAD])
@pytest.mark.parametrize("kind", ["all_true", "nodata5", "pits", "nodata_rows"])
def test_masked_lean_rows_vs_oracle(kind, enc):
    """Rasters WITH a mask keep the lean rows (every reference call passes one): an explicit all-true mask, 5 % nodata,
    pits / nodata codes inside E/SE/S/SW rows, and whole inactive rows and columns; bit-exact against the oracle at
    a size where every tile kind occurs (interior, raster edge, ragged)."""
    ny, nx = 1100, 1700
    rng = np.random.default_rng(31)
    fd = ops.synth_scheidegger_host(ny, nx, seed=31)
    fd[rng.random((ny, nx)) < 0.02] = 3                                     # some SW too
    mask = np.ones((ny, nx), bool)
    if kind == "nodata5":
        mask = rng.random((ny, nx)) > 0.05
    elif kind == "pits":
        fd[rng.random((ny, nx)) < 0.01] = -1
        fd[rng.random((ny, nx)) < 0.01] = 8
        mask = rng.random((ny, nx)) > 0.01
    elif kind == "nodata_rows":
        mask[200:203] = False
        mask[:, 700:702] = False
        mask[63:65, :300] = False
        mask[:, 255:257] = False
    exp = oracle.basin_acum(fd, mask)
    fd_in = helpers.to_keypad(fd) if enc == _lib.ENC_KEYPAD else fd
    got = run(fd_in, mask, _lib.ALGO_SWEEP, enc=enc)
    assert np.array_equal(got, exp.astype(np.int32)), f"{(got != exp).sum()} cells differ"
    from wmf_heavy_b200 import stripes
    d = torch.from_numpy(np.ascontiguousarray(fd_in, dtype=np.int8)).cuda()
    m = torch.from_numpy(np.ascontiguousarray(mask, dtype=np.uint8)).cuda()
    st = stripes.accumulate_striped_single_device(d, 3, encoding=enc, mask_t=m).cpu().numpy()
    assert np.array_equal(st, exp.astype(np.int64))
