#!/usr/bin/env python
"""bench.py -- headline benchmark of the WMF raster-drainage hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--size 16384]

Primary metric (BASELINE.json): D8 flow-accumulation throughput in Mcells/s on a synthetic
16384 x 16384 Scheidegger DIR raster (int8 DIR resident in HBM -> int32 ACC resident in HBM; every
pass of the accumulation inside the timed region).  A "step" is one full accumulation of the raster.
The same JSON line carries
  parity_checked  the GPU accumulation of the benchmark raster equals the CPU oracle's, cell for cell (N = 1);
  secondary       the same accumulation, same size, on the other input classes of SURVEY 8(d): steepest-descent
                  terrain with all eight directions and pits, Scheidegger with an explicit all-true mask (what every
                  reference call passes) and with 5 % nodata -- with the share of tiles each tile solver took;
  shia            SHIA cell-timesteps/s (BASELINE configs 3 and 5): the Fortran 5-tank update with type-1 and type-2
                  speeds, the wmf_py 3-tank fp64 update, a population of 512 parameter sets, each with its roofline
                  figure, and the CPU restatement timed beside it.

N > 1 (torchrun, one rank per GPU): the raster is (N*size) x size, row-striped, one stripe per rank (weak scaling);
SHIA cells and parameter sets are sharded across the ranks with their gather inside the timed region.
Timing: CUDA events on the launching stream, barrier + synchronize on both sides, max over ranks.  Inputs (268 MB DIR +
1 GB ACC) exceed the 126 MB L2, so no explicit flush is needed between iterations (stated in config.l2).

`--impl reference` times the CPU restatement of the reference algorithm (oracle/, gcc -O2, 1 thread: the reference is
single-threaded by construction) on a bounded window of the same raster; it imports nothing of the product.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "D8 flow-accum Mcells/s (16k² raster); SHIA cell-timesteps/s at 1/2/4/8 B200"
SEED = 20260101


def stripe_rows_per_rank(n: int, world: int) -> int:
    """Rows of one rank's stripe.  N = 1: the n x n raster of BASELINE config 2.  N > 1: n - 64 rows (one tile row less)
    on EVERY rank count, so that the whole raster stays below 2^31 cells up to N = 8 and the accumulation type (int32)
    -- hence the bytes written per cell -- is the same at every N: the weak-scaling curve compares like with like.
    (The N = 8 raster of 8 x n x n cells is exactly 2^31, one more than int32 counts; it is timed beside it as
    `int64_variant`.)"""
    return n if world == 1 else n - 64


def workload_config(n: int, world: int, acc_name: str, algo: str = "auto") -> dict:
    """`config` of the JSON line: the same for both arms (the reference arm names its bounded sample in cpu_baseline)."""
    rows = stripe_rows_per_rank(n, world)
    return {"workload": f"D8 flow accumulation, synthetic Scheidegger DIR {rows}x{n} per GPU (seed {SEED}), "
                        f"int8 DIR -> {acc_name} ACC, mask = all valid", "algo": algo,
            "parallelism": "1 GPU" if world == 1 else
            f"{world} row stripes of {rows} rows of one {world * rows}x{n} raster, one NCCL all-gather of the "
            "boundary records per accumulation",
            "l2": "inputs (DIR 0.27 GB + ACC 1.07 GB) exceed the 126 MB L2; no explicit flush"}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def profiled(key: str):
    """Figures that only a profiler can give (DRAM bytes per accumulation, issue-slot utilisation), taken from the
    committed ncu captures of THIS build by tools/profile_figures.py -> profiles/r02_figures.json; None when absent."""
    p = os.path.join(ROOT, "profiles", "r02_figures.json")
    try:
        return json.load(open(p)).get(key)
    except Exception:
        return None


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region (NVML polled from a thread every ~2 ms;
    `nvidia-smi -lms` cannot deliver a sample inside a 30 ms region)."""

    def __init__(self, index: int):
        self.index, self.rows, self.stop, self.t, self.h = index, [], False, None, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(self._physical_index(index))
            self.max_sm = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.h = None

    @staticmethod
    def _physical_index(i: int) -> int:
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            try:
                return int(vis.split(",")[i])
            except Exception:
                return i
        return i

    def _loop(self):
        nv = self.nv
        while not self.stop:
            try:
                sm = float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    why = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                except Exception:
                    why = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                self.rows.append((sm, why))
            except Exception:
                pass
            time.sleep(0.002)

    def __enter__(self):
        if self.h is not None:
            self.t = threading.Thread(target=self._loop, daemon=True)
            self.t.start()
        return self

    def __exit__(self, *a):
        self.stop = True
        if self.t is not None:
            self.t.join(timeout=2)

    def summary(self):
        if self.h is None or not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        nv = self.nv
        bits = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
        reasons = sorted({nm for _, why in self.rows for nm, b in bits.items() if why & b})
        return {"sm_mhz": statistics.median(r[0] for r in self.rows), "sm_max_mhz": self.max_sm, "reasons": reasons,
                "samples": len(self.rows)}


# ------------------------------------------------------------------------------------------ CPU legs (oracle only)
def cpu_oracle_accum(rows: int, n: int, steps: int = 1):
    """Time the CPU restatement (oracle, 1 thread) on the first `rows` rows of the n x n benchmark raster (a true
    window of it: same generator, same keys).  Returns (Mcells/s, times, the accumulation of the window)."""
    import oracle
    fd = oracle.synth_scheidegger(rows, n, seed=SEED, nx_total=n)
    mask = np.ones((rows, n), dtype=bool)
    oracle.basin_acum(fd[:64, :64], mask[:64, :64])           # build/load outside the timed region
    times, acc = [], None
    for _ in range(steps):
        t0 = time.perf_counter()
        acc = oracle.basin_acum(fd, mask)
        times.append(time.perf_counter() - t0)
    return rows * n / 1e6 / statistics.median(times), times, acc


def run_reference(args, rank: int, world: int):
    """The reference arm: the reference's algorithm (wmf_py basin_acum, a sequential in-degree queue) as its C
    restatement, on the box's host cores.  The algorithm is single-threaded by construction, so it uses one."""
    if rank != 0:
        return
    n = args.size
    # rows of the raster one step accumulates: the whole raster (what the GPU arm accumulates per step, ~9 s) as long as
    # the run stays within a few minutes, else a true window of it (its first rows: Scheidegger flow never runs upwards)
    rows = min(args.ref_rows, n) if args.ref_rows > 0 else (n if args.steps <= 20 else max(64, (n * 20 // args.steps) // 64 * 64))
    for _ in range(1 if args.warmup else 0):
        cpu_oracle_accum(min(rows, 512), n, 1)
    t0 = time.perf_counter()
    mcells, times, _ = cpu_oracle_accum(rows, n, args.steps)
    line = {
        "impl": "reference", "metric": METRIC, "value": mcells, "unit": "Mcells/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * statistics.median(times),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
        "config": workload_config(n, max(args.gpus, 1), "int32"),
        "cpu_baseline": {"value": mcells, "unit": "Mcells/s", "cores": 1, "kind": "port",
                         "sample": (f"per step: the whole {n}x{n} benchmark raster; " if rows == n else
                                    f"per step: the first {rows} rows x {n} columns of the benchmark raster ({rows * n} cells); ") +
                                   "C restatement of wmf_py basin_acum (oracle/wmf_oracle.c, gcc -O2, 1 thread; the algorithm is "
                                   "a sequential queue), ~150x faster than the pure-Python reference, which cannot travel to this box"},
        "e2e": {"value": mcells, "unit": "Mcells/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": time.perf_counter() - t0,
    }
    print(json.dumps(line), flush=True)


def cpu_oracle_shia(n_cells: int, n_reg: int):
    """The Fortran 5-tank update as its C restatement, 1 thread, on a bounded sample (cells x steps)."""
    import oracle
    rng = np.random.default_rng(7)
    f32 = np.float32
    n_rec = 31
    rec = np.vstack([np.zeros((1, n_cells), np.int32), (1000 * rng.gamma(2.0, 2.0, (n_rec - 1, n_cells))).astype(np.int32)])
    pos = np.where(rng.random(n_reg) < 0.7, 1, rng.integers(2, n_rec + 1, n_reg)).astype(np.int32)
    inp = oracle.Shia5Inputs(
        v_coef=(rng.uniform(0.5, 1.5, (4, n_cells)) * np.array([[1e-2], [3e-3], [1e-3], [1e-4]])).astype(f32),
        h_coef=(rng.uniform(0.5, 1.5, (4, n_cells)) * np.array([[0.5], [0.05], [0.005], [1.0]])).astype(f32),
        h_exp=np.ones((4, n_cells), f32), max_capilar=rng.uniform(50, 150, n_cells).astype(f32),
        max_gravita=rng.uniform(10, 50, n_cells).astype(f32), max_aquifer=rng.uniform(200, 800, n_cells).astype(f32),
        hill_long=np.full(n_cells, 30.0, f32), elem_area=np.full(n_cells, 900.0, f32), rain_records=rec, pos_evento=pos,
        evp=(0.1 + 0.1 * np.sin(np.arange(n_reg) / 7.0)).astype(f32), dt=300.0, speed_type=(1, 1, 1))
    sto = np.full((5, n_cells), 0.001, f32)
    t0 = time.perf_counter()
    oracle.f_shia_v1(inp, np.ones(11, f32), sto, show_storage=False, show_mean_speed=False)
    dt = time.perf_counter() - t0
    return n_cells * n_reg / dt, dt


# ------------------------------------------------------------------------------------------ SHIA (config 3 / 5)
def shia_inputs(torch, dev, n_cells: int, n_reg: int, seed: int = 7):
    g = torch.Generator(device=dev).manual_seed(seed)
    u = lambda lo, hi, *s: lo + (hi - lo) * torch.rand(*s, device=dev, generator=g)
    cs = torch.empty((17, n_cells), device=dev)
    cs[0:4] = u(0.5, 1.5, 4, n_cells) * torch.tensor([[1e-2], [3e-3], [1e-3], [1e-4]], device=dev)
    cs[4:8] = u(0.5, 1.5, 4, n_cells) * torch.tensor([[0.5], [0.05], [0.005], [1.0]], device=dev)
    cs[8:12] = 1.0
    cs[12], cs[13], cs[14] = u(50, 150, n_cells), u(10, 50, n_cells), u(200, 800, n_cells)
    cs[15], cs[16] = 30.0, 900.0
    n_rec = 301
    rec = (1000.0 * torch.distributions.Gamma(2.0, 0.5).sample((n_rec, 1)).to(dev)
           * u(0.0, 1.0, n_rec, n_cells)).to(torch.int32)
    rec[0] = 0
    gp = torch.Generator().manual_seed(13)
    pos = torch.where(torch.rand(n_reg, generator=gp) < 0.7, torch.ones(n_reg, dtype=torch.int64),
                      torch.randint(2, n_rec + 1, (n_reg,), generator=gp)).to(torch.int32).to(dev)
    evp = (0.1 + 0.1 * torch.sin(torch.arange(n_reg) / 7.0)).to(dev)
    return cs, rec, pos, evp, float((pos != 1).float().mean())


def timed_runs(torch, fn, setup, steps: int, warmup: int, world: int, dist):
    """Median CUDA-event time [ms] of fn(state) over `steps` runs; max over ranks."""
    times = []
    for it in range(warmup + steps):
        state = setup()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn(state)
        e1.record()
        torch.cuda.synchronize()
        if it >= warmup:
            times.append(e0.elapsed_time(e1))
    ms = statistics.median(times)
    if world > 1:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return ms


def bench_shia(torch, ops, dev, args, rank: int, world: int, peak: float):
    """SHIA cell-timesteps/s.  N = 1: the runs of BASELINE config 3 (type-1 speeds, type-2 speeds, the fp64 3-tank model)
    and config 5 (512 parameter sets x 250 k cells).  N > 1: config 3 with the basin's cells sharded across the ranks
    (distinct cell ranges, per-step diagnostics gathered in rank order inside the timed region) and config 5 with the
    parameter sets sharded (balance series all-gathered inside the timed region)."""
    import torch.distributed as dist
    from wmf_heavy_b200 import _lib
    from wmf_heavy_b200 import dist as wdist
    out = {}
    n_cells, n_reg = args.shia_cells, args.shia_steps
    # ---- config 3: this rank's cells (the whole basin at N = 1)
    cs, rec, pos, evp, f_rain = shia_inputs(torch, dev, n_cells, n_reg, seed=7 + rank)
    calib = torch.ones((1, 11), device=dev)
    algo_bytes = 4.0 * f_rain + 108.0 / n_reg
    new_sto = lambda: torch.full((1, 5, n_cells), 0.001, device=dev)
    for name, st in (("c3_type1", (1, 1, 1)), ("c3_type2", (2, 2, 2))):
        if world == 1:
            run = lambda sto, st=st: ops.shia5_run(cs, rec, pos, evp, calib, sto, None, dt=300.0, speed_type=st, calc_niter=5)
        else:
            run = lambda sto, st=st: wdist.shia5_run_cell_sharded(cs, rec, pos, evp, calib, sto, None, dt=300.0, speed_type=st,
                                                                  calc_niter=5)
        ms = timed_runs(torch, run, new_sto, 3, 1, world, dist)
        v = world * n_cells * n_reg / (ms * 1e-3)
        gbs = algo_bytes * n_cells * n_reg / (ms * 1e-3) / 1e9
        out[name] = {"value": v, "unit": "cell-timesteps/s", "ms_per_run": ms, "cells": world * n_cells, "intervals": n_reg,
                     "mode": f"Fortran 5-tank as shipped, fp32, time-fused, speed_type {st}, calc_niter 5" +
                             ("" if world == 1 else f"; cells sharded over {world} ranks, diagnostics gathered (NCCL) in the timed region"),
                     "roofline": {"bound": "hbm", "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak,
                                  "algorithmic_bytes_per_cell_step": algo_bytes, "f_rain": f_rain,
                                  "issue_slot_utilisation": profiled(f"shia_{name}_issue_util"),
                                  "note": "off the HBM roof by construction: the time loop is fused (unfused: 112 B/cell-step); "
                                          "the kernel is instruction-issue bound"}}
    del cs, rec, pos, evp
    # ---- S4: the wmf_py 3-tank model, fp64, 8 B/cell-step of rain
    if world == 1:
        nt3 = 200
        rain = torch.rand((nt3, n_cells), device=dev, dtype=torch.float64) * 20.0
        prm = _lib.Shia3Params(300.0, 100.0, 30.0, 500.0, 0.0, 0.01, 0.02, 0.01, 0.005, 1.0e6, 300.0 / 3600.0, 0, 0)
        mk = lambda: tuple(torch.full((n_cells,), v, device=dev, dtype=torch.float64) for v in (10.0, 5.0, 50.0))
        ms = timed_runs(torch, lambda s: ops.shia3_run(rain, None, s[0], s[1], s[2], prm, nt3), mk, 3, 1, world, dist)
        gbs = 8.0 * n_cells * nt3 / (ms * 1e-3) / 1e9
        out["s4_fp64"] = {"value": n_cells * nt3 / (ms * 1e-3), "unit": "cell-timesteps/s", "ms_per_run": ms, "cells": n_cells,
                          "intervals": nt3, "mode": "wmf_py 3-tank balance, fp64, time-fused",
                          "roofline": {"bound": "hbm", "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak,
                                       "algorithmic_bytes_per_cell_step": 8.0,
                                       "issue_slot_utilisation": profiled("shia_s4_issue_util")}}
        del rain
    # ---- config 5: NSGA-II population, 512 parameter sets x 250 k cells (sets sharded across the ranks)
    P, n5 = args.pop_sets, args.pop_cells
    cs, rec, pos, evp, f_rain = shia_inputs(torch, dev, n5, n_reg, seed=11)
    gc = torch.Generator(device=dev).manual_seed(11)
    calibs = 0.5 + torch.rand((P, 11), device=dev, generator=gc)
    sto0 = torch.full((5, n5), 0.001, device=dev)
    if world == 1:
        run = lambda _s: ops.shia5_run(cs, rec, pos, evp, calibs, sto0.expand(P, 5, n5).contiguous(), None, dt=300.0)
    else:
        run = lambda _s: wdist.shia5_run_population_sharded(cs, rec, pos, evp, calibs, sto0, dt=300.0)
    ms = timed_runs(torch, run, lambda: None, 2, 1, world, dist)
    out["c5_population"] = {"value": P * n5 * n_reg / (ms * 1e-3), "unit": "cell-timesteps/s", "ms_per_run": ms, "sets": P,
                            "cells": n5, "intervals": n_reg,
                            "mode": f"{P} parameter sets x one basin, sets sharded over {world} rank(s)" +
                                    ("" if world == 1 else ", balance series all-gathered (NCCL) in the timed region")}
    return out


# ------------------------------------------------------------------------------------------ basin path (D2-D5)
def bench_basin(torch, ops, dev, peak: float, n: int = 2048):
    """Basin delineation and the vector-form routines at the size of BASELINE configs 1 / 3 / 5: the largest basin of a
    n x n Scheidegger raster (keypad codes), outlet at the cell of largest accumulation.  Algorithmic bytes:
    1 B per raster cell + 16 B per basin cell (12 B structure + 4 B acum)."""
    from wmf_heavy_b200 import _lib
    d = ops.synth_scheidegger(n, n, seed=SEED + 1)
    acc = ops.d8_accum(d, None)
    flat = int(torch.argmax(acc.view(-1)))
    r0, c0 = flat // n, flat % n
    key = ops.synth_scheidegger(n, n, seed=SEED + 1, encoding=_lib.ENC_KEYPAD)

    def timed(fn, reps=3):
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            r = fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps, r

    t_mask, (mask, nb) = timed(lambda: ops.basin_mask(d, None, (r0, c0)))
    t_st, st = timed(lambda: ops.basin_structure(key, c0 + 1, r0 + 1))
    nb = int(st.shape[1])
    dem = torch.rand((n, n), device=dev) + 100.0
    demv = ops.basin_map2basin(st, dem)
    dirv = ops.basin_map2basin(st, key).to(torch.int32)
    t_bas, bas = timed(lambda: ops.basin_basics(st, demv, dirv, 30.0, 0.0, 0.0, 30.0, n))
    t_nod, _ = timed(lambda: ops.stream_nod(st, bas[0], 100))
    by = 1.0 * n * n + 16.0 * nb
    gbs = by / (t_st * 1e-3) / 1e9
    return {"raster": f"{n}x{n} Scheidegger (keypad codes)", "basin_cells": nb,
            "basin_mask_ms": t_mask, "basin_structure_ms": t_st, "basin_basics_ms": t_bas, "stream_nod_ms": t_nod,
            "structure_basin_cells_per_s": nb / (t_st * 1e-3),
            "roofline": {"bound": "hbm", "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak,
                         "algorithmic_bytes": by,
                         "note": "basin_find + basin_cut (BFS-ordered structure): membership by the tile contraction (labels per tile "
                                 "in shared memory, pointer jumping over the tiles' exit cells), order by one radix sort; "
                                 "includes the host round trips of the call (outlet code, cell count)"}}


# ------------------------------------------------------------------------------------------ the GPU arm
def run_ours(args, rank: int, world: int, local_rank: int):
    import torch
    import torch.distributed as dist
    from wmf_heavy_b200 import _lib, ops, synth
    from wmf_heavy_b200.cu_py import basics

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    saved_stdout = None
    if world > 1:
        # NCCL may print its version banner on stdout; the driver wants ONE JSON line there
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        dist.init_process_group("nccl", device_id=dev)
    n = args.size
    peak, peak_src = measured_peaks()
    algo = {"auto": _lib.ALGO_AUTO, "walk": _lib.ALGO_WALK, "sweep": _lib.ALGO_SWEEP}[args.algo]

    # ---- workload: this rank's stripe of the (world*rows) x n Scheidegger raster, generated on the device
    rows = stripe_rows_per_rank(n, world)
    d = ops.synth_scheidegger(rows, n, seed=SEED, row0=rank * rows, nx_total=n)
    if world == 1:
        acc = torch.empty((n, n), dtype=torch.int32, device=dev)

        def step():
            ops.d8_accum(d, None, algo=algo, out=acc)
    else:
        # row stripes with one boundary exchange (all-gather over NCCL); int64 ACC when the whole raster can
        # exceed 2^31 - 1 cells
        from wmf_heavy_b200 import stripes
        wide = world * rows * n > 2**31 - 1
        acc = torch.empty((rows, n), dtype=torch.int64 if wide else torch.int32, device=dev)
        stripe_dtype = _lib.ACC_I64 if wide else _lib.ACC_I32

        def step():
            stripes.accumulate_striped(d, rank * rows, world * rows, acc_dtype=stripe_dtype, out=acc)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    with ClockSampler(local_rank) as clk:            # sampled over warm-up + timed region (the same load)
        for _ in range(max(args.warmup, 3)):
            step()
        barrier()
        ev[0].record()
        for i in range(args.steps):
            step()
            ev[i + 1].record()
        barrier()
    total_ms = ev[0].elapsed_time(ev[-1])
    per = [ev[i].elapsed_time(ev[i + 1]) for i in range(args.steps)]
    if world > 1:
        t = torch.tensor([total_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    ms_per_step = total_ms / args.steps
    cells_total = float(rows) * n * world
    value = cells_total / (ms_per_step * 1e-3) / 1e6
    op_ms = statistics.mean(per)
    acc_name = str(acc.dtype).replace("torch.", "")
    bytes_per_cell = 1.0 + acc.element_size()                       # int8 DIR read + ACC write
    achieved = bytes_per_cell * rows * n / (op_ms * 1e-3) / 1e9

    # ---- e2e through the reference-facing API: host flowdir + mask in (pinned), float64 raster out
    e2e = None
    if args.e2e_steps > 0 and world == 1:
        h_dir = torch.empty((n, n), dtype=torch.int8).pin_memory()
        h_dir.copy_(d)
        h_mask = torch.ones((n, n), dtype=torch.bool).pin_memory()
        h_out = torch.empty((n, n), dtype=torch.float64).pin_memory()
        fd_np, mk_np, out_np = h_dir.numpy(), h_mask.numpy(), h_out.numpy()
        basics.basin_acum(fd_np, mk_np, out=out_np)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            basics.basin_acum(fd_np, mk_np, out=out_np)
        torch.cuda.synchronize()
        e2e_s = (time.perf_counter() - t0) / args.e2e_steps
        assert float(out_np[-1, -1]) == float(acc[-1, -1].item())
        # the reference's own signature, no `out=`: the result arrives in a page-locked block of torch's host allocator
        r = basics.basin_acum(fd_np, mk_np)
        assert r.dtype == np.float64 and float(r[-1, -1]) == float(acc[-1, -1].item()) and float(r[n // 2, n // 3]) == float(acc[n // 2, n // 3].item())
        del r
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            r = basics.basin_acum(fd_np, mk_np)
            del r
        torch.cuda.synchronize()
        e2e_plain_s = (time.perf_counter() - t0) / args.e2e_steps
        e2e = {"value": cells_total / e2e_plain_s / 1e6, "unit": "Mcells/s", "h2d_bytes_per_step": 2 * n * n,
               "d2h_bytes_per_step": 8 * n * n, "ms_per_step": 1e3 * e2e_plain_s,
               "api": "wmf_heavy_b200.cu_py.basics.basin_acum(flowdir int8, mask bool) -> float64: exactly the reference's "
                      "signature (wmf_py/cu_py/basics.py:124); inputs in pinned host arrays, the result comes back in a "
                      "page-locked block of torch's caching host allocator",
               "with_out_buffer": {"value": cells_total / e2e_s / 1e6, "ms_per_step": 1e3 * e2e_s,
                                   "api": "the same with out=<pinned float64 buffer of the caller> (an extension of the signature)"}}
        del h_dir, h_mask, h_out

    if args.e2e_steps > 0 and world > 1:
        from wmf_heavy_b200 import stripes
        h_dir = torch.empty((rows, n), dtype=torch.int8).pin_memory()
        h_dir.copy_(d)
        h_out = torch.empty((rows, n), dtype=acc.dtype).pin_memory()

        def e2e_step():
            dd = h_dir.to(dev, non_blocking=True)
            a = stripes.accumulate_striped(dd, rank * rows, world * rows, acc_dtype=stripe_dtype)
            h_out.copy_(a)

        e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            e2e_step()
        torch.cuda.synchronize()
        e2e_s = (time.perf_counter() - t0) / args.e2e_steps
        t = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
        e2e = {"value": cells_total / e2e_s / 1e6, "unit": "Mcells/s", "h2d_bytes_per_step": rows * n,
               "d2h_bytes_per_step": acc.element_size() * rows * n, "ms_per_step": 1e3 * e2e_s,
               "api": "wmf_heavy_b200.stripes.accumulate_striped(host int8 stripe) -> host stripe, per rank"}
        del h_dir, h_out

    # ---- N > 1: the same stripes accumulated into int64 at the size where int32 no longer counts the raster (N * n * n
    # cells: 2^31 at N = 8), and BASELINE config 4 itself -- ONE 65536^2 raster over the N ranks (strong scaling)
    striped_extra = None
    if world > 1 and not args.no_secondary:
        from wmf_heavy_b200 import stripes
        striped_extra = {}

        def timed_stripes(fn, steps_):
            for _ in range(2):
                fn()
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps_):
                fn()
            e1.record()
            barrier()
            t = torch.tensor([e0.elapsed_time(e1) / steps_], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())

        d64 = ops.synth_scheidegger(n, n, seed=SEED, row0=rank * n, nx_total=n)
        a64 = torch.empty((n, n), dtype=torch.int64, device=dev)
        ms64 = timed_stripes(lambda: stripes.accumulate_striped(d64, rank * n, world * n, acc_dtype=_lib.ACC_I64, out=a64), 10)
        striped_extra["int64_variant"] = {"raster": f"{world * n}x{n}, {n} rows per rank", "dtype": "int64", "ms_per_step": ms64,
                                          "value": float(n) * n * world / (ms64 * 1e-3) / 1e6, "unit": "Mcells/s",
                                          "algorithmic_bytes_per_cell": 9.0}
        del d64, a64
        big = args.c4_size
        if big % world == 0 and (big // world) * big <= 2**33:
            rows_r = big // world
            n_local = max(1, -(-rows_r * big // (2**30)))          # stripes of at most 2^30 cells (< 2^31 each)
            while rows_r % n_local:
                n_local += 1
            dbig = ops.synth_scheidegger(rows_r, big, seed=SEED + 4, row0=rank * rows_r, nx_total=big)
            abig = torch.empty((rows_r, big), dtype=torch.int64, device=dev)
            msb = timed_stripes(lambda: stripes.accumulate_striped_multi(dbig, rank * rows_r, big, n_local, acc_dtype=_lib.ACC_I64,
                                                                         out=abig), 3)
            striped_extra["c4_strong"] = {"raster": f"{big}x{big} (BASELINE config 4), {rows_r} rows per rank in {n_local} stripe(s)",
                                          "dtype": "int64", "scaling": "strong", "ms_per_step": msb,
                                          "value": float(big) * big / (msb * 1e-3) / 1e6, "unit": "Mcells/s",
                                          "roofline_frac_per_gpu": 9.0 * rows_r * big / (msb * 1e-3) / 1e9 / peak}
            del dbig, abig
        torch.cuda.empty_cache()

    # ---- the other input classes, same size (N = 1)
    secondary = None
    if not args.no_secondary and world == 1:
        secondary = {}
        ter = synth.terrain_dir(n, n, seed=5)
        m1 = torch.ones((n, n), dtype=torch.uint8, device=dev)
        m95 = synth.nodata_mask(n, n, 0.05, seed=9)
        for name, dd, mm, desc in (
                ("terrain", ter, None, "steepest descent of a tilted plane + smooth multi-scale noise: all 8 directions, pits"),
                ("scheidegger_all_true_mask", d, m1, "the benchmark raster with an explicit all-true mask (what basin_acum(flowdir, mask) passes)"),
                ("scheidegger_5pct_nodata", d, m95, "the benchmark raster with 5 % of the cells masked out at random"),
                ("terrain_5pct_nodata", ter, m95, "the terrain raster with 5 % of the cells masked out")):
            st = {}
            for _ in range(2):
                ops.d8_accum(dd, mm, algo=algo, out=acc)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                ops.d8_accum(dd, mm, algo=algo, out=acc)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 5
            ops.d8_accum(dd, mm, algo=algo, out=acc, stats=st)
            bpc = 5.0 + (1.0 if mm is not None else 0.0)
            gbs = bpc * n * n / (ms * 1e-3) / 1e9
            secondary[name] = {"input": desc, "value": n * n / (ms * 1e-3) / 1e6, "unit": "Mcells/s", "ms_per_step": ms,
                               "vs_primary": (n * n / (ms * 1e-3) / 1e6) / value,
                               "roofline_frac": gbs / peak, "algorithmic_bytes_per_cell": bpc,
                               "tiles": st.get("tile_classes"), "rounds": st.get("jump_rounds")}
        del ter, m1, m95
        ops.d8_accum(d, None, algo=algo, out=acc)           # (the parity check below looks at the benchmark raster)

    shia = None
    if not args.no_shia:
        try:
            shia = bench_shia(torch, ops, dev, args, rank, world, peak)
        except Exception as e:           # the primary metric must still print
            shia = {"error": repr(e)}

    basin = None
    if not args.no_secondary and world == 1:
        try:
            basin = bench_basin(torch, ops, dev, peak)
        except Exception as e:
            basin = {"error": repr(e)}

    if rank == 0:
        cpu, parity = None, None
        if not args.no_cpu_baseline and world == 1:
            rows = args.cpu_rows or n                    # the whole raster by default: ~10 s of one host core
            mc, times, ref = cpu_oracle_accum(rows, n, 1)
            got = acc[:rows].cpu().numpy()
            parity = bool(np.array_equal(got, ref.astype(np.int32)))      # (the window's accumulation is exact: Scheidegger flow never runs upwards)
            cpu = {"value": mc, "unit": "Mcells/s", "cores": 1, "kind": "port",
                   "sample": (f"the whole {n}x{n} raster" if rows == n else f"the first {rows} rows of the same raster")
                             + ", C restatement of wmf_py basin_acum (oracle/wmf_oracle.c, gcc -O2, 1 thread); "
                             f"{times[0]:.2f} s"}
            del ref, got
            if shia is not None and "error" not in shia:
                v, dt_s = cpu_oracle_shia(1 << 20, 50)
                shia["cpu_baseline"] = {"value": v, "unit": "cell-timesteps/s", "cores": 1, "kind": "port",
                                        "sample": f"1 Mi cells x 50 intervals, C restatement of Modelos_heavy.f90 shia_v1 "
                                                  f"(oracle/wmf_oracle.c, 1 thread); {dt_s:.2f} s"}
        # kernels of libwmfb200.so per accumulation: sweepA, jumpA, generalA, g_build, g_walk, g_unres, sweepC, jumpC,
        # generalC; stripes = local (those six of phases A + G, macro_links, stripe_paths, stripe_summary) + boundary forest
        # (3) + finish (route, sweepC, jumpC, generalC)
        launches_per_step = ({"walk": 2, "sweep": 9, "auto": 9}[args.algo] if world == 1 else 16)
        line = {
            "metric": METRIC, "value": value, "unit": "Mcells/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": acc_name, "data": "synthetic",
            "config": workload_config(n, world, acc_name, args.algo),
            "parity_checked": parity,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         # dram__bytes_read + write over the kernels of one accumulation, from the ncu launch list of this
                         # build (profiles/r02_figures.json); None when that file does not match this size / algorithm
                         "traffic": profiled("d8_traffic_bytes") if (world == 1 and n == 16384 and args.algo != "walk") else None,
                         "peak_source": peak_src,
                         "kernel": "wmf_d8_accum, all passes of one accumulation "
                                   f"({bytes_per_cell:.0f} B/cell algorithmic)"},
            "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches_per_step * args.steps,
            "clocks": clk.summary(), "secondary": secondary, "striped": striped_extra, "basin": basin, "shia": shia,
        }
        if saved_stdout is not None:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--size", type=int, default=16384)
    ap.add_argument("--algo", default="auto", choices=["auto", "walk", "sweep"])
    ap.add_argument("--ref-rows", type=int, default=0,
                    help="rows of the raster one step of --impl reference accumulates (0: the whole raster up to 20 steps, else a window)")
    ap.add_argument("--cpu-rows", type=int, default=0, help="rows for cpu_baseline (0 = the whole raster)")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--shia-cells", type=int, default=1 << 20)
    ap.add_argument("--shia-steps", type=int, default=1000)
    ap.add_argument("--pop-sets", type=int, default=512)
    ap.add_argument("--pop-cells", type=int, default=250000)
    ap.add_argument("--c4-size", type=int, default=65536, help="side of the strong-scaling raster of BASELINE config 4 (N > 1)")
    ap.add_argument("--no-shia", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
