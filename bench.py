#!/usr/bin/env python
"""bench.py -- headline benchmark of the WMF raster-drainage hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--size 16384]

Primary metric (BASELINE.json): D8 flow-accumulation throughput in Mcells/s on a synthetic
16384 x 16384 Scheidegger DIR raster (int8 DIR resident in HBM -> int32 ACC resident in HBM; every
pass of the accumulation inside the timed region).  A "step" is one full accumulation of the raster.
Secondary (reported in the same line under "shia"): SHIA cell-timesteps/s on a synthetic 1 Mi-cell
basin x 1000 intervals (BASELINE config 3).

N > 1 (torchrun, one rank per GPU): the raster is (N*size) x size, row-striped, one stripe per rank
(weak scaling); see DESIGN.md "Multi-GPU".  Timing: CUDA events on the launching stream, barrier +
synchronize on both sides, max over ranks.  Inputs (268 MB DIR + 1 GB ACC) exceed the 126 MB L2, so no
explicit flush is needed between iterations (stated in config.l2).

`--impl reference` times the CPU restatement of the reference algorithm (oracle/, gcc -O2, 1 thread: the
reference is single-threaded by construction) on a bounded window of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "D8 flow-accum Mcells/s (16k² raster); SHIA cell-timesteps/s at 1/2/4/8 B200"
SEED = 20260101
ALGO_BYTES_PER_CELL = 5.0      # 1 B int8 DIR read + 4 B int32 ACC write


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


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


def cpu_oracle_accum(sample: int, steps: int = 1):
    """Time the CPU restatement (oracle, 1 thread) on the top-left sample x sample window of the workload."""
    import oracle
    from wmf_heavy_b200 import ops
    fd = ops.synth_scheidegger_host(sample, sample, seed=SEED, nx_total=sample)
    mask = np.ones((sample, sample), dtype=bool)
    oracle.basin_acum(fd[:64, :64], mask[:64, :64])           # build/load outside the timed region
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        acc = oracle.basin_acum(fd, mask)
        times.append(time.perf_counter() - t0)
    return sample * sample / 1e6 / statistics.median(times), times, acc


def run_reference(args, rank: int, world: int):
    if rank != 0:
        return
    sample = args.ref_sample
    for _ in range(args.warmup and 1):
        cpu_oracle_accum(min(sample, 1024), 1)
    t0 = time.perf_counter()
    mcells, times, _ = cpu_oracle_accum(sample, args.steps)
    line = {
        "impl": "reference", "metric": METRIC, "value": mcells, "unit": "Mcells/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * statistics.median(times),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
        "config": {"workload": f"D8 accumulation, Scheidegger DIR (seed {SEED}), {sample}x{sample} window of the "
                               f"{args.size}x{args.size} raster", "sample_cells": sample * sample},
        "cpu_baseline": {"value": mcells, "unit": "Mcells/s", "cores": 1, "kind": "port",
                         "sample": f"{sample}x{sample} top-left window, C restatement of wmf_py basin_acum "
                                   "(oracle/wmf_oracle.c, gcc -O2), the Python reference cannot travel to this box"},
        "e2e": {"value": mcells, "unit": "Mcells/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": time.perf_counter() - t0,
    }
    print(json.dumps(line), flush=True)


def bench_shia(torch, ops, dev, n_cells: int, n_reg: int, steps: int, warmup: int):
    """SHIA 5-tank (Fortran semantics, as shipped), BASELINE config 3: ~1 Mi cells x 1000 intervals."""
    g = torch.Generator(device=dev).manual_seed(7)
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
    calib = torch.ones((1, 11), device=dev)
    f_rain = float((pos != 1).float().mean())
    times = []
    for it in range(warmup + steps):
        sto = torch.full((1, 5, n_cells), 0.001, device=dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        ops.shia5_run(cs, rec, pos, evp, calib, sto, None, dt=300.0, speed_type=(1, 1, 1))
        e1.record()
        torch.cuda.synchronize()
        if it >= warmup:
            times.append(e0.elapsed_time(e1))
    ms = statistics.median(times)
    return {"value": n_cells * n_reg / (ms * 1e-3), "unit": "cell-timesteps/s", "ms_per_run": ms, "cells": n_cells,
            "intervals": n_reg, "mode": "Fortran 5-tank as shipped, fp32, time-fused, speed_type (1,1,1)",
            "algorithmic_bytes_per_cell_step": 4.0 * f_rain + 108.0 / n_reg, "f_rain": f_rain}


def run_ours(args, rank: int, world: int, local_rank: int):
    import torch
    import torch.distributed as dist
    from wmf_heavy_b200 import _lib, ops
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

    # ---- workload: this rank's stripe of the (world*n) x n Scheidegger raster, generated on the device
    d = ops.synth_scheidegger(n, n, seed=SEED, row0=rank * n, nx_total=n)
    if world == 1:
        acc = torch.empty((n, n), dtype=torch.int32, device=dev)

        def step():
            ops.d8_accum(d, None, algo=algo, out=acc)
    else:
        # row stripes with one boundary exchange (all-gather over NCCL); int64 ACC: the whole raster can
        # exceed 2^31 cells
        from wmf_heavy_b200 import stripes
        wide = world * n * n > 2**31 - 1
        acc = torch.empty((n, n), dtype=torch.int64 if wide else torch.int32, device=dev)
        stripe_dtype = _lib.ACC_I64 if wide else _lib.ACC_I32

        def step():
            stripes.accumulate_striped(d, rank * n, world * n, acc_dtype=stripe_dtype, out=acc)

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
    cells_total = float(n) * n * world
    value = cells_total / (ms_per_step * 1e-3) / 1e6
    op_ms = statistics.mean(per)
    bytes_per_cell = 1.0 + acc.element_size()                       # int8 DIR read + ACC write
    achieved = bytes_per_cell * n * n / (op_ms * 1e-3) / 1e9

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
        if world > 1:
            t = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e2e_s = float(t.item())
        assert float(out_np[-1, -1]) == float(acc[-1, -1].item())
        e2e = {"value": cells_total / e2e_s / 1e6, "unit": "Mcells/s", "h2d_bytes_per_step": 2 * n * n,
               "d2h_bytes_per_step": 8 * n * n, "ms_per_step": 1e3 * e2e_s,
               "api": "wmf_heavy_b200.cu_py.basics.basin_acum(flowdir int8, mask bool) -> float64"}
        del h_dir, h_mask, h_out

    if args.e2e_steps > 0 and world > 1:
        from wmf_heavy_b200 import stripes
        h_dir = torch.empty((n, n), dtype=torch.int8).pin_memory()
        h_dir.copy_(d)
        h_out = torch.empty((n, n), dtype=acc.dtype).pin_memory()

        def e2e_step():
            dd = h_dir.to(dev, non_blocking=True)
            a = stripes.accumulate_striped(dd, rank * n, world * n, acc_dtype=stripe_dtype)
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
        e2e = {"value": cells_total / e2e_s / 1e6, "unit": "Mcells/s", "h2d_bytes_per_step": n * n,
               "d2h_bytes_per_step": acc.element_size() * n * n, "ms_per_step": 1e3 * e2e_s,
               "api": "wmf_heavy_b200.stripes.accumulate_striped(host int8 stripe) -> host int64 stripe, per rank"}
        del h_dir, h_out

    shia = None
    if not args.no_shia:
        try:
            shia = bench_shia(torch, ops, dev, args.shia_cells, args.shia_steps, 3, 1)
            if world > 1:
                t = torch.tensor([shia["ms_per_run"]], device=dev, dtype=torch.float64)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                shia["ms_per_run"] = float(t.item())
                shia["value"] = world * shia["cells"] * shia["intervals"] / (shia["ms_per_run"] * 1e-3)
                shia["cells"] *= world
        except Exception as e:           # the primary metric must still print
            shia = {"error": repr(e)}

    if rank == 0:
        cpu = None
        if not args.no_cpu_baseline and world == 1:
            cs = args.cpu_sample or n                    # the whole raster by default: ~10-30 s of one host core
            mc, times, _ = cpu_oracle_accum(cs, 1)
            cpu = {"value": mc, "unit": "Mcells/s", "cores": 1, "kind": "port",
                   "sample": (f"the whole {n}x{n} raster" if cs == n else f"{cs}x{cs} top-left window of the same raster")
                             + ", C restatement "
                             "of wmf_py basin_acum (oracle/wmf_oracle.c, gcc -O2, 1 thread); "
                             f"{times[0]:.2f} s"}
        # kernels of libwmfb200.so per accumulation (profiles/*launches*.csv): walk = indegree + walk; sweep = sweepA,
        # generalA, g_build, g_walk, g_unres, sweepC, generalC; stripes = 8 (local: + macro_links, stripe_paths,
        # stripe_summary) + 6 (boundary forest) + 3 (route, sweepC, generalC; + widen for an int64 accumulation)
        launches_per_step = ({"walk": 2, "sweep": 7, "auto": 7}[args.algo] if world == 1
                             else (17 if world * n * n <= 0xFFFFFFFF else 18))   # (wider rasters also widen the inflows)
        line = {
            "metric": METRIC, "value": value, "unit": "Mcells/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": str(acc.dtype).replace("torch.", ""), "data": "synthetic",
            "config": {"workload": f"D8 flow accumulation, synthetic Scheidegger DIR {n}x{n} per GPU (seed {SEED}), "
                                   f"int8 DIR -> {str(acc.dtype).replace('torch.', '')} ACC, mask = all valid", "algo": args.algo,
                       "parallelism": "1 GPU" if world == 1 else
                       f"{world} row stripes of {n} rows of one {world * n}x{n} raster, one NCCL all-gather of the "
                       "boundary records per accumulation",
                       "l2": "inputs (DIR 0.27 GB + ACC 1.07 GB) exceed the 126 MB L2; no explicit flush"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         # dram__bytes_read+write of all kernels of one accumulation (ncu, profiles/r01_summary.md), 16384^2 int32
                         "traffic": 1.997e9 if (world == 1 and n == 16384 and args.algo != "walk") else None,
                         "peak_source": peak_src,
                         "kernel": "wmf_d8_accum, all passes of one accumulation "
                                   f"({bytes_per_cell:.0f} B/cell algorithmic)"},
            "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches_per_step * args.steps,
            "clocks": clk.summary(), "shia": shia,
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
    ap.add_argument("--ref-sample", type=int, default=8192, help="window edge of one step of --impl reference")
    ap.add_argument("--cpu-sample", type=int, default=0, help="window edge for cpu_baseline (0 = the whole raster)")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--shia-cells", type=int, default=1 << 20)
    ap.add_argument("--shia-steps", type=int, default=1000)
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
