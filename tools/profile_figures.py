"""Figures only a profiler can give, for bench.py's JSON line -> profiles/r02_figures.json.

    python tools/profile_figures.py gpurun_out/r02_launches.csv [shia_launches.csv] > profiles/r02_figures.json

Input: the ncu launch list tools/prof_step.sh writes (gpu__time_duration, smsp__inst_executed, dram__bytes_read/write per
launch of `bench.py --steps 2 --warmup 3` at the default size).  Output keys:
  d8_traffic_bytes     dram__bytes_read.sum + dram__bytes_write.sum over ALL kernels of one accumulation
                       (total over the accumulation kernels / number of accumulations in the capture)
  d8_kernels           per kernel: launches per accumulation, average us, share of the step, DRAM bytes, warp instructions
  shia_<name>_issue_util  warp instructions / (duration * 148 SMs * 4 schedulers * SM clock), when a SHIA list is given
"""
import collections
import csv
import json
import sys


def read(path):
    lines = open(path).read().splitlines()
    i = [k for k, l in enumerate(lines) if l.startswith('"ID"')][0]
    per = collections.OrderedDict()
    for r in csv.DictReader(lines[i:]):
        name = r["Kernel Name"].split("(")[0].replace("void <unnamed>::", "").replace("<unnamed>::", "").replace("void ", "")
        per.setdefault((int(r["ID"]), name), {})[r["Metric Name"]] = float(r["Metric Value"].replace(",", ""))
    return per


def main():
    per = read(sys.argv[1])
    skip = ("scheid", "at::", "elementwise", "fill", "memset", "memcpy")
    agg = collections.OrderedDict()
    for (_, name), m in per.items():
        if any(s in name for s in skip):
            continue
        a = agg.setdefault(name.split("<")[0], {"n": 0, "ns": 0.0, "bytes": 0.0, "inst": 0.0})
        a["n"] += 1
        a["ns"] += m.get("gpu__time_duration.sum", 0.0)
        a["bytes"] += m.get("dram__bytes_read.sum", 0.0) + m.get("dram__bytes_write.sum", 0.0)
        a["inst"] += m.get("smsp__inst_executed.sum", 0.0)
    n_acc = max(a["n"] for k, a in agg.items() if k.startswith("k_sweepC")) if any(k.startswith("k_sweepC") for k in agg) else 1
    step_ns = sum(a["ns"] for a in agg.values()) / n_acc
    out = {"source": sys.argv[1], "accumulations_in_capture": n_acc,
           "d8_traffic_bytes": sum(a["bytes"] for a in agg.values()) / n_acc,
           "d8_step_us_under_ncu": step_ns / 1e3,
           "d8_kernels": {k: {"launches_per_accumulation": a["n"] / n_acc, "avg_us": a["ns"] / a["n"] / 1e3,
                              "share": a["ns"] / n_acc / step_ns, "dram_bytes": a["bytes"] / a["n"],
                              "warp_inst": a["inst"] / a["n"]} for k, a in agg.items()}}
    if len(sys.argv) > 2:
        sm_hz, slots = 1.965e9, 148 * 4
        for (_, name), m in read(sys.argv[2]).items():
            if name.startswith("k_shia5") and "gpu__time_duration.sum" in m:
                key = "shia_c3_type2_issue_util" if "2, 2, 2" in name or "T2" in name else "shia_c3_type1_issue_util"
                out.setdefault(key, m.get("smsp__inst_executed.sum", 0.0) / (m["gpu__time_duration.sum"] * 1e-9 * sm_hz * slots))
    json.dump(out, sys.stdout, indent=1)


if __name__ == "__main__":
    main()
