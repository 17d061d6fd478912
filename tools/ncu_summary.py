"""Summarise an .ncu-rep (ncu --set full) per kernel into Markdown: python tools/ncu_summary.py prof.ncu-rep"""
import csv
import io
import subprocess
import sys

raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
want = [("gpu__time_duration.sum", "duration"), ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput % of peak"),
        ("sm__inst_executed.avg.per_cycle_elapsed", "IPC (per SM)"), ("smsp__inst_executed.sum", "warp instructions"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active % of peak"),
        ("launch__registers_per_thread", "registers/thread"),
        ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe %"),
        ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe %"),
        ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe %"),
        ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU pipe %")]
for r in rows[2:]:
    name = r[idx["Kernel Name"]].split("(")[0].replace("void <unnamed>::", "").replace("<unnamed>::", "")
    print(f"### {name}\n")
    print("| metric | value |\n|---|---|")
    for k, label in want:
        if k in idx and r[idx[k]] not in ("", "n/a"):
            print(f"| {label} | {r[idx[k]]} {units[idx[k]]} |")
    st = [(float(r[idx[h]]), h) for h in hdr if "smsp__average_warps_issue_stalled" in h and h.endswith("per_issue_active.ratio")
          and "not_issued" not in h and r[idx[h]] not in ("", "n/a")]
    top = ", ".join(f"{h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')} {v:.2f}"
                    for v, h in sorted(st, reverse=True)[:5])
    print(f"| top stalls (warps per issue) | {top} |\n")
