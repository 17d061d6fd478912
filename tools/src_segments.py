"""Group the SASS of one kernel in an `ncu --page source --csv` dump into runs of equal execution count.

python tools/src_segments.py src.csv [kernel-substring] [min Minstr]
"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
want = sys.argv[2] if len(sys.argv) > 2 else ""
floor = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
for si, s in enumerate(starts):
    name = rows[s][1]
    if want not in name:
        continue
    end = starts[si + 1] if si + 1 < len(starts) else len(rows)
    hdr = rows[s + 1]
    ix = {h: i for i, h in enumerate(hdr)}
    if "Instructions Executed" not in ix:
        continue
    data = [r for r in rows[s + 2:end] if len(r) == len(hdr)]
    base = int(data[0][ix["Address"]], 16)
    tot = sum(int(r[ix["Instructions Executed"]]) for r in data)
    samp = sum(int(r[ix["# Samples"]]) for r in data)
    print(f"== {name[:60]}: {tot/1e6:.1f} M warp instructions, {samp} samples")
    segs, cur = [], None
    for r in data:
        a = int(r[ix["Address"]], 16) - base
        n = int(r[ix["Instructions Executed"]])
        sm = int(r[ix["# Samples"]])
        if cur is None or abs(n - cur["n"]) > 0.05 * max(n, cur["n"]) + 1000:
            cur = {"a0": a, "n": n, "cnt": 0, "inst": 0, "samp": 0}
            segs.append(cur)
        cur["a1"] = a
        cur["cnt"] += 1
        cur["inst"] += n
        cur["samp"] += sm
    for g in segs:
        if g["inst"] > floor * 1e6:
            print(f"{g['a0']:6x}-{g['a1']:6x} static={g['cnt']:4d} exec={g['n']:9d} inst={g['inst']/1e6:7.1f}M "
                  f"({100*g['inst']/tot:4.1f}%) samples={g['samp']} ({100*g['samp']/samp:4.1f}%)")
    break
