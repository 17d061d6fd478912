"""SASS census of libwmfb200.so: python tools/sass_census.py > profiles/r02_sass_census.md
Counts, per kernel family, the instructions that show how data moves (cp.async = LDGSTS, TMA bulk copy = UBLKCP,
256-bit stores, shared / global atomics, mbarrier waits)."""
import collections
import os
import re
import subprocess

lib = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "wmf_heavy_b200", "libwmfb200.so")
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
pats = collections.OrderedDict([("LDGSTS (cp.async)", r"\bLDGSTS"), ("UBLKCP (cp.async.bulk, TMA)", r"\bUBLKCP"),
                                ("SYNCS (mbarrier)", r"\bSYNCS"), ("STG.*256", r"\bSTG\.E\.\S*256"), ("LDG.*128", r"\bLDG\.E\.\S*128"),
                                ("ATOMS / RED shared", r"\bATOMS"), ("ATOMG / RED global", r"\b(ATOMG|RED)\b"), ("SHFL", r"\bSHFL"),
                                ("MUFU", r"\bMUFU"), ("UTMALDG", r"\bUTMALDG"), ("HMMA / UTC*MMA", r"\b(HMMA|UTC\wMMA)")])
fam = collections.OrderedDict()
cur = None
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", "-p", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = name.replace("(anonymous namespace)::", "").split("<")[0].split("(")[0]
        cur = fam.setdefault(name, collections.Counter())
        cur["variants"] += 1
        continue
    if cur is None:
        continue
    if re.search(r"/\*[0-9a-f]{4,}\*/\s+\S", line):
        cur["instructions"] += 1
        for k, p in pats.items():
            if re.search(p, line):
                cur[k] += 1
cols = ["variants", "instructions"] + list(pats)
print("# SASS census of `libwmfb200.so` (sm_100a), all template variants of a kernel summed\n")
print("Built by `tools/sass_census.py` from `cuobjdump -sass`.  No tensor-core instruction appears anywhere: the path is integer / fp32")
print("elementwise work (SURVEY 8(d)).  `UBLKCP` + `SYNCS` are the TMA bulk-copy staging variants of the fast sweep kernels")
print("(opt-in, `WMF_D8_BULK=1`: measured 1 % slower than the `LDGSTS` variants, DESIGN.md 4.1).\n")
print("| kernel | " + " | ".join(cols) + " |")
print("|---|" + "---|" * len(cols))
tot = collections.Counter()
for name, c in fam.items():
    print(f"| `{name}` | " + " | ".join(str(c.get(k, 0)) for k in cols) + " |")
    tot.update(c)
print("| **total** | " + " | ".join(str(tot.get(k, 0)) for k in cols) + " |")
