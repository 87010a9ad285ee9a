"""Per-kernel SASS instruction counts of libflowstate_b200.so (cuobjdump -sass): the mnemonics that prove the
Blackwell paths (UTCHMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UBLKCP = cp.async.bulk, SYNCS = mbarrier,
FFMA2 / FADD2 / FMUL2 = packed FP32, MUFU, VIMNMX3, HFMA2 ...).   python scripts/sass_summary.py > profiles/r02_sass_summary.md"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "flowstate_b200", "libflowstate_b200.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip() or n
KEYS = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UBLKCP", "SYNCS", "FFMA2", "FADD2", "FMUL2", "FFMA", "MUFU", "VIMNMX3",
        "SHFL", "VOTE", "REDUX", "LDS", "STS", "LDG", "STG", "DFMA", "DADD", "BAR", "ELECT"]
kernels = collections.OrderedDict()
cur = None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = demangle(m.group(1))
        cur = re.sub(r"\(.*", "", cur).replace("void ", "").replace("fs::", "")
        kernels[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,5}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        op = m.group(1)
        kernels[cur]["total"] += 1
        base = op.split(".")[0]
        for k in KEYS:
            if base == k or (k in ("FFMA2", "FADD2", "FMUL2") and base == k):
                kernels[cur][k] += 1
print("# SASS summary of `flowstate_b200/libflowstate_b200.so` (round 2)\n")
print("`python scripts/sass_summary.py` = `cuobjdump -sass` of the in-tree library, static instruction counts per kernel "
      "(sm_100a).  UTCHMMA = `tcgen05.mma`, UTCBAR = `tcgen05.commit`, LDTM / STTM = `tcgen05.ld / st`, UBLKCP = "
      "`cp.async.bulk` (TMA), SYNCS = mbarrier operations, FFMA2 / FADD2 / FMUL2 = packed FP32 pairs.\n")
cols = ["total"] + KEYS
print("| kernel | " + " | ".join(cols) + " |")
print("|---|" + "---|" * len(cols))
for name, c in sorted(kernels.items(), key=lambda kv: -kv[1]["total"]):
    print("| `%s` | " % name[:70] + " | ".join(str(c[k]) if c[k] else "" for k in cols) + " |")
tot = collections.Counter()
for c in kernels.values():
    tot.update(c)
print("\nLibrary totals: " + ", ".join("%s %d" % (k, tot[k]) for k in cols if tot[k]))
