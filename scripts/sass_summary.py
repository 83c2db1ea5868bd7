"""python scripts/sass_summary.py > profiles/rNN_sass_summary.txt

Per-kernel SASS mnemonic counts of the in-tree libbjx.so (`cuobjdump -sass`): the evidence that the hot kernels are
hand-written tcgen05 / TMEM / TMA code (UTCHMMA, LDTM/STTM, UTMALDG) and not a recompiled library."""
import collections
import re
import subprocess
import sys
from pathlib import Path

LIB = Path(__file__).resolve().parent.parent / "bijax_b200" / "_lib" / "libbjx.so"
MNEMONICS = ["UTCHMMA", "UTMALDG", "UBLKPF", "LDTM", "STTM", "UTCBAR", "SYNCS", "MUFU", "RED", "DFMA", "HMMA"]
pat = re.compile(r"\b(" + "|".join(MNEMONICS) + r")(?:\.|\b)")
sass = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True, check=True).stdout
cur, counts = None, collections.defaultdict(collections.Counter)
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        continue
    if cur and re.match(r"\s+/\*[0-9a-f]{4,}\*/", line):
        counts[cur]["instr"] += 1
        m = pat.search(line)
        if m:
            counts[cur][m.group(1)] += 1
names = subprocess.run(["c++filt"], input="\n".join(counts), capture_output=True, text=True).stdout.splitlines()
rows = sorted(((re.sub(r"\(.*", "", n), counts[k]) for n, k in zip(names, counts)), key=lambda r: (-r[1]["UTCHMMA"], -r[1]["instr"]))
print("# SASS mnemonic counts per kernel of bijax_b200/_lib/libbjx.so (sm_100a only), from `cuobjdump -sass` (scripts/sass_summary.py).")
print("# UTCHMMA = tcgen05.mma (kind::f16, bf16 operands) | UTMALDG = cp.async.bulk.tensor (TMA load) | UBLKPF = cp.async.bulk.prefetch.L2")
print("# LDTM / STTM = tcgen05.ld / tcgen05.st (TMEM) | UTCBAR = tcgen05.commit | SYNCS = mbarrier | MUFU = special-function unit")
print("# RED = fire-and-forget global reduction | DFMA = fp64 FMA | HMMA = legacy mma.sync (must be 0: nothing here is a recompiled wmma kernel)")
cols = ["instr"] + MNEMONICS
print(f"{'kernel':64s} " + " ".join(f"{c:>7s}" for c in cols))
tot = collections.Counter()
for name, c in rows:
    tot.update(c)
    print(f"{name[:64]:64s} " + " ".join(f"{c[x]:7d}" for x in cols))
print(f"{'TOTAL (' + str(len(rows)) + ' kernels)':64s} " + " ".join(f"{tot[x]:7d}" for x in cols))
