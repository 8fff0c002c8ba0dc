#!/usr/bin/env python3
"""Per-kernel counts of the SASS mnemonics that prove the Blackwell paths (cuobjdump -sass of the in-tree library):
UTCHMMA (tcgen05.mma), LDTM (tcgen05.ld), UTCBAR (tcgen05.commit), UTMALDG (cp.async.bulk.tensor), LDGSTS (cp.async),
SYNCS (mbarrier), ACQBULK / UTMACCTL (tensormap proxy fence).   python profiles/sass_summary.py > profiles/r02_sass_summary.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "mfas_b200", "_mfas_b200.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
cur, counts = None, collections.OrderedDict()
keys = ["UTCHMMA", "LDTM", "UTCBAR", "UTMALDG", "UTMACCTL", "LDGSTS", "SYNCS", "UTCATOMSWS"]
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        counts[cur] = collections.Counter()
        continue
    if cur:
        for k in keys:
            if re.search(r"\b" + k + r"\b|\b" + k + r"\.", line):
                counts[cur][k] += 1
print("kernel".ljust(70), " ".join(k.rjust(9) for k in keys))
for fn, c in counts.items():
    if sum(c.values()):
        print(fn[:70].ljust(70), " ".join(str(c[k]).rjust(9) for k in keys))
tot = collections.Counter()
for c in counts.values():
    tot.update(c)
print("TOTAL".ljust(70), " ".join(str(tot[k]).rjust(9) for k in keys))
