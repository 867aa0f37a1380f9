#!/usr/bin/env python
"""Opcode evidence from the built library (no GPU needed):  python profiles/sass_summary.py > profiles/r2_sass_summary.txt
Counts, per kernel of libqtb200.so, the SASS mnemonics that prove the tcgen05 / TMA / TMEM data path (B200_PROFILING.md):
UTCxMMA (tcgen05.mma by kind), UTMALDG / UTMASTG (TMA loads, tiled and im2col, and stores), LDTM / STTM (TMEM access),
UTCBAR (tcgen05.commit), SYNCS (mbarrier), ELECT (elect.sync issue), plus register / spill figures from cuobjdump -res-usage."""
import collections
import os
import re
import subprocess
import sys

LIB = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "pytorch_quantize_impls_b200", "csrc", "libqtb200.so")
PAT = re.compile(r"\b(UTC[A-Z]*MMA(?:\.2CTA)?|UTMALDG(?:\.[0-9]D)?(?:\.IM2COL)?(?:\.2CTA)?|UTMASTG|LDTM|STTM|UTCBAR(?:\.2CTA)?|SYNCS|ELECT|POPC|IMMA|HMMA|STL|LDL)\b")
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
kern, counts, size = None, {}, {}
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        counts[kern] = collections.Counter()
        size[kern] = 0
        continue
    if kern and re.search(r"/\*[0-9a-f]{4,}\*/\s+\S", line):
        size[kern] += 1
        for op in PAT.findall(line):
            counts[kern][op] += 1
res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True).stdout
regs = {}
cur = None
for line in res.splitlines():
    m = re.search(r"Function (\S+):", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
    m = re.search(r"REG:(\d+).*?STACK:(\d+)", line)
    if m and cur:
        regs[cur] = (int(m.group(1)), int(m.group(2)))
fam = collections.defaultdict(list)
for k in counts:
    fam[re.sub(r"<.*", "", k)].append(k)
print("libqtb200.so: %d kernels, %d SASS instructions" % (len(counts), sum(size.values())))
for f in sorted(fam):
    ks = fam[f]
    tot = collections.Counter()
    for k in ks:
        tot.update(counts[k])
    r = [regs[k] for k in ks if k in regs]
    print("\n%s  (%d instantiations, %d instructions; registers %s, stack bytes max %s)" % (
        f, len(ks), sum(size[k] for k in ks), ("%d-%d" % (min(a for a, _ in r), max(a for a, _ in r))) if r else "?",
        max(b for _, b in r) if r else "?"))
    print("   " + ", ".join("%s %d" % kv for kv in sorted(tot.items())) if tot else "   (no tensor / TMA / TMEM opcodes: CUDA-core kernel)")
    if len(ks) <= 64 and any(counts[k] for k in ks) and "--per-kernel" in sys.argv:
        for k in sorted(ks):
            print("     %s: %s" % (k[len(f):], dict(counts[k])))
