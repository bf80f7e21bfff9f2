"""Counts the SASS mnemonics that prove a Blackwell-native kernel (B200_PROFILING.md: tcgen05.mma -> UTC*MMA,
tcgen05.ld -> LDTM, TMA -> UTMALDG / UTMASTG, mma.sync -> HMMA) per kernel of the built objects.

    python profiles/sass_summary.py > profiles/r02_sass_summary.txt
"""
import collections
import glob
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = ("UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "UTCBAR", "SYNCS", "HMMA", "LDGSTS")

rows = []
for obj in sorted(glob.glob(os.path.join(ROOT, "gemmgan_b200", "build", "*.o"))):
    sass = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    kernel, counts, total = None, collections.Counter(), 0
    def flush():
        if kernel and (total > 0):
            rows.append((os.path.basename(obj), kernel, total, dict(counts)))
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            flush()
            kernel = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            kernel = re.sub(r"\(.*", "", kernel)
            counts, total = collections.Counter(), 0
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m:
            total += 1
            op = m.group(1)
            for k in KEYS:
                if op.startswith(k):
                    counts[k] += 1
    flush()
print(f"{'object':18s} {'kernel':64s} {'instrs':>7s}  " + "  ".join(f"{k:>7s}" for k in KEYS))
for obj, kernel, total, c in rows:
    if not any(c.get(k) for k in KEYS[:10]):
        continue
    print(f"{obj:18s} {kernel[:64]:64s} {total:7d}  " + "  ".join(f"{c.get(k, 0):7d}" for k in KEYS))
