#!/usr/bin/env python
"""Per-kernel counts of the SASS mnemonics that prove which hardware path a kernel uses (tcgen05: UTCHMMA / UTCBAR / LDTM;
TMA: UTMALDG / UBLKCP; legacy tensor core: HMMA; MUFU; clock reads) in the in-tree library, for profiles/.
   python tools/sass_summary.py [path/to/lib.so] > profiles/r2_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "xr_image_segmentation_b200", "libxrseg.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
pats = {"UTCHMMA": r"\bUTCHMMA", "UTCBAR": r"\bUTCBAR", "LDTM": r"\bLDTM", "UTMALDG": r"\bUTMALDG", "UBLKCP": r"\bUBLKCP",
        "UTMASTG": r"\bUTMASTG", "HMMA": r"\bHMMA", "LDSM": r"\bLDSM", "MUFU.TANH": r"MUFU\.TANH", "MUFU.EX2": r"MUFU\.EX2",
        "CLOCK": r"SR_CLOCK|CS2R\S* R\d+, SR_CLOCK", "ST.256": r"STG\.E\.ENL2\.256|STG\.E\.256", "SYNCS": r"\bSYNCS"}
cur = None
counts = collections.OrderedDict()
arch = set()
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"\(.*", "", cur).replace("xrseg::", "").replace("(anonymous namespace)::", "")
        counts[cur] = collections.Counter()
        continue
    m = re.search(r"arch = (sm_\w+)", line)
    if m:
        arch.add(m.group(1))
    if cur is None:
        continue
    counts[cur]["instr"] += 1 if re.search(r"/\*[0-9a-f]{4}\*/", line) else 0
    for k, p in pats.items():
        if re.search(p, line):
            counts[cur][k] += 1
print(f"# SASS summary of {os.path.relpath(lib, ROOT)} (cuobjdump -sass; arch {', '.join(sorted(arch))})")
keys = list(pats)
print("kernel".ljust(58) + " instr " + " ".join(k.rjust(9) for k in keys))
for k, c in counts.items():
    print(k[:57].ljust(58) + f"{c['instr']:6d} " + " ".join((str(c[x]) if c[x] else ".").rjust(9) for x in keys))
