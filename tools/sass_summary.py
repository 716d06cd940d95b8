#!/usr/bin/env python3
"""Static SASS evidence for profiles/: per kernel of liborbb200.so, the number of instructions and the counts of the
mnemonics the design relies on (cuobjdump -sass; no GPU needed).
usage: python tools/sass_summary.py [lib.so] > profiles/rXX_sass_mnemonics.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "jetracer-orbslam2_b200", "liborbb200.so")
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
WATCH = ["POPC", "LOP3", "VABSDIFF4", "VIMNMX3", "VIMNMX", "IDP.4A", "IDP.2A", "PRMT", "SHF", "IMAD.HI", "REDUX", "MATCH", "VOTE",
         "SHFL", "LDG.E.128", "LDG.E.64", "LDG.E.U8", "LDS.128", "STS.128", "ATOMS", "ATOMG", "RED", "UTMALDG", "SYNCS",
         "ACQBULK", "BAR", "HMMA", "UTCHMMA", "DMUL", "DADD", "DFMA", "FFMA", "F2I"]
cur, arch = None, None
stats = collections.OrderedDict()
for line in txt.splitlines():
    m = re.search(r"arch = (sm_\w+)", line)
    if m:
        arch = m.group(1)
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        cur = stats.setdefault(name.replace("void ", ""), collections.Counter())
        cur["arch:" + (arch or "?")] += 0
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur is not None:
        op = m.group(1)
        cur["_total"] += 1
        for w in WATCH:
            if op == w or op.startswith(w + "."):
                cur[w] += 1
print(f"# cuobjdump -sass {os.path.basename(so)}: instruction counts per kernel (static), selected mnemonics")
print("# tensor-core mnemonics (HMMA / UTCHMMA) are expected to be 0 everywhere: nothing on this path is a dense contraction")
for name, c in stats.items():
    arch = [k for k in c if k.startswith("arch:")]
    parts = [f"{w}={c[w]}" for w in WATCH if c[w]]
    print(f"{name}  [{arch[0][5:] if arch else '?'}]  total={c['_total']}  " + "  ".join(parts))
