#!/usr/bin/env python3
"""Share of k_fast_cells' executed warp instructions per phase of the kernel, from a tools/ncu_sass_dump.py dump.
Phase boundaries are found from marker statements in k_fast.cu, so they follow the source as it changes; instructions
inlined from CUDA headers are attributed to the phase of the last k_fast.cu line before them (address order).
usage: fast_phases.py dump.txt [k_fast.cu]"""
import re, sys, os
dump = sys.argv[1]
srcp = sys.argv[2] if len(sys.argv) > 2 else os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "jetracer-orbslam2_b200", "csrc", "k_fast.cu")
src = open(srcp).read().splitlines()
def line_of(pat, nth=0):
    hits = [i + 1 for i, l in enumerate(src) if pat in l]
    return hits[nth]
marks = [(1, "arc_score fn" ), (line_of("k_fast_cells(const CUtensorMap"), "prologue"), (line_of("const int gx = c.x0 - 4"), "staging"),
         (line_of("score_bytes >> 4", 1), "score clear + setup"), (line_of("auto precheck"), "precheck"), (line_of("const int c_own"), "scan"),
         (line_of("const int xoff = 4 * U - 64"), "queue fill"), (line_of("int cn = 0;"), "arc loop"), (line_of("if (DUMP) {"), "dump"),
         (line_of("        kn = 0;"), "nms"), (line_of("int *counter"), "emit")]
marks.sort()
def phase(l):
    name = marks[0][1]
    for ln, nm in marks:
        if l >= ln: name = nm
    return "arc loop" if name == "arc_score fn" and l > 30 else name
agg, last, tot, cells = {}, "prologue", 0, None
for ln in open(dump):
    if ln.startswith("#"): continue
    p = ln.split(); f, l = p[1].rsplit(":", 1); cnt = int(p[2])
    if cells is None and cnt: cells = cnt
    if f == "k_fast.cu": last = phase(int(l))
    agg[last] = agg.get(last, 0) + cnt; tot += cnt
print(f"warp-instr {tot}, {tot / cells:.0f} per cell-warp")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1]): print(f"{k:22s} {100 * v / tot:5.1f}%  {v / cells:7.1f} warp-instr/cell")
