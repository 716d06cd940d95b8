#!/usr/bin/env python3
"""Per-source-line LSU work for one kernel of an .ncu-rep: shared-memory wavefronts (ideal / excessive), global tag
requests and instruction counts, grouped by source line (same SASS <-> line join as ncu_lines.py) and by opcode.
usage: ncu_wavefronts.py report.ncu-rep lib.so kernel_substring [top_n]"""
import collections, csv, io, os, re, subprocess, sys, tempfile

rep, so, kname = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
blocks, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "ins": []}
        blocks.append(cur)
    elif cur is not None and r and r[0] == "Address":
        cur["hdr"] = r
    elif cur is not None and cur["hdr"] and r and r[0].startswith("0x"):
        cur["ins"].append(r)
blk = next(b for b in blocks if kname.split("<")[0] in b["name"])
h = blk["hdr"]
col = {n: h.index(n) for n in ("Source", "Instructions Executed", "L1 Wavefronts Shared", "L1 Wavefronts Shared Ideal",
                               "L1 Wavefronts Shared Excessive", "L1 Tag Requests Global", "# Samples")}
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, capture_output=True)
lines = None
for f in sorted(os.listdir(tmp)):
    txt = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(tmp, f)], capture_output=True, text=True).stdout
    for sec in re.split(r"\n//-+ \.text\.", txt)[1:]:
        lst, cur_line, cur_file = [], 0, ""
        for ln in sec.splitlines():
            m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
            if m:
                if "inlined at" not in ln or cur_line == 0:
                    cur_line, cur_file = int(m.group(2)), os.path.basename(m.group(1))
            elif re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+(@!?U?P\w+\s+)?[A-Z]", ln):
                lst.append((cur_file, cur_line))
        if len(lst) == len(blk["ins"]):
            dem = subprocess.run(["c++filt", sec.split()[0]], capture_output=True, text=True).stdout
            want = [t for t in re.findall(r"\d+", kname.split("<", 1)[1] if "<" in kname else "") if len(t) > 1]
            if kname.split("<")[0] in dem and all(t in dem for t in want):
                lines = lst
if lines is None:
    sys.exit("could not align SASS with line info")
num = lambda s: int(float(s)) if s not in ("", "-") else 0
per_line = collections.defaultdict(lambda: [0, 0, 0, 0, 0])
per_op = collections.defaultdict(lambda: [0, 0, 0, 0, 0])
tot = [0, 0, 0, 0, 0]
for (f, l), r in zip(lines, blk["ins"]):
    op = re.sub(r"^@!?U?P\w+\s+", "", r[col["Source"]].strip()).split()[0]
    v = [num(r[col["Instructions Executed"]]), num(r[col["L1 Wavefronts Shared"]]), num(r[col["L1 Wavefronts Shared Ideal"]]),
         num(r[col["L1 Wavefronts Shared Excessive"]]), num(r[col["L1 Tag Requests Global"]])]
    for a in (per_line[(f, l)], per_op[op], tot):
        for i in range(5):
            a[i] += v[i]
print(f"{blk['name'][:70]}: warp-instr {tot[0]}, shared wavefronts {tot[1]} (ideal {tot[2]}, excessive {tot[3]}), global tag requests {tot[4]}")
src = {}
print("-- by source line (sorted by shared wavefronts + global tag requests)")
for (f, l), a in sorted(per_line.items(), key=lambda kv: -(kv[1][1] + kv[1][4]))[:top]:
    if f not in src:
        p = os.path.join(os.path.dirname(os.path.abspath(so)), "csrc", f)
        src[f] = open(p).read().splitlines() if os.path.exists(p) else []
    text = src[f][l - 1].strip()[:70] if 0 < l <= len(src[f]) else ""
    print(f"{f}:{l:4d} inst {100*a[0]/tot[0]:5.1f}%  smem wf {100*a[1]/max(tot[1],1):5.1f}% (x{a[1]/max(a[2],1):.2f} of ideal)  gmem req {100*a[4]/max(tot[4],1):5.1f}%  {text}")
print("-- by opcode")
for op, a in sorted(per_op.items(), key=lambda kv: -kv[1][0])[:28]:
    print(f"{op:24s} inst {100*a[0]/tot[0]:5.1f}%  smem wf {100*a[1]/max(tot[1],1):5.1f}%  gmem req {100*a[4]/max(tot[4],1):5.1f}%")
