#!/usr/bin/env python3
"""Per-SASS-instruction dump of one kernel of an .ncu-rep joined with -lineinfo source lines:
address, file:line, executed warp instructions, share, stall samples, instruction text (in address order).
usage: ncu_sass_dump.py report.ncu-rep lib.so kernel_substring > out.txt"""
import csv, io, re, subprocess, sys, tempfile, os

rep, so, kname = sys.argv[1:4]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
blocks, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "ins": []}; blocks.append(cur)
    elif cur is not None and r and r[0] == "Address":
        cur["hdr"] = r
    elif cur is not None and cur["hdr"] and r and r[0].startswith("0x"):
        cur["ins"].append(r)
blk = next(b for b in blocks if kname in b["name"])
h = blk["hdr"]
ci, cs = h.index("Instructions Executed"), h.index("# Samples")
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, capture_output=True)
lines = None
for f in sorted(os.listdir(tmp)):
    txt = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(tmp, f)], capture_output=True, text=True).stdout
    for sec in re.split(r"\n//-+ \.text\.", txt)[1:]:
        dem = subprocess.run(["c++filt", sec.split()[0]], capture_output=True, text=True).stdout
        if kname.split("<")[0] not in dem:
            continue
        cur_line, cur_file, lst = 0, "", []
        for ln in sec.splitlines():
            m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
            if m:
                if "inlined at" not in ln or cur_line == 0:
                    cur_line = int(m.group(2)); cur_file = os.path.basename(m.group(1))
            elif re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+(@!?U?P\w+\s+)?[A-Z]", ln):
                lst.append((cur_file, cur_line, ln.split("*/", 1)[1].split("/*")[0].strip()))
        if len(lst) == len(blk["ins"]):
            lines = lst
if lines is None:
    sys.exit("could not align SASS with line info")
tot = sum(int(r[ci]) for r in blk["ins"])
print(f"# {blk['name'][:100]}: warp-instr {tot}")
for k, ((f, l, text), r) in enumerate(zip(lines, blk["ins"])):
    print(f"{k:5d} {f}:{l:<4d} {int(r[ci]):>10d} {100*int(r[ci])/tot:5.2f}% s{int(r[cs]):<5d} {text}")
