#!/usr/bin/env python3
"""Per-source-line instruction and stall-sample totals for one kernel of an .ncu-rep.

ncu's `--page source --csv` is per SASS instruction without line numbers; `nvdisasm --print-line-info` on the
cubin (built with -lineinfo) has the line of every instruction.  Both list a function's instructions in address
order, so they join by position.   usage: ncu_lines.py report.ncu-rep lib.so kernel_substring [top_n]
"""
import csv, io, re, subprocess, sys, tempfile, os, collections

rep, so, kname = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
# split per kernel block: a block starts with a row ["Kernel Name", name]
blocks, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "ins": []}
        blocks.append(cur)
    elif cur is not None and r and r[0] == "Address":
        cur["hdr"] = r
    elif cur is not None and cur["hdr"] and r and r[0].startswith("0x"):
        cur["ins"].append(r)
blk = next(b for b in blocks if kname in b["name"])
h = blk["hdr"]
ci, ct, cs = h.index("Instructions Executed"), h.index("Thread Instructions Executed"), h.index("# Samples")
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, capture_output=True)
lines = None
cand_lens = []
for f in sorted(os.listdir(tmp)):
    txt = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(tmp, f)], capture_output=True, text=True).stdout
    # find the .text section of the mangled kernel whose demangled name matches: use section headers
    secs = re.split(r"\n//-+ \.text\.", txt)
    for sec in secs[1:]:
        mangled = sec.split()[0]
        dem = subprocess.run(["c++filt", mangled], capture_output=True, text=True).stdout
        if kname.split("<")[0] in dem:
            cur_line, lst = 0, []
            for ln in sec.splitlines():
                m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
                if m:
                    if "inlined at" not in ln or cur_line == 0:
                        cur_line = int(m.group(2)); cur_file = os.path.basename(m.group(1))
                elif re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+(@!?U?P\w+\s+)?[A-Z]", ln):
                    lst.append((cur_file, cur_line, ln.split("*/", 1)[1].strip()[:60]))
            cand_lens.append(len(lst))
            if len(lst) == len(blk["ins"]):
                lines = lst
if lines is None:
    print("could not align SASS with line info", len(blk["ins"]), cand_lens); sys.exit(1)
agg = collections.defaultdict(lambda: [0, 0, 0])
for (f, l, _), r in zip(lines, blk["ins"]):
    a = agg[(f, l)]
    a[0] += int(r[ci]); a[1] += int(r[ct]); a[2] += int(r[cs])
tot_i = sum(a[0] for a in agg.values()); tot_s = sum(a[2] for a in agg.values())
print(f"{blk['name'][:80]}: warp-instr {tot_i}, samples {tot_s}")
src = {}
for (f, l), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    if f not in src:
        path = os.path.join(os.path.dirname(os.path.abspath(so)), "csrc", f)
        src[f] = open(path).read().splitlines() if os.path.exists(path) else []
    text = src[f][l - 1].strip()[:90] if 0 < l <= len(src[f]) else ""
    print(f"{f}:{l:4d} inst {100*a[0]/tot_i:5.1f}% samples {100*a[2]/max(tot_s,1):5.1f}%  {text}")
