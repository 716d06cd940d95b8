#!/usr/bin/env python3
"""Per-kernel launch counts, mean duration and share of the extraction kernels from an ncu launch list
(`ncu --metrics gpu__time_duration.sum --csv`), next to bench.py's live stage times of the same build.
usage: launch_shares.py launch_list.csv bench.json > profiles/rXX_launch_shares.txt"""
import csv, json, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = next(r for r in rows if r and r[0] == "ID")
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
acc = collections.OrderedDict()
for r in rows:
    if r and r[0].isdigit():
        name = r[ki].split("(")[0].strip()
        a = acc.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += float(r[vi].replace(",", "")) / 1e3
ours = {k: v for k, v in acc.items() if "orbb::" in k or k.startswith("k_") or "k_" in k.split("::")[-1]}
ext = {k: v for k, v in ours.items() if any(s in k for s in ("k_level0", "k_resize", "k_fast_cells", "k_octree", "k_blur", "k_angle_orb"))}
tot = sum(v[1] for v in ext.values())
print("# ncu --metrics gpu__time_duration.sum --clock-control none -c 400 python bench.py --steps 2 --warmup 1 --no-cpu --no-rgbd --no-refgpu --no-cfg5 --no-configs --no-parity --sustain-s 0")
print("# (cold-cache, serialised launches: compare SHARES with bench.py's live stages_ms, not absolutes)")
if len(sys.argv) > 2:
    d = next(json.loads(l) for l in open(sys.argv[2]) if l.startswith("{"))
    st = d["stages_ms"]; s = sum(st.values())
    print("# live stages_ms of the same build (bench.py, 256 frames): " + ", ".join(f"{k} {v:.3f} ({100 * v / s:.1f}%)" for k, v in st.items()))
print("kernel,launches,mean_us,share_of_extraction_kernels")
for k, v in sorted(ext.items(), key=lambda kv: -kv[1][1]):
    print(f"{k},{v[0]},{v[1] / v[0]:.1f},{100 * v[1] / tot:.1f}%")
other = [k for k in ours if k not in ext]
print("# other repo kernels in the same command: " + ", ".join(other))
