#!/usr/bin/env python3
"""The result table BASELINE.md section 4 asks for, as markdown, from the bench lines committed under profiles/:
one row per config x {CPU 1 core, CPU all cores, B200 x 1 / 2 / 4 / 8}.
usage: python tools/results_table.py [final_scaling.jsonl] [earlier_scaling.jsonl] > table.md"""
import json, os, sys

root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
final = sys.argv[1] if len(sys.argv) > 1 else os.path.join(root, "profiles", "r02k_scaling.jsonl")
earlier = sys.argv[2] if len(sys.argv) > 2 else os.path.join(root, "profiles", "r02e_scaling.jsonl")


def lines(path):
    return [json.loads(l) for l in open(path) if l.startswith('{"metric"')]


def row(cfg, dev, n, fps, hbm, gp, popc, parity, note=""):
    f = lambda v, fmt: "-" if v is None else format(v, fmt)
    ms = None if not fps else 1e3 / fps
    print(f"| {cfg} | {dev} | {n} | {f(fps, ',.0f')} | {f(ms, '.4f')} | {f(hbm, '.3f')} | {f(gp, ',.1f')} | {f(popc, '.2f')} | {parity} | {note} |")


print("| config | device | cores / GPUs | frames/s | ms/frame | B·fps / HBM peak | matcher Gpairs/s | per GPU ÷ POPC roof | parity (exact / tolerated / failed) | source |")
print("|---|---|---|---|---|---|---|---|---|---|")
fin = {d["n_gpus"]: d for d in lines(final)}
old = {d["n_gpus"]: d for d in lines(earlier)}
d1 = fin[1]
cb = d1["cpu_baseline"]
name = "cfg 1/5 geometry: 640×480, 1000 kp, 8 levels"
row(name, "CPU oracle port", 1, cb["one_core_value"], None, cb["matcher_gpairs_one_core"], None, "reference for the parity column", os.path.basename(final))
row(name, "CPU oracle port", cb["cores"], cb["value"], None, cb["matcher_gpairs"], None, "", os.path.basename(final))
for n in (1, 2, 4, 8):
    d, src = (fin[n], final) if n in fin else (old.get(n), earlier)
    if d is None:
        continue
    p = d.get("parity", {})
    par = f"{p.get('exact', '-')} / {p.get('tolerated', '-')} / {p.get('failed', '-')} keypoints of {p.get('frames', '-')} frames" if p else "-"
    m = d["matcher"]
    kind = {0: "POPC", 1: "mma.sync", 2: "tcgen05"}.get(m.get("kernel_kind", 1), "mma.sync")
    row(name, "B200", n, d["value"], d["step_roofline"]["frac_of_hbm"], m["value"], m["vs_popc_roof"],
        par, f"{os.path.basename(src)} ({kind} matcher; e2e {d['e2e']['value']:,.0f} frames/s; cfg 5 pipeline {d['cfg5']['ms']['total']:.2f} ms, hash {d['cfg5']['records_sha256'][:8]})")
for key, label in (("cfg2_848x480_1200kp_batch64", "cfg 2: 848×480, 1200 kp, batch 64"),
                   ("cfg3_848x800_1000kp_stereo16_match", "cfg 3: 848×800 stereo ×16 + 2-NN matching"),
                   ("cfg4_1280x720_2000kp_batch256", "cfg 4: 1280×720, 2000 kp, batch 256")):
    c = d1.get("other_configs", {}).get(key)
    if c:
        row(label, "B200", 1, c["frames_per_s"], c["frac_of_hbm"], None, None, f"{c['parity_frames_exact']} sampled frames bit-exact", os.path.basename(final))
