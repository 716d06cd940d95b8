#!/usr/bin/env python3
"""tcgen05 matcher (k_match_umma; ORBB_MATCH_UMMA=2: every pair keyed in the epilogue) against the XOR / POPC matcher on the
same inputs, then its throughput.
The matcher's kernel choice is read once per process, so every form runs in a child process (bounded by a timeout: a
kernel whose MMA never completes traps after a bounded wait (~10 s) instead of hanging).  Shapes cover full and partial train tiles,
split-T, k = 1 and 2, and duplicate train rows (tie rule: lowest train index).
usage (GPU box): python tools/umma_probe.py            # parent: compares, then times the forms that matched
                 python tools/umma_probe.py child out.npz   # child: results of the form selected by the environment
UMMA_PROBE_NOTIME=1 skips the throughput part (tests/test_gpu_parity.py runs the comparison that way); exit code 1 if any
form failed to run or differs."""
import importlib, os, subprocess, sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHAPES = ((2000, 2000), (257, 50000), (5000, 333), (300, 1), (1027, 4099), (40000, 12345), (3000, 128), (1500, 262144))


def child(out):
    sys.path.insert(0, ROOT)
    import torch
    orbb = importlib.import_module("jetracer-orbslam2_b200.orbb")
    ex = orbb.ORBextractor(100, 1.2, 2, 20, 7, width=640, height=480, max_batch=1)
    st = torch.cuda.current_stream()
    res = {}
    rng = np.random.default_rng(7)
    for nq, nt in SHAPES:
        t = rng.integers(0, 256, (nt, 32), dtype=np.uint8)
        if nt > 8:
            t[nt // 2:nt // 2 + nt // 8] = t[:nt // 8]  # duplicates: ties on the exact distance
        q = rng.integers(0, 256, (nq, 32), dtype=np.uint8)
        if nt > 8:
            sel = rng.integers(0, nt, nq // 2)
            q[:nq // 2] = t[sel] ^ (rng.integers(0, 256, (nq // 2, 32), dtype=np.uint8) & rng.integers(0, 2, (nq // 2, 32), dtype=np.uint8))
        qd, td = torch.from_numpy(q).cuda(), torch.from_numpy(t).cuda()
        for k in (1, 2):
            idx = torch.full((nq, 2), -7, dtype=torch.int32, device="cuda"); dist = torch.full_like(idx, -7)
            ex.match_keypoints(qd, nq, td, nt, idx, dist, k=k, stream=st)
            torch.cuda.synchronize()
            res[f"idx_{nq}_{nt}_{k}"] = idx.cpu().numpy(); res[f"dist_{nq}_{nt}_{k}"] = dist.cpu().numpy()
    np.savez(out, **res)
    if os.environ.get("UMMA_PROBE_TIME"):
        g = torch.Generator(device="cuda").manual_seed(1)
        for nq, nt in ((257020, 50000), (100000, 200000), (2000, 2000)):
            q = torch.randint(0, 256, (nq, 32), dtype=torch.uint8, device="cuda", generator=g)
            t = torch.randint(0, 256, (nt, 32), dtype=torch.uint8, device="cuda", generator=g)
            idx = torch.zeros((nq, 2), dtype=torch.int32, device="cuda"); dist = torch.zeros_like(idx)
            for k in (1, 2):
                for _ in range(2):
                    ex.match_keypoints(q, nq, t, nt, idx, dist, k=k, stream=st)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                it = 5 if nq * nt > 1e9 else 200
                e0.record(st)
                for _ in range(it):
                    ex.match_keypoints(q, nq, t, nt, idx, dist, k=k, stream=st)
                e1.record(st)
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / it
                print(f"  nq={nq} nt={nt} k={k}: {ms:.4f} ms, {nq * nt / ms / 1e6:.1f} Gpairs/s", flush=True)


def run(name, env, out, time_it=False):
    e = dict(os.environ); e.update(env)
    if time_it:
        e["UMMA_PROBE_TIME"] = "1"
    try:
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "child", out], env=e, capture_output=True, text=True, timeout=240)
    except subprocess.TimeoutExpired:
        print(f"{name}: TIMEOUT", flush=True)
        return None
    if r.returncode != 0:
        print(f"{name}: exit {r.returncode}: {r.stderr.strip().splitlines()[-3:]}", flush=True)
        return None
    if r.stdout.strip():
        print(f"{name} timing:\n{r.stdout.rstrip()}", flush=True)
    return np.load(out)


def main():
    tmp = os.path.join(ROOT, "gpurun_out")
    os.makedirs(tmp, exist_ok=True)
    ref = run("popc", {"ORBB_MATCH_POPC": "1"}, os.path.join(tmp, "umma_ref.npz"))
    if ref is None:
        sys.exit(1)
    good = []
    for name, env in (("umma (tiles expanded inside the CTAs)", {"ORBB_MATCH_UMMA": "1", "ORBB_MATCH_PRE": "0"}),
                      ("umma+pre (train set expanded once, B tiles by bulk copy: the default)", {"ORBB_MATCH_UMMA": "1", "ORBB_MATCH_PRE": "1"}),
                      ("umma(plain epilogue)", {"ORBB_MATCH_UMMA": "2"}),
                      ("imma", {"ORBB_MATCH_UMMA": "0"})):
        if os.environ.get("UMMA_PROBE_ONLY") and not any(name.startswith(o) for o in os.environ["UMMA_PROBE_ONLY"].split(",")):
            continue
        got = run(name, env, os.path.join(tmp, "umma_got.npz"))
        if got is None:
            continue
        bad = [k for k in ref.files if not np.array_equal(ref[k], got[k])]
        if bad:
            k0 = bad[0]
            diff = np.argwhere(ref[k0] != got[k0])
            print(f"{name}: MISMATCH in {len(bad)}/{len(ref.files)} arrays ({bad}); first {k0}: {len(diff)} entries differ, e.g. row {diff[0].tolist()} "
                  f"ref {ref[k0][tuple(diff[0])]} got {got[k0][tuple(diff[0])]}", flush=True)
            for kk in bad[:4]:  # where the wrong rows sit inside a 256-query CTA (warp = 32 rows), and a few of them
                rows = np.unique(np.argwhere(ref[kk] != got[kk])[:, 0])
                print(f"   {kk}: {len(rows)}/{len(ref[kk])} rows wrong; per warp slot {np.bincount((rows % 256) // 32, minlength=8).tolist()}; "
                      f"rows {rows[:6].tolist()} ref {ref[kk][rows[:6]].tolist()} got {got[kk][rows[:6]].tolist()}", flush=True)
        else:
            print(f"{name}: all {len(ref.files)} result arrays identical to the POPC matcher", flush=True)
            good.append((name, env))
    for name, env in good:
        if os.environ.get("UMMA_PROBE_NOTIME"):
            break
        if name.startswith("umma") and "plain" not in name:
            run(name, env, os.path.join(tmp, "umma_got.npz"), time_it=True)
    n_forms = 4 if not os.environ.get("UMMA_PROBE_ONLY") else None
    for f in ("umma_ref.npz", "umma_got.npz"):
        try:
            os.remove(os.path.join(tmp, f))
        except OSError:
            pass
    if n_forms is not None and len(good) != n_forms:
        sys.exit(1)


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "child":
        child(sys.argv[2])
    else:
        main()
