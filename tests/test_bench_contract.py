"""CPU-side checks of bench.py's contract: the reference arm prints one JSON line with the keys the driver reads
(run here on a tiny sample), and the product arm refuses to run without a CUDA device instead of falling back."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args, timeout=600):
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True,
                          timeout=timeout, env=env, cwd=ROOT)


def test_reference_arm_line():
    r = _run("--impl", "reference", "--steps", "1", "--warmup", "0", "--no-refgpu", "--frames", "16")
    assert r.returncode == 0, r.stderr[-400:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"].startswith("ORB frames/s") and d["unit"] == "frames/s"
    assert d["value"] > 0 and d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None
    assert d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 0 and d["ms_per_step"] > 0
    assert d["dtype"] == "u8" and d["data"] == "synthetic" and "workload" in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["unit"] == "frames/s" and cb["sample"]
    assert cb["one_core_value"] > 0 and cb["matcher_gpairs"] > 0 and cb["matcher_gpairs_one_core"] > 0
    assert d["config"]["frames_per_gpu_per_step"] == 16 and d["value_stats"]["best"] >= d["value_stats"]["median"] > 0
    e = d["e2e"]
    assert e["value"] == d["value"] and e["unit"] == "frames/s" and e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0


def test_reference_arm_other_ranks_stay_silent():
    env_rank = dict(os.environ, CUDA_VISIBLE_DEVICES="", RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                        "--warmup", "0"], capture_output=True, text=True, timeout=120, env=env_rank, cwd=ROOT)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_product_arm_has_no_cpu_fallback():
    r = _run("--steps", "1", "--warmup", "0")
    assert r.returncode != 0
    assert "no CPU fallback" in (r.stderr + r.stdout)
