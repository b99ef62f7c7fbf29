"""bench.py contract on a machine without a GPU: the reference arm (the unmodified reference from oracle/_ref - or the oracle
port when that copy is missing - on the host cores) prints ONE JSON line with the keys the driver reads, and the product arm
refuses to run without a CUDA device (no CPU fallback)."""
import json
import os
import subprocess
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args, timeout=600):
    env = dict(os.environ, PYTHONPATH=ROOT)
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, cwd=ROOT,
                          timeout=timeout, env=env)


def test_reference_arm_json_line():
    from oracle import ref_harness as R
    R.install()                       # no-op without /root/reference (then the arm falls back to the oracle port)
    r = _run("--impl", "reference", "--steps", "1", "--warmup", "0")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.strip().splitlines() if ln.startswith("{")]
    assert len(lines) == 1, r.stdout[-2000:]
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 1 and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["metric"] == "composited_ray_samples_x_slots_per_sec" and d["unit"] == "sample-slots/s" and d["data"] == "synthetic"
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["scaling"] == "weak"
    assert d["config"]["workload"] == "room_chair_eval_K5_F128_D64_N4" and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == ("reference" if R.available() else "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["unit"] == d["unit"] and cb["sample"]
    e = d["e2e"]
    assert e["value"] == d["value"] and e["unit"] == d["unit"] and e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0
    assert d["gpu_launches"] == 0
    sys.path.insert(0, ROOT)
    import bench
    assert d["config"] == bench.headline_config()      # the two arms print the same `config`


def test_product_arm_fails_loudly_without_gpu():
    if torch.cuda.is_available():
        return                       # on a GPU box the product arm runs; the -m gpu tests and the bench cover it
    for extra in ([], ["--mode", "train"]):
        r = _run("--steps", "1", "--warmup", "1", *extra, timeout=300)
        assert r.returncode != 0
        assert "CUDA" in (r.stderr + r.stdout)
