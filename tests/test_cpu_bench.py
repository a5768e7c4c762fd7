"""CPU suite: bench.py's contract without a GPU -- the reference arm prints one JSON line with the agreed keys, and the
product arm refuses to run (there is no CPU fallback for the hot path)."""
import json
import os
import subprocess
import sys

from conftest import ROOT


def _run(*args, timeout=600):
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=timeout, env=env, cwd=ROOT)


def test_reference_arm_emits_one_json_line():
    # (the real arm runs the whole 1000-step sampler once, ~2 min; the contract is checked on a shortened schedule at C2)
    p = _run("--impl", "reference", "--steps", "1", "--warmup", "0", "--config", "C2", "--ddpm-steps", "20")
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "shapes/s" and d["higher_is_better"] is True
    assert d["metric"] == "shapes/sec end-to-end (1000-step sample + 512\u00b3 UDF extract) at 1/2/4/8 B200"     # BASELINE.json, verbatim
    assert d["value"] > 0 and d["steps"] == 1 and d["warmup"] == 0
    assert d["e2e"] == {"value": d["value"], "unit": "shapes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["cores"] >= 1 and cb["kind"] in ("port", "port+reference-mc", "reference") and cb["sample"]
    assert d["config"]["resolution"] == 256 and d["config"]["ddpm_steps"] == 20 and d["config"]["batch_per_gpu"] == 8
    assert d["config"]["name"] == "C2" and "measured once" in d["config"]["workload"]


def test_default_configuration_is_the_metric():
    sys.path.insert(0, ROOT)
    import importlib
    b = importlib.import_module("bench")
    base = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    assert b.METRIC == base["metric"]
    assert b.CONFIGS["C3"]["res"] == 512 and b.CONFIGS["C3"]["batch"] == 8 and b.CONFIG == "C3" and b.RES == 512
    assert b.CONFIGS["C5"]["guidance"] == 4.0 and b.CONFIGS["C4"]["latent"] == 64 and b.CONFIGS["C4"]["batch"] == 4


def test_product_arm_has_no_cpu_fallback():
    p = _run("--steps", "1", "--warmup", "0", timeout=300)
    assert p.returncode != 0
    assert "CUDA" in (p.stderr + p.stdout)
