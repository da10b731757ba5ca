"""CPU-only: the parts of bench.py's contract that do not need a GPU -- the reference arm's JSON line, the refusal to run the
GPU arm on a CPU (no fallback), and the algorithmic-byte model the roofline block is built on."""
import json
import os
import subprocess
import sys

from conftest import ROOT


def run_bench(*args):
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, env=env, timeout=600)


def test_reference_arm_line():
    res = run_bench("--impl", "reference", "--steps", "1", "--warmup", "1")
    assert res.returncode == 0, res.stderr
    line = json.loads(res.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "frames/sec (corr+warp+DLT forward)" and line["unit"] == "frames/s"
    assert line["higher_is_better"] is True and line["vs_baseline"] is None and line["value"] > 0
    assert line["config"]["workload"].startswith("256/512 crops")
    assert line["e2e"] == {"value": line["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == line["value"] and "pair(s)" in cb["sample"]
    assert line["gpu_launches"] == 0


def test_gpu_arm_refuses_to_run_without_a_device():
    res = run_bench("--steps", "1")
    assert res.returncode != 0
    assert "no CPU fallback" in (res.stderr + res.stdout)


def test_algorithmic_bytes_model():
    """bytes_xcorr = 4*C*(Hx*Wx + h*w + Ho*Wo); M1 per pair at 256/512 = 54.28 MB, at 127/255 = 13.38 MB."""
    from hdn_b200 import engine
    ab = engine.algorithmic_bytes_per_pair("256/512")
    assert ab["k1"] == 6 * 5786624 and ab["k2"] == 6 * 2583552 and ab["k3"] == 3932160
    assert abs(ab["total"] - 54.28e6) < 0.01e6
    nat = engine.algorithmic_bytes_per_pair("127/255")
    assert nat["k1"] == 6 * 1526784 and nat["k2"] == 6 * 519168 and abs(nat["total"] - 13378728) <= 200
    shared = engine.algorithmic_bytes_per_pair("256/512", B=64, shared_template=True)
    assert shared["k1"] < ab["k1"]  # a shared template is read once per batch
    assert engine.xcorr_flops(61, 61, 29, 29, False) == 2 * 256 * 33 * 33 * 841
