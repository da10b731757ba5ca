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
    res = run_bench("--impl", "reference", "--steps", "2", "--warmup", "1", "--batch", "4")
    assert res.returncode == 0, res.stderr
    line = json.loads(res.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "frames/sec (corr+warp+DLT forward)" and line["unit"] == "frames/s"
    assert line["higher_is_better"] is True and line["vs_baseline"] is None and line["value"] > 0
    assert line["config"]["workload"].startswith("256/512 crops")
    # `value` = the M1 chain (the GPU arm's `value` scope), `e2e` = the fused chain from neck features (the GPU arm's `e2e` scope):
    # each driver-computed ratio compares like with like.  No device is involved: zero bytes cross a bus.
    e = line["e2e"]
    assert e["unit"] == "frames/s" and e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0 and 0 < e["value"] < line["value"]
    assert "fused" in line["config"]["e2e_workload"] and line["config"]["same_config"] is True
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == line["value"] and "pair(s)" in cb["sample"]
    assert cb["fused"]["value"] == e["value"] and cb["fused"]["pairs_per_step"] == 4
    assert line["gpu_launches"] == 0


def test_fused_chain_work_model():
    """Bytes that cross the plug-in boundary per 256/512 pair (neck features + crops in, maps + scores + H out) and the dense work."""
    from hdn_b200 import head_engine as he
    ab = he.algorithmic_bytes_per_pair("256/512")
    assert ab["in"] == 4 * (3 * 256 * (63 * 63 + 31 * 31) + 3 * 512 * 512 + 127 * 127 + 16) and abs(ab["in"] - 18.35e6) < 0.05e6
    fl = he.flops_per_pair("256/512")
    assert fl["conv_search"] == 6 * 2 * 256 * 256 * 9 * (61 * 61 + 29 * 29) and 30e9 < fl["conv_search"] < 35e9


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
