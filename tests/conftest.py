import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    import torch

    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    d = np.load(os.path.join(GOLDEN, name + ".npz"))
    return {k: d[k] for k in d.files}


def golden_names(prefix):
    return sorted(f[:-4] for f in os.listdir(GOLDEN) if f.startswith(prefix) and f.endswith(".npz"))


def regen_image(g):
    """K3 goldens store big inputs as (seed, shape): same generator as oracle/gen_golden.py."""
    if "img" in g:
        return g["img"]
    shape = tuple(int(v) for v in g["shape"])
    return (np.random.default_rng(int(g["seed"])).random(shape) * 255.0).astype(np.float32)


# Parity criterion of SURVEY.md section 8(d): rtol 1e-3, atol 1e-4 * max|ref|.
def assert_close(got, ref, rtol=1e-3, atol_scale=1e-4, what=""):
    got = np.asarray(got, np.float64)
    ref = np.asarray(ref, np.float64)
    assert got.shape == ref.shape, (what, got.shape, ref.shape)
    atol = atol_scale * max(float(np.max(np.abs(ref))), 1e-30)
    err = np.abs(got - ref) - (atol + rtol * np.abs(ref))
    assert np.all(err <= 0), "%s: %d / %d out of tolerance, worst excess %.3e (atol %.3e)" % (
        what, int(np.sum(err > 0)), err.size, float(err.max()), atol)
