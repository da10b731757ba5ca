"""Mirror of the reference's importable packages (`hdn`, `homo_estimator`, a thin `toolkit.datasets`).

The reference's tools (`tools/test.py`, `tools/demo.py`) import these names; putting this directory on
sys.path (``hdn_b200.compat.activate()`` or ``PYTHONPATH=<repo>/hdn_b200/compat``) makes them resolve to
the B200 implementation with no edit to the tools.  Only the inference hot path is mirrored
(SURVEY.md section 8b); training, evaluation and dataset-preparation modules are deliberately absent.
"""
import os
import sys

COMPAT_DIR = os.path.dirname(os.path.abspath(__file__))


def activate():
    """Idempotently put the mirror packages at the front of sys.path."""
    if COMPAT_DIR not in sys.path:
        sys.path.insert(0, COMPAT_DIR)
    return COMPAT_DIR
