"""Re-export of the deterministic weight fixture (hdn_b200/synthetic.py) for the tests."""
from hdn_b200.synthetic import GATES, SCALES, fill_weights as fill  # noqa: F401
