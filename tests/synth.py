"""Re-export of hdn_b200.synthetic for the tests (kept so test modules can `import synth`)."""
from hdn_b200.synthetic import crop_tensor, sequence, texture  # noqa: F401
