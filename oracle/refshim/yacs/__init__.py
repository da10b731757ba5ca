"""Stand-in for the `yacs` package, which is not installed in this image.

TEST INFRASTRUCTURE ONLY: lets oracle/gen_golden.py import the reference's
hdn/core/config.py (it does `from yacs.config import CfgNode`).  The product
has its own config node (hdn_b200/compat/hdn/core/config.py).
"""
