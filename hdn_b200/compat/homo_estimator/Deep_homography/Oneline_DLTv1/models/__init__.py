"""Part of the hdn_b200 mirror of the reference API (see hdn_b200/compat/__init__.py)."""
