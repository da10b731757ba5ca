"""Stub: the reference imports matplotlib at module scope but never plots on the hot path."""
def use(*a, **k):
    pass
