"""Stub for memory_profiler: `profile` is an identity decorator."""
def profile(fn=None, **k):
    if fn is None:
        return lambda f: f
    return fn
