class Polygon:  # placeholder; evaluation code is out of scope
    def __init__(self, *a, **k):
        raise NotImplementedError("shapely stub")
