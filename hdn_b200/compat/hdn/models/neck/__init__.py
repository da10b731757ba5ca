"""Neck registry -- mirror of hdn/models/neck/__init__.py:13-19."""
from hdn.models.neck.neck import AdjustAllLayer, AdjustLayer

NECKS = {"AdjustLayer": AdjustLayer, "AdjustAllLayer": AdjustAllLayer}


def get_neck(name, **kwargs):
    return NECKS[name](**kwargs)
