"""Head registry -- mirror of hdn/models/head/__init__.py:9-19."""
from hdn.models.head.ban import DepthwiseBAN, MultiBAN
from hdn.models.head.ban_lp import DepthwiseCircBAN, MultiCircBAN

BANS = {"DepthwiseBAN": DepthwiseBAN, "MultiBAN": MultiBAN, "DepthwiseCircBAN": DepthwiseCircBAN, "MultiCircBAN": MultiCircBAN}


def get_ban_head(name, **kwargs):
    return BANS[name](**kwargs)
