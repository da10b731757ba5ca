"""Heads of the log-polar (scale / rotation) branch -- mirror of hdn/models/head/ban_lp.py.

Identical to the similarity heads except that the correlation wraps along the angle axis and replicates along
the log-radius axis (xcorr_depthwise_circular, ban_lp.py:38) and the localisation branch has 4 channels (:47).
"""
from hdn.core.xcorr import xcorr_depthwise_circular
from hdn.models.head.ban import DepthwiseBAN, DepthwiseXCorr, MultiBAN


class DepthwiseXCorrCirc(DepthwiseXCorr):
    correlate = staticmethod(xcorr_depthwise_circular)
    circular = True


class DepthwiseCircBAN(DepthwiseBAN):
    branch = DepthwiseXCorrCirc
    loc_channels = 4


class MultiCircBAN(MultiBAN):
    level_head = DepthwiseCircBAN
