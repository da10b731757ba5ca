"""Mirror of hdn/core/xcorr.py: the two depth-wise correlations the heads import by name
(ban.py:10 `xcorr_depthwise`, ban_lp.py:10 `xcorr_depthwise_circular`), bound to the sm_100a kernels."""
from hdn_b200.ops import xcorr_depthwise, xcorr_depthwise_circular, xcorr_depthwise_multi  # noqa: F401

__all__ = ["xcorr_depthwise", "xcorr_depthwise_circular", "xcorr_depthwise_multi"]
