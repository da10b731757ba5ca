"""Log-polar transforms -- mirror of hdn/models/logpolar.py.

`STN_Polar` (logpolar.py:50-134) is the sm_100a kernel (hdn_logpolar_f32): the sampling grid is analytic in the
kernel, nothing is built on the host or uploaded per call.  `getPolarImg` (:11-29) is the host-side cv2.logPolar
used once per template crop.  STN_LinearPolar / Polar_Pick are training-only or unused on the tracking path and
are not mirrored.
"""
import math

import cv2
import numpy as np

from hdn_b200.ops import STN_Polar  # noqa: F401


def getPolarImg(img, original=None):
    """cv2.logPolar of a square crop about its centre, M = W / ln(W/2), bilinear + fill outliers."""
    rows, cols = img.shape[0], img.shape[1]
    scale = cols / math.log(cols / 2)
    centre = tuple(np.round(original)) if original is not None else (rows // 2, cols // 2)
    return cv2.logPolar(img, centre, scale, cv2.WARP_FILL_OUTLIERS + cv2.INTER_LINEAR)


def getLinearPolarImg(img, original=None):
    rows, cols = img.shape[0], img.shape[1]
    centre = tuple(np.round(original)) if original is not None else (rows // 2, cols // 2)
    return cv2.linearPolar(img, centre, cols / 2, cv2.WARP_FILL_OUTLIERS + cv2.INTER_LINEAR)
