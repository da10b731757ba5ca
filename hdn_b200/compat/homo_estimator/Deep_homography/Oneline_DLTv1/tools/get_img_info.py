"""Host-side packing of the homography estimator's inputs -- mirror of Oneline_DLTv1/tools/get_img_info.py:8-103.

`get_template_info` / `get_search_info` (:8-37, :42-70): BGR crop tensor [1,3,h,w] -> float64 NumPy gray
((img - mean) / std, averaged over channels) [1,127,127] plus the raw CHW copy.
`merge_tmp_search` (:72-103): stacks the pair and emits the constants the network call needs
(`patch_indices` = arange(127*127), `four_points` = the patch corners in (TL, BL, BR, TR) order).
"""
import cv2
import numpy as np

_MEAN = np.array([118.93, 113.97, 102.60]).reshape(1, 1, 3)
_STD = np.array([69.85, 68.81, 72.45]).reshape(1, 1, 3)
_SIDE = 127


def _gray_info(batch):
    hwc = batch[0].cpu().permute(1, 2, 0).numpy()
    if hwc.shape[0] != _SIDE or hwc.shape[1] != _SIDE:
        hwc = cv2.resize(hwc, (_SIDE, _SIDE))
    shown = np.transpose(hwc.copy(), [2, 0, 1])
    gray = np.mean((hwc - _MEAN) / _STD, axis=2, keepdims=True)
    return np.transpose(gray, [2, 0, 1]), shown


def get_template_info(template):
    return _gray_info(template)


def get_search_info(search):
    return _gray_info(search)


def merge_tmp_search(tmp, search):
    pair = np.concatenate([tmp, search], axis=0)
    height, width = pair.shape[1], pair.shape[2]
    ys, xs = np.mgrid[0:_SIDE, 0:_SIDE]
    corners = [(0, 0), (0, _SIDE), (_SIDE, _SIDE), (_SIDE, 0)]
    return {"org_imgs": pair, "input_tensors": pair[:, 0:_SIDE, 0:_SIDE], "patch_indices": (ys.reshape(-1) * width + xs.reshape(-1)),
            "four_points": np.reshape(corners, (-1))}
