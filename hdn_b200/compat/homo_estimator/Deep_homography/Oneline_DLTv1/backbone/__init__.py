"""Homography-estimator backbone factory -- mirror of Oneline_DLTv1/backbone/__init__.py:20-52.

`pretrained=True` in the reference downloads ImageNet weights through model_zoo (:40-49) and drops conv1/fc.
The tracker always overwrites every weight from its own checkpoint right after construction
(tools/test.py:69), so the download only matters to training; here `pretrained` is accepted and ignored
(no network access on the hot path)."""
from homo_estimator.Deep_homography.Oneline_DLTv1.backbone import resnet


def get_backbone(model_name, pretrained=False, **kwargs):
    return resnet.build(model_name, **kwargs)
