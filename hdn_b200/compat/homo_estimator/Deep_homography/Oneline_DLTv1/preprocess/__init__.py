"""Pre-processing registry -- mirror of Oneline_DLTv1/preprocess/__init__.py:18-26 (MaskGenerator is never called on
the tracking path, model_builder...py:184-186, and is not mirrored)."""
from homo_estimator.Deep_homography.Oneline_DLTv1.preprocess.input_feature_extractor import PreShareFeature

head = {"PreShareFeature": PreShareFeature}


def get_pre(name, **kwargs):
    return head[name](**kwargs)
