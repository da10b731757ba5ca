"""`HomoModelBuilder` -- mirror of Oneline_DLTv1/models/homo_model_builder.py:96-113: the container that owns the
homography estimator's sub-modules (ShareFeature, backbone, avgpool, fc -> 8 corner offsets).  State-dict keys:
ShareFeature.ShareFeature.{0..7}.*, backbone.*, fc.{weight,bias}.

Its reference `forward` (:115-217) is the TRAINING path (losses on triplets); inference re-implements the chain
inline in ModelBuilder.track_proj (model_builder...py:161-217), which is what this repo mirrors.  `forward` here
is the inference chain for a batch: gray pair -> offsets -> (K5) H -> (K4) warped template.
"""
import torch
import torch.nn as nn

from hdn.core.config import cfg
from homo_estimator.Deep_homography.Oneline_DLTv1.backbone import get_backbone
from homo_estimator.Deep_homography.Oneline_DLTv1.preprocess import get_pre
from homo_estimator.Deep_homography.Oneline_DLTv1.utils import DLT_solve, dlt_warp, transform  # noqa: F401


class HomoModelBuilder(nn.Module):
    def __init__(self, pretrained=False):
        super().__init__()
        self.ShareFeature = get_pre("PreShareFeature")
        name = cfg.BACKBONE_HOMO.TYPE
        self.backbone = get_backbone(name, pretrained, **cfg.BACKBONE_HOMO.KWARGS)
        self.avgpool = nn.AdaptiveAvgPool2d(1)
        self.fc = nn.Linear(512 if name in ("resnet18", "resnet34") else 2048, 8)

    def offsets(self, input_tensors):
        """[B,2,h,w] gray (template, search) -> (offsets [B,8], patch_1, patch_2)."""
        p1 = self.ShareFeature(input_tensors[:, :1])
        p2 = self.ShareFeature(input_tensors[:, 1:])
        y = self.backbone(torch.cat((p1, p2), dim=1))
        return self.fc(self.avgpool(y).flatten(1)), p1, p2

    def forward(self, data):
        off, p1, p2 = self.offsets(data["input_tensors"])
        H, warped = dlt_warp(data["h4p"], off, data["org_imgs"][:, :1])
        return {"offsets": off, "H": H, "pred_I2": warped, "patch_1": p1, "patch_2": p2}
