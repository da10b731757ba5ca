"""Backbone registry -- mirror of hdn/models/backbone/__init__.py:10-22.  Only the ResNet-atrous family is
provided: no shipped configuration selects AlexNet / MobileNetV2 (component 4)."""
from hdn.models.backbone.resnet_atrous import resnet18, resnet34, resnet50

BACKBONES = {"resnet18": resnet18, "resnet34": resnet34, "resnet50": resnet50}


def get_backbone(name, **kwargs):
    if name not in BACKBONES:
        raise KeyError("backbone %r is not part of the inference path mirrored here (have: %s)" % (name, sorted(BACKBONES)))
    return BACKBONES[name](**kwargs)
