"""Global configuration node `cfg` -- mirror of hdn/core/config.py (reference: yacs CfgNode, 554 lines of defaults).

Same access pattern (`cfg.TRACK.INSTANCE_SIZE`, `cfg.CUDA = False`, `cfg.merge_from_file(yaml)`), same key names
and default values for every key the inference path reads (reference lines cited per section).  Training /
dataset sections are not pre-declared: the node accepts new keys when a YAML brings them (the reference would
reject unknown keys; that strictness only matters to training).  No yacs dependency.
"""
import copy

import yaml


class CfgNode(dict):
    """dict with attribute access and recursive merge."""

    def __init__(self, init=None, new_allowed=True):
        super().__init__()
        for k, v in (init or {}).items():
            self[k] = CfgNode(v) if isinstance(v, dict) else v

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError("config has no key %r" % k)

    def __setattr__(self, k, v):
        self[k] = v

    def update_from(self, other):
        for k, v in other.items():
            if isinstance(v, dict) and isinstance(self.get(k), CfgNode):
                self[k].update_from(v)
            else:
                self[k] = CfgNode(v) if isinstance(v, dict) else v
        return self

    def merge_from_file(self, path):
        with open(path, "r") as fh:
            self.update_from(yaml.safe_load(fh) or {})

    def merge_from_list(self, pairs):
        for key, val in zip(pairs[0::2], pairs[1::2]):
            node = self
            *parents, leaf = key.split(".")
            for p in parents:
                node = node.setdefault(p, CfgNode())
            node[leaf] = val

    def clone(self):
        return copy.deepcopy(self)

    def freeze(self):  # yacs API compatibility; values stay assignable like the reference's usage needs
        return self

    defrost = freeze


_DEFAULTS = {
    "META_ARC": "hdn_r50_l234",                       # config.py:12
    "CUDA": True,                                     # :14
    "BASE": {"PROJ_PATH": "./", "BASE_PATH": "./", "DATA_PATH": "./", "DATA_ROOT": "./"},  # :19-25 (author's home dirs there)
    "TRAIN": {"EXEMPLAR_SIZE": 127, "SEARCH_SIZE": 255, "BASE_SIZE": 8, "OUTPUT_SIZE": 25, "OUTPUT_SIZE_LP": 13, "BATCH_SIZE": 32},  # :30-56
    "BACKBONE": {"TYPE": "res50", "KWARGS": {}, "PRETRAINED": "", "TRAIN_LAYERS": ["layer2", "layer3", "layer4"], "LAYERS_LR": 0.1,
                 "TRAIN_EPOCH": 10, "IF_PRETRAINED": False},                                           # :398-420
    "BACKBONE_HOMO": {"TYPE": "res34", "KWARGS": {}, "PRETRAINED": "", "TRAIN_LAYERS": ["layer2", "layer3", "layer4"], "LAYERS_LR": 0.1,
                      "TRAIN_EPOCH": 10, "IF_PRETRAINED": False},                                      # :424-445
    "ADJUST": {"ADJUST": True, "KWARGS": {}, "HOMO_KWARGS": {}, "TYPE": "AdjustAllLayer"},            # :449-458
    "BAN": {"BAN": False, "TYPE": "MultiBAN", "KWARGS": {}},                                           # :462-470
    "BAN_LP": {"BAN": False, "TYPE": "MultiCircBAN", "KWARGS": {}},                                    # :475-483
    "HOMO_CORR": {"CORR": False, "TYPE": "HomoCorr", "KWARGS": {}},                                    # :490-498
    "POINT": {"STRIDE": 8, "STRIDE_LP": 8},                                                            # :505-509
    "TRACK": {"TYPE": "hdnTracker", "PENALTY_K": 0.14, "PENALTY_K_LP": 0.14, "WINDOW_INFLUENCE": 0.45, "LR": 0.3, "LR_LP": 0.6,
              "EXEMPLAR_SIZE": 127, "INSTANCE_SIZE": 255, "SHALLOW_SIZE": 127, "BASE_SIZE": 8, "CONTEXT_AMOUNT": 0.5, "CONTEXT_MUL": 2.0,
              "HOMO_CONTEXT": 1.0, "BASE_SC_FAC": 2.0, "SCALE_SCORE_THRESH": 0.5},                     # :515-554
}

cfg = CfgNode(_DEFAULTS)
