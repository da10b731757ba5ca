"""Checkpoint loading -- mirror of hdn/utils/model_load.py:48-79 (`load_pretrain`) and :82-112 (`restore_from`).

Accepts the reference's checkpoints unchanged: a raw state dict or {'state_dict': ...}, optional 'module.' prefix
(DataParallel), keys containing 'rf' dropped, non-strict load.  Works with or without a GPU."""
import logging

import torch

logger = logging.getLogger("global")


def remove_prefix(state_dict, prefix):
    return {(k[len(prefix):] if k.startswith(prefix) else k): v for k, v in state_dict.items()}


def check_keys(model, pretrained_state_dict):
    ckpt, own = set(pretrained_state_dict.keys()), set(model.state_dict().keys())
    used = own & ckpt
    missing = sorted(k for k in own - ckpt if not k.endswith("num_batches_tracked"))
    if missing:
        logger.info("[Warning] missing keys: %s", missing)
    logger.info("used keys: %d, unused checkpoint keys: %d", len(used), len(ckpt - own))
    if not used:
        raise RuntimeError("load NONE from pretrained checkpoint")
    return True


def _read(path):
    if torch.cuda.is_available():
        dev = torch.cuda.current_device()
        return torch.load(path, map_location=lambda storage, loc: storage.cuda(dev))
    return torch.load(path, map_location="cpu")


def load_pretrain(model, pretrained_path):
    logger.info("load pretrained model from %s", pretrained_path)
    blob = _read(pretrained_path)
    state = remove_prefix(blob["state_dict"] if "state_dict" in blob else blob, "module.")
    try:
        check_keys(model, state)
    except RuntimeError:
        state = {"features." + k: v for k, v in state.items()}  # backbone-only checkpoints (model_load.py:64-71)
        check_keys(model, state)
    state = {k: v for k, v in state.items() if "rf" not in k}
    model.load_state_dict(state, strict=False)
    return model


def restore_from(model, optimizer, ckpt_path):
    blob = _read(ckpt_path)
    merged = model.state_dict()
    state = remove_prefix(blob["state_dict"], "module.")
    check_keys(model, state)
    merged.update(state)
    model.load_state_dict(merged)
    return model, optimizer, 0
