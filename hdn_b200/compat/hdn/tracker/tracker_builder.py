"""`build_tracker` -- mirror of hdn/tracker/tracker_builder.py:12-19 (dispatch on cfg.TRACK.TYPE)."""
from hdn.core.config import cfg
from hdn.tracker.hdn_tracker import hdnTracker
from hdn.tracker.hdn_tracker_proj_e2e import hdnTrackerHomo as hdnTrackerHomoProje2e

TRACKS = {"hdnTracker": hdnTracker, "hdnTrackerHomoProje2e": hdnTrackerHomoProje2e}


def build_tracker(model):
    return TRACKS[cfg.TRACK.TYPE](model)
