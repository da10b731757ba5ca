"""Lock-step tracking of many independent sequences on one GPU.

Frames of ONE sequence are serially dependent (H_total feedback, hdn_tracker_proj_e2e.py:154,262-266), but sequences
are independent, and at batch 1 the backbone leaves most of a B200 idle (13.6 ms for a 255x255 ResNet-50 forward in
fp32, 3.3 ms per frame at batch 8).  `LockstepTrackers` keeps S `hdnTrackerHomo` states on one shared model and
advances them together: every tracker's `_track_steps` generator runs its host-side phase (warp, crop, decode) on a
worker thread until it asks for a network stage; the S requests of a stage are stacked and answered by ONE batched
call (template / stage 1 / stage 2 / stage 3 + K6 read-back), then every generator resumes with its own slice.

Each tracker runs the same code as single-sequence `track_new`, so its trajectory is the single-sequence trajectory up
to cuDNN/cuBLAS choosing different (fp32) algorithms for different batch sizes.
"""
from concurrent.futures import ThreadPoolExecutor


from hdn_b200 import compat

compat.activate()


class LockstepTrackers:
    def __init__(self, model, n, workers=None):
        from hdn.tracker.hdn_tracker_proj_e2e import hdnTrackerHomo
        self.model = model
        self.trackers = [hdnTrackerHomo(model) for _ in range(n)]
        self.pool = ThreadPoolExecutor(max_workers=workers or min(n, 16))

    def __len__(self):
        return len(self.trackers)

    def init(self, imgs, bboxes, polys, gt_points, first_points):
        gens = [t._init_steps(*a) for t, a in zip(self.trackers, zip(imgs, bboxes, polys, gt_points, first_points))]
        self._drive(gens)

    def track_new(self, fr_idx, imgs):
        """One frame for every sequence -> list of the per-sequence result dicts (same dict as hdnTrackerHomo.track_new)."""
        return self._drive([t._track_steps(fr_idx, img) for t, img in zip(self.trackers, imgs)])

    def _drive(self, gens):
        from hdn.tracker.hdn_tracker_proj_e2e import serve, stack_requests
        n = len(gens)
        results = [None] * n

        def advance(i, value, first=False):
            try:
                return ("req",) + tuple(next(gens[i]) if first else gens[i].send(value))
            except StopIteration as done:
                return ("done", done.value, None)

        pending = list(self.pool.map(lambda i: advance(i, None, True), range(n)))
        while True:
            live = [i for i in range(n) if pending[i][0] == "req"]
            for i in range(n):
                if pending[i][0] == "done" and results[i] is None:
                    results[i] = pending[i][1] if pending[i][1] is not None else True
            if not live:
                break
            kinds = {pending[i][1] for i in live}
            if len(kinds) != 1:
                raise RuntimeError("trackers left lock-step: %s" % sorted(kinds))
            answers = serve(self.model, kinds.pop(), stack_requests([pending[i][2] for i in live]))
            nxt = list(self.pool.map(lambda ia: advance(ia[0], ia[1]), zip(live, answers)))
            for i, p in zip(live, nxt):
                pending[i] = p
        return results
