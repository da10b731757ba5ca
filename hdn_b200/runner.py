"""Tracker-level runner: whole sequences through the mirrored reference API, sharded over GPUs.

    python -m hdn_b200.runner [--sequences 8] [--frames 60] [--size 720x1280] [--snapshot model.pth] [--graphs 1]
    torchrun --nproc-per-node N -m hdn_b200.runner ...          (one rank per GPU, sequences round-robin)

This is BASELINE.json config 4 (a POT-shaped stream: per-sequence serial tracking at the native 127/255 crops,
sequences spread over the GPUs).  Each sequence goes through exactly what tools/test.py does per video
(tools/test.py:115-172): `tracker.init` on frame 0 from the ground-truth polygon, `tracker.track_new` on every later
frame.  Frames of one sequence are serially dependent (H_total feedback, hdn_tracker_proj_e2e.py:154,262-266), so
the unit of sharding is the sequence; the only exchange is the final gather of the predicted polygons.

Without --snapshot the weights are the seeded fixture of hdn_b200.synthetic (no checkpoint ships with the
reference); sequences are synthetic planar objects under a smooth random homography walk.
Prints ONE JSON line on rank 0: frames/s per rank and aggregate, per-stage host/device time split.
"""
import argparse
import json
import os
import sys
import time

import numpy as np


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sequences", type=int, default=8)
    ap.add_argument("--frames", type=int, default=60)
    ap.add_argument("--size", default="720x1280", help="frame HxW (POT videos are 1280x720)")
    ap.add_argument("--snapshot", default="")
    ap.add_argument("--config", default=os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "experiments", "tracker_homo_config",
                                                     "proj_e2e_GOT_unconstrained_v2.yaml"))
    ap.add_argument("--graphs", type=int, default=1, help="replay the three network stages from CUDA graphs")
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--lockstep", type=int, default=0, help="advance this many sequences per GPU in lock-step (0 = one at a time)")
    ap.add_argument("--results", default="", help="directory for per-sequence result files (8 corner coordinates per line, the format tools/test.py writes)")
    return ap.parse_args()


def build(cfg_path, snapshot, graphs):
    import torch
    from hdn_b200 import compat, synthetic
    compat.activate()
    from hdn.core.config import cfg
    cfg.merge_from_file(cfg_path)
    cfg.CUDA = True
    from hdn.models.model_builder_e2e_unconstrained_v2 import ModelBuilder
    from hdn.tracker.tracker_builder import build_tracker
    from hdn.utils.model_load import load_pretrain
    model = ModelBuilder()
    model = load_pretrain(model, snapshot) if snapshot else synthetic.fill_weights(model)
    model = model.cuda().eval()
    if graphs:
        model.enable_graphs()
    return build_tracker(model), model


def track_sequence(tracker, frames, polys):
    """-> polygons [n-1, 4, 2], seconds spent in track_new."""
    from hdn.utils.bbox import get_min_max_bbox, get_w_h_from_poly
    import torch
    out = []
    gt = polys[0]
    cx, cy, w, h = get_min_max_bbox(np.array(gt))
    tracker.init(frames[0], [cx - (w - 1) / 2, cy - (h - 1) / 2, w, h], get_w_h_from_poly(np.array(gt)), gt, np.array([gt[:2]]))
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for idx in range(1, len(frames)):
        out.append(np.asarray(tracker.track_new(idx, frames[idx], None, None, None)["polygon"], np.float64))
    torch.cuda.synchronize()
    return np.asarray(out), time.perf_counter() - t0


def first_frame_args(gt):
    """What tools/test.py:118-130 derives from the first ground-truth polygon."""
    from hdn.utils.bbox import get_min_max_bbox, get_w_h_from_poly
    cx, cy, w, h = get_min_max_bbox(np.array(gt))
    return [cx - (w - 1) / 2, cy - (h - 1) / 2, w, h], get_w_h_from_poly(np.array(gt)), gt, np.array([gt[:2]])


def track_lockstep(model, seqs):
    """seqs: list of (frames, polys) of equal length -> (polygons [S, n-1, 4, 2], seconds in track_new)."""
    import torch
    from hdn_b200.batched import LockstepTrackers
    group = LockstepTrackers(model, len(seqs))
    firsts = [first_frame_args(p[0]) for _, p in seqs]
    group.init([f[0] for f, _ in seqs], *[list(col) for col in zip(*firsts)])
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = []
    for idx in range(1, len(seqs[0][0])):
        res = group.track_new(idx, [f[idx] for f, _ in seqs])
        out.append([np.asarray(r["polygon"], np.float64) for r in res])
    torch.cuda.synchronize()
    return np.asarray(out).transpose(1, 0, 2, 3), time.perf_counter() - t0


def main():
    a = parse()
    import torch
    from hdn_b200 import shard, synthetic
    rank, local_rank, world = shard.init()
    torch.cuda.set_device(local_rank)
    H, W = (int(v) for v in a.size.lower().split("x"))
    tracker, model = build(a.config, a.snapshot, a.graphs)
    mine = shard.round_robin(a.sequences, rank, world)
    seqs = {s: synthetic.sequence(100 + s, a.frames, size=(H, W), obj=(H // 3, W // 3)) for s in mine}
    # warm-up on a short private sequence (cuDNN autotune, graph capture)
    wf, wp = synthetic.sequence(99, a.warmup + 1, size=(H, W), obj=(H // 3, W // 3))
    track_sequence(tracker, wf, wp)
    if a.lockstep:
        track_lockstep(model, [(wf, wp)] * min(a.lockstep, len(mine)))
    shard.barrier()
    t_all0 = time.perf_counter()
    busy, n_frames, results = 0.0, 0, {}
    if a.lockstep:
        for lo in range(0, len(mine), a.lockstep):
            chunk = mine[lo:lo + a.lockstep]
            polys, dt = track_lockstep(model, [seqs[s] for s in chunk])
            for s, p in zip(chunk, polys):
                results[s] = p
            busy += dt
            n_frames += polys.shape[0] * polys.shape[1]
    else:
        for s in mine:
            polys, dt = track_sequence(tracker, *seqs[s])
            results[s] = polys
            busy += dt
            n_frames += len(polys)
    torch.cuda.synchronize()
    wall = time.perf_counter() - t_all0
    wall_max = shard.max_over_ranks(wall, torch.device("cuda", local_rank))
    total_frames = a.sequences * (a.frames - 1)
    # the one exchange: every rank's polygons to rank 0 (fixed-size slots, sequence id -> slot)
    slot = torch.zeros((a.sequences, a.frames - 1, 8), dtype=torch.float64, device="cuda")
    for s, p in results.items():
        slot[s] = torch.from_numpy(p.reshape(a.frames - 1, 8)).cuda()
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(slot)  # disjoint slots -> sum == gather
    accuracy = None
    if rank == 0:
        # accuracy of the gathered trajectories against the synthetic ground truth, scored like the reference's HomoBenchmark
        # (alignment-error precision, first frame dropped); optionally the result files tools/test.py would write
        from toolkit.evaluation import HomoBenchmark
        pred = slot.cpu().numpy()

        class _V:
            pass
        videos = []
        for sid in range(a.sequences):
            polys = seqs[sid][1] if sid in seqs else synthetic.sequence(100 + sid, a.frames, size=(H, W), obj=(H // 3, W // 3))[1]
            gt = np.asarray(polys, np.float64).reshape(a.frames, 8)
            v = _V()
            v.name, v.gt_traj, v.pred_trajs = "seq%03d" % sid, gt, {"hdn_b200": np.concatenate([gt[:1], pred[sid]], axis=0)}
            videos.append(v)
            if a.results:
                os.makedirs(os.path.join(a.results, "hdn_b200"), exist_ok=True)
                with open(os.path.join(a.results, "hdn_b200", v.name + ".txt"), "w") as fh:
                    for x in v.pred_trajs["hdn_b200"]:
                        fh.write(" ".join(str(float(i)) for i in x) + "\n")
        bench = HomoBenchmark(type("DS", (list,), {"tracker_names": ["hdn_b200"], "tracker_path": a.results or None})(videos))
        accuracy = HomoBenchmark.summary(bench.eval_4pts_precision()["hdn_b200"])
    if rank == 0:
        line = {"metric": "tracker frames/sec (hdnTrackerHomo.track_new, native 127/255 crops)", "value": total_frames / wall_max, "unit": "frames/s",
                "n_gpus": world, "sequences": a.sequences, "frames_per_sequence": a.frames, "frame_size": [H, W],
                "per_rank_fps": n_frames / busy if busy else None, "ms_per_frame": 1e3 * busy / max(n_frames, 1), "cuda_graphs": bool(a.graphs), "lockstep_per_gpu": a.lockstep,
                "weights": a.snapshot or "seeded fixture (hdn_b200.synthetic.fill_weights)", "data": "synthetic homography walk",
                "polygon_checksum": float(slot.abs().sum().item()),
                "accuracy_vs_synthetic_gt": accuracy,
                "accuracy_note": "HomoBenchmark alignment-error precision; %s" % ("checkpoint " + a.snapshot if a.snapshot else "UNTRAINED seeded weights -- the numbers only exercise the scoring path, parity with the reference on the same weights is what tests/test_gpu_model.py checks")}
        print(json.dumps(line), flush=True)
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == "__main__":
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    main()
