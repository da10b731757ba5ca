"""Thin stand-in for toolkit.datasets (the reference's benchmark readers, SURVEY 2 component 20: out of scope to
re-implement).  Only what tools/test.py needs to iterate a POT-format benchmark is provided:
`DatasetFactory.create_dataset(name=, dataset_root=, load_img=)` -> iterable of videos, each iterable as
(BGR frame, ground truth) with `.name`.  JSON schema = toolkit/datasets/pot.py:70-89
(video_dir, init_rect, img_names, gt_rect, flag, homography)."""
import json
import os

import cv2


class Video:
    def __init__(self, name, root, meta, load_img=False):
        self.name = name
        self.video_dir = meta.get("video_dir", name)
        self.init_rect = meta["init_rect"]
        self.gt_traj = meta["gt_rect"]
        self.attr = "0"
        self.tags = {"flag": meta.get("flag"), "homo": meta.get("homography")}
        self.img_names = [os.path.join(root, p) for p in meta["img_names"]]
        self.imgs = [cv2.imread(p) for p in self.img_names] if load_img else None
        first = self.imgs[0] if self.imgs else cv2.imread(self.img_names[0])
        if first is None:
            raise FileNotFoundError(self.img_names[0])
        self.height, self.width = first.shape[:2]
        self.pred_trajs = {}
        self.tracker_names = []

    def load_tracker(self, path, tracker_names=None, store=True):
        """Results of `tracker_names` for this video (toolkit/datasets/pot.py:35-67, video.py:31-56): one line of 8 space-separated
        corner coordinates per frame.  `<path>/<tracker>/<video>.txt` is what tools/test.py:237-243 writes (every frame);
        `<path>/<tracker>/<video>_<tracker>.txt` is the POT submission layout, of which the reference keeps frame 0 and the odd frames."""
        if isinstance(tracker_names, str):
            tracker_names = [tracker_names]
        for name in tracker_names or sorted(os.listdir(path)):
            plain = os.path.join(path, name, self.name + ".txt")
            pot = os.path.join(path, name, "%s_%s.txt" % (self.name, name))
            src = plain if os.path.exists(plain) else (pot if os.path.exists(pot) else None)
            if src is None:
                continue
            with open(src) as fh:
                traj = [[float(v) for v in line.replace(",", " ").split()] for line in fh if line.strip()]
            if src == pot:
                traj = [x for i, x in enumerate(traj) if i == 0 or i % 2 == 1]
            if not store:
                return traj
            self.pred_trajs[name] = traj
        self.tracker_names = list(self.pred_trajs)

    def __len__(self):
        return len(self.img_names)

    def __getitem__(self, i):
        return (self.imgs[i] if self.imgs else cv2.imread(self.img_names[i])), self.gt_traj[i]

    def __iter__(self):
        for i in range(len(self)):
            yield self[i]


class POTDataset:
    def __init__(self, name, dataset_root, load_img=False, eval_mode=False):
        self.name, self.dataset_root = name, dataset_root
        with open(os.path.join(dataset_root, name + ".json")) as fh:
            meta = json.load(fh)
        self.videos = {k: Video(k, dataset_root, v, load_img) for k, v in meta.items()}
        self.tracker_path, self.tracker_names = None, []

    def set_tracker(self, path, tracker_names):
        """toolkit/datasets/dataset.py:23-30."""
        self.tracker_path, self.tracker_names = path, [tracker_names] if isinstance(tracker_names, str) else list(tracker_names)

    def __len__(self):
        return len(self.videos)

    def __getitem__(self, key):
        return self.videos[key] if isinstance(key, str) else self.videos[sorted(self.videos)[key]]

    def __iter__(self):
        for k in sorted(self.videos):
            yield self.videos[k]


class DatasetFactory:
    @staticmethod
    def create_dataset(**kwargs):
        name = kwargs["name"]
        if name[:3] in ("POT", "UCS", "POI"):
            return POTDataset(**kwargs)
        raise Exception("unknow dataset {} (only POT-format benchmarks are mirrored)".format(name))
