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
