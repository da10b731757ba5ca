"""Seeded synthetic inputs shared by the golden generator (reference side) and the tests (hdn_b200 side)."""
import cv2
import numpy as np


def crop_tensor(seed, shape):
    """A smooth-ish random BGR crop in [0,255] as float32 NCHW (what get_subwindow would hand the model)."""
    rng = np.random.default_rng(seed)
    b, c, h, w = shape
    coarse = rng.random((b, c, h // 8 + 2, w // 8 + 2)).astype(np.float32)
    out = np.empty(shape, np.float32)
    for i in range(b):
        for j in range(c):
            up = cv2.resize(coarse[i, j], (w, h), interpolation=cv2.INTER_CUBIC)
            out[i, j] = np.clip(up * 200.0 + 27.0 + rng.standard_normal((h, w)).astype(np.float32) * 6.0, 0, 255)
    return out


def texture(seed, h, w):
    """A planar texture with structure at several scales (uint8 BGR)."""
    rng = np.random.default_rng(seed)
    img = np.zeros((h, w, 3), np.float32)
    for cells in (4, 9, 19, 41):
        layer = rng.random((cells, cells, 3)).astype(np.float32)
        img += cv2.resize(layer, (w, h), interpolation=cv2.INTER_NEAREST if cells > 15 else cv2.INTER_CUBIC) / 4.0
    return np.clip(img * 255.0, 0, 255).astype(np.uint8)


def sequence(seed, n_frames, size=(360, 480), obj=(120, 160)):
    """A planar object on a background, moved by a smooth random homography walk.
    -> frames [n] uint8 BGR, gt polygons [n,8] (x1,y1,..,x4,y4 TL,TR,BR,BL of the object)."""
    rng = np.random.default_rng(seed)
    H_img, W_img = size
    oh, ow = obj
    bg = texture(seed + 1, H_img, W_img)
    fg = texture(seed + 2, oh, ow)
    x0, y0 = (W_img - ow) / 2.0, (H_img - oh) / 2.0
    base = np.array([[x0, y0], [x0 + ow, y0], [x0 + ow, y0 + oh], [x0, y0 + oh]], np.float32)
    frames, polys = [], []
    cur = base.copy()
    vel = np.zeros((4, 2), np.float32)
    src = np.array([[0, 0], [ow, 0], [ow, oh], [0, oh]], np.float32)
    for t in range(n_frames):
        if t > 0:
            vel = 0.7 * vel + rng.normal(0, 0.9, (4, 2)).astype(np.float32) + rng.normal(0, 1.2, (1, 2)).astype(np.float32)
            cur = cur + vel
        Hm = cv2.getPerspectiveTransform(src, cur)
        frame = bg.copy()
        warped = cv2.warpPerspective(fg, Hm, (W_img, H_img))
        mask = cv2.warpPerspective(np.full((oh, ow), 255, np.uint8), Hm, (W_img, H_img))
        frame[mask > 127] = warped[mask > 127]
        frames.append(frame)
        polys.append(cur.reshape(-1).copy())
    return frames, np.asarray(polys, np.float32)
