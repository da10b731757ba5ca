"""CPU: the NumPy restatement of OpenCV's fixed-point resampling (oracle/cv_port.py) against cv2 ITSELF, bit for bit, on the
calls the reference's tracker makes (hdn_tracker_proj_e2e.py:154, base_tracker.py:118, transform.py:98) -- this pins the
oracle the device-side pre-processing kernels (hdn_b200/csrc/preproc.cu) are checked against."""
import cv2
import numpy as np

from oracle import cv_port


def rand_homography(rng, w, h, mag):
    src = np.array([[0, 0], [w, 0], [w, h], [0, h]], np.float32)
    return cv2.getPerspectiveTransform(src, src + rng.normal(0, 20 * mag, (4, 2)).astype(np.float32))


def test_resize_linear_is_bit_exact():
    rng = np.random.default_rng(0)
    for t in range(120):
        h = int(rng.integers(20, 400))
        w = h if t % 2 == 0 else int(rng.integers(20, 400))
        src = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        ds = 127 if t % 3 else 255
        assert np.array_equal(cv2.resize(src, (ds, ds)), cv_port.resize_linear_u8(src, (ds, ds))), (h, w, ds)
    for n in (254, 510):  # exact 2x decimation: INTER_LINEAR is silently INTER_AREA
        src = rng.integers(0, 256, (n, n, 3), dtype=np.uint8)
        assert np.array_equal(cv2.resize(src, (n // 2, n // 2)), cv_port.resize_linear_u8(src, (n // 2, n // 2)))


def test_warp_perspective_is_bit_exact():
    rng = np.random.default_rng(1)
    for t in range(24):
        h, w = (int(rng.integers(50, 300)), int(rng.integers(50, 400))) if t % 3 else (360, 480)
        src = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        M = np.linalg.inv(rand_homography(rng, w, h, 1.0 if t % 2 else 3.0))
        assert np.array_equal(cv2.warpPerspective(src, M, (w, h), borderMode=cv2.BORDER_REPLICATE), cv_port.warp_perspective_u8(src, M)), t
    src = rng.integers(0, 256, (90, 130, 3), dtype=np.uint8)  # identity: the (0, 0) table entry [32767, 0, 0, 1] must still reproduce the image
    assert np.array_equal(cv_port.warp_perspective_u8(src, np.eye(3)), src)


def test_warp_affine_cubic_is_bit_exact():
    rng = np.random.default_rng(2)
    for t in range(16):
        h, w = (int(rng.integers(50, 300)), int(rng.integers(50, 400))) if t % 3 else (360, 480)
        src = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        rot, cx, cy = rng.normal(0, 0.5), rng.uniform(0, w), rng.uniform(0, h)
        cc, ss = np.cos(rot), np.sin(rot)
        M = np.array([[cc, -ss, cx - cx * cc + cy * ss], [ss, cc, cy - cy * cc - cx * ss]])  # transform.py:88-90
        assert np.array_equal(cv2.warpAffine(src, M, (w, h), flags=2, borderMode=cv2.BORDER_REPLICATE), cv_port.warp_affine_cubic_u8(src, M)), t


def test_tables():
    b, c = cv_port.bilinear_table(), cv_port.cubic_table()
    assert b.shape == (32, 32, 4) and np.all(b.sum(-1) == 32768) and list(b[0, 0]) == [32767, 0, 0, 1]
    assert c.shape == (32, 32, 16) and np.all(c.sum(-1) == 32768)


def test_crop_resize_equals_the_trackers_crop_window():
    """The restated crop (frame-coordinate window + mean padding + resize) == the compat tracker's crop_window, which is itself
    pinned to the reference's get_subwindow outputs (tests/test_compat_host.py)."""
    from hdn_b200 import compat, synthetic
    compat.activate()
    from hdn.tracker.base_tracker import crop_window
    frames, _ = synthetic.sequence(11, 2)
    img = frames[1]
    avg = np.mean(img, axis=(0, 1))
    rng = np.random.default_rng(4)
    cases = [([240.3, 180.9], 127, 181.0), ([20.0, 340.5], 255, 363.0), ([200.0, 200.0], 127, 127), ([470.0, 10.0], 255, 510.0)]
    cases += [([float(rng.uniform(-40, 520)), float(rng.uniform(-40, 400))], int(rng.choice([127, 255])), float(np.floor(rng.uniform(60, 600))))
              for _ in range(60)]
    for pos, msz, osz in cases:
        ref, _ = crop_window(img, np.array(pos), msz, osz, avg)
        assert np.array_equal(ref, cv_port.crop_resize(img, pos, msz, osz, avg)), (pos, msz, osz)


def test_gray_normalise_equals_get_search_info():
    from hdn_b200 import compat
    compat.activate()
    import torch
    from homo_estimator.Deep_homography.Oneline_DLTv1.tools.get_img_info import get_search_info
    crop = np.random.default_rng(5).integers(0, 256, (1, 3, 127, 127)).astype(np.float32)
    ref, _ = get_search_info(torch.from_numpy(crop))
    assert np.array_equal(ref, cv_port.gray_normalise(crop))
