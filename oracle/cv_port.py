"""NumPy restatement of the three OpenCV resampling routines on the tracker's per-frame path.

TEST INFRASTRUCTURE ONLY (same rule as oracle/hdn_oracle.c): the checker for hdn_b200/csrc/preproc.cu.

The arithmetic lives in a THIRD-PARTY dependency that is not under /root/reference: OpenCV (`opencv-python`, unpinned by the
reference's INSTALL.md; 4.13.0 in this image).  The reference calls it at
    cv2.warpPerspective(img, inv(H_total), (w, h), borderMode=BORDER_REPLICATE)      hdn/tracker/hdn_tracker_proj_e2e.py:154
    cv2.resize(im_patch, (model_sz, model_sz))                                        hdn/tracker/base_tracker.py:118,195
    cv2.warpAffine(img, M, (w, h), flags=2, borderMode=BORDER_REPLICATE)              hdn/utils/transform.py:98 (img_rot_around_center)
Restated here from OpenCV's published algorithm (modules/imgproc/src/imgwarp.cpp: WarpPerspectiveInvoker, warpAffine,
remapBilinear / remapBicubic, initInterTab2D; resize.cpp: resizeGeneric_Invoker / HResizeLinear / VResizeLinear<uchar>,
ResizeAreaFastVec) for 8-bit 3-channel images:
  * all three work in FIXED POINT: 5 fractional bits for the sampling position, 15-bit interpolation weights from a 32x32 table
    whose entries are nudged so that every kernel sums to exactly 2^15 (warps); 11-bit row / column weights and a
    `(((b0 * (S0 >> 4)) >> 16) + ((b1 * (S1 >> 4)) >> 16) + 2) >> 2` vertical pass (resize);
  * INTER_LINEAR resize by exactly 1/2 silently becomes INTER_AREA (2x2 box mean, `(sum + 2) >> 2`).
Pinned: tests/test_cv_port.py compares every function BIT FOR BIT with cv2 itself (which is importable both here and on the
GPU box) on random images, the tracker's own crop geometry and the degenerate cases (borders, exact 2x, identity).
"""
import cv2
import numpy as np


def _clip(x, a, b):
    """OpenCV's clip(x, a, b): clamp to [a, b - 1]."""
    return np.where(x >= a, np.where(x < b, x, b - 1), a)


# ---------------------------------------------------------------------------------------------------- resize
def resize_linear_u8(src, dsize):
    """cv2.resize(src, dsize) (INTER_LINEAR) for uint8 [H, W, C]."""
    H, W = src.shape[:2]
    dw, dh = dsize
    scale_x, scale_y = 1.0 / (dw / W), 1.0 / (dh / H)
    s = src.astype(np.int32)
    if scale_x == 2.0 and scale_y == 2.0:  # resize.cpp: INTER_LINEAR with an exact 2x2 decimation is routed to INTER_AREA
        return ((s[0::2, 0::2] + s[0::2, 1::2] + s[1::2, 0::2] + s[1::2, 1::2] + 2) >> 2).astype(np.uint8)[:dh, :dw]

    def coeffs(n_dst, scale):
        d = np.arange(n_dst)
        f = ((d + 0.5) * scale - 0.5).astype(np.float32)
        i = np.floor(f).astype(np.int32)
        return i, (f - i.astype(np.float32)).astype(np.float32)

    sx, fx = coeffs(dw, scale_x)
    lo = sx < 0
    fx, sx = np.where(lo, np.float32(0), fx), np.where(lo, 0, sx)
    hi = sx >= W - 1
    fx, sx = np.where(hi, np.float32(0), fx), np.where(hi, W - 1, sx)
    a0 = np.rint((np.float32(1) - fx) * np.float32(2048)).astype(np.int32)
    a1 = np.rint(fx * np.float32(2048)).astype(np.int32)
    sx1 = np.minimum(sx + 1, W - 1)
    sy, fy = coeffs(dh, scale_y)
    b0 = np.rint((np.float32(1) - fy) * np.float32(2048)).astype(np.int32)
    b1 = np.rint(fy * np.float32(2048)).astype(np.int32)
    y0, y1 = _clip(sy, 0, H), _clip(sy + 1, 0, H)
    shape = (slice(None), slice(None)) + (None,) * (src.ndim - 2)
    hrow = s[:, sx] * a0[None][shape] + s[:, sx1] * a1[None][shape]
    vshape = (slice(None), None) + (None,) * (src.ndim - 2)
    out = (((b0[vshape] * (hrow[y0] >> 4)) >> 16) + ((b1[vshape] * (hrow[y1] >> 4)) >> 16) + 2) >> 2
    return np.clip(out, 0, 255).astype(np.uint8)


# ---------------------------------------------------------------------------------------------------- interpolation tables
def bilinear_table():
    """BilinearTab_i of initInterTab2D(INTER_LINEAR, fixpt): [32 fy][32 fx][4] 15-bit weights.  Every entry is exact except (0, 0),
    where 1.0 * 2^15 saturates to 32767 and the sum correction lands on the last tap."""
    f = np.arange(32)
    tab = np.stack([np.outer(32 - f, 32 - f), np.outer(32 - f, f), np.outer(f, 32 - f), np.outer(f, f)], -1) * 32
    tab[0, 0] = [32767, 0, 0, 1]
    return tab.astype(np.int32)


def cubic_table():
    """BicubicTab_i (A = -0.75): [32 fy][32 fx][16]; each 4x4 kernel is forced to sum to 2^15 by adjusting the largest (or
    smallest) of its four central taps (initInterTab2D)."""
    A = np.float32(-0.75)
    t1 = np.zeros((32, 4), np.float32)
    for i in range(32):
        x = np.float32(i / 32.0)
        c0 = ((A * (x + 1) - 5 * A) * (x + 1) + 8 * A) * (x + 1) - 4 * A
        c1 = ((A + 2) * x - (A + 3)) * x * x + 1
        c2 = ((A + 2) * (1 - x) - (A + 3)) * (1 - x) * (1 - x) + 1
        t1[i] = [c0, c1, c2, np.float32(1) - c0 - c1 - c2]
    tab = np.zeros((32, 32, 16), np.int32)
    for i in range(32):
        for j in range(32):
            v = (t1[i][:, None] * t1[j][None, :]).astype(np.float32).reshape(-1)
            it = np.clip(np.rint(v * np.float32(32768)), -32768, 32767).astype(np.int32)
            diff = int(it.sum()) - 32768
            if diff:
                big = small = 2 * 4 + 2
                for k1 in (2, 3):
                    for k2 in (2, 3):
                        k = k1 * 4 + k2
                        if it[k] < it[small]:
                            small = k
                        elif it[k] > it[big]:
                            big = k
                it[big if diff < 0 else small] -= diff
            tab[i, j] = it
    return tab


_BTAB = _CTAB = None


# ---------------------------------------------------------------------------------------------------- warpPerspective
def warp_perspective_u8(src, M, dsize=None):
    """cv2.warpPerspective(src, M, (w, h), borderMode=BORDER_REPLICATE) (INTER_LINEAR) for uint8 [H, W, C]; output size = input size."""
    global _BTAB
    if _BTAB is None:
        _BTAB = bilinear_table()
    H, W = src.shape[:2]
    m = cv2.invert(np.asarray(M, np.float64))[1].reshape(-1)  # warpPerspective inverts the matrix with cv::invert
    bw = min(1024 // min(16, H), W)  # the invoker walks 16-row x 64-column blocks: positions are X0(block origin) + M0 * x1
    xs = np.arange(W)
    xb = (xs // bw) * bw
    x0, x1 = xb.astype(np.float64)[None, :], (xs - xb).astype(np.float64)[None, :]
    y = np.arange(H, dtype=np.float64)[:, None]
    X0, Y0, W0 = m[0] * x0 + m[1] * y + m[2], m[3] * x0 + m[4] * y + m[5], m[6] * x0 + m[7] * y + m[8]
    Wd = W0 + m[6] * x1
    with np.errstate(divide="ignore", invalid="ignore"):
        Wd = np.where(Wd != 0, 32.0 / Wd, 0.0)
    X = np.rint(np.clip((X0 + m[0] * x1) * Wd, -2147483648.0, 2147483647.0)).astype(np.int64)
    Y = np.rint(np.clip((Y0 + m[3] * x1) * Wd, -2147483648.0, 2147483647.0)).astype(np.int64)
    sx, sy = np.clip(X >> 5, -32768, 32767), np.clip(Y >> 5, -32768, 32767)
    w = _BTAB[Y & 31, X & 31]
    s = src.astype(np.int64).reshape(H, W, -1)
    sx0, sx1, sy0, sy1 = _clip(sx, 0, W), _clip(sx + 1, 0, W), _clip(sy, 0, H), _clip(sy + 1, 0, H)
    t = s[sy0, sx0] * w[..., 0:1] + s[sy0, sx1] * w[..., 1:2] + s[sy1, sx0] * w[..., 2:3] + s[sy1, sx1] * w[..., 3:4]
    return np.clip((t + (1 << 14)) >> 15, 0, 255).astype(np.uint8).reshape(src.shape)


# ---------------------------------------------------------------------------------------------------- warpAffine (cubic)
def invert_affine(M):
    """The inversion cv::warpAffine applies to a forward 2x3 matrix (imgwarp.cpp), same operation order."""
    M = np.asarray(M, np.float64)
    D = M[0, 0] * M[1, 1] - M[0, 1] * M[1, 0]
    D = 1.0 / D if D != 0 else 0.0
    m = np.zeros(6)
    m[0], m[1], m[3], m[4] = M[1, 1] * D, M[0, 1] * (-D), M[1, 0] * (-D), M[0, 0] * D
    m[2] = -m[0] * M[0, 2] - m[1] * M[1, 2]
    m[5] = -m[3] * M[0, 2] - m[4] * M[1, 2]
    return m


def warp_affine_cubic_u8(src, M):
    """cv2.warpAffine(src, M, (w, h), flags=2 (INTER_CUBIC), borderMode=BORDER_REPLICATE) for uint8 [H, W, C]."""
    global _CTAB
    if _CTAB is None:
        _CTAB = cubic_table()
    H, W = src.shape[:2]
    m = invert_affine(M)
    xs, ys = np.arange(W), np.arange(H)
    adelta, bdelta = np.rint(m[0] * xs * 1024).astype(np.int64), np.rint(m[3] * xs * 1024).astype(np.int64)
    X0 = np.rint((m[1] * ys + m[2]) * 1024).astype(np.int64) + 16
    Y0 = np.rint((m[4] * ys + m[5]) * 1024).astype(np.int64) + 16
    X, Y = (X0[:, None] + adelta[None, :]) >> 5, (Y0[:, None] + bdelta[None, :]) >> 5
    sx, sy = np.clip(X >> 5, -32768, 32767) - 1, np.clip(Y >> 5, -32768, 32767) - 1
    w = _CTAB[Y & 31, X & 31]
    s = src.astype(np.int64).reshape(H, W, -1)
    t = np.zeros(s.shape, np.int64)
    for k1 in range(4):
        yy = _clip(sy + k1, 0, H)
        for k2 in range(4):
            t += s[yy, _clip(sx + k2, 0, W)] * w[..., k1 * 4 + k2][..., None]
    return np.clip((t + (1 << 14)) >> 15, 0, 255).astype(np.uint8).reshape(src.shape)


# ---------------------------------------------------------------------------------------------------- the tracker's crop
def crop_geometry(pos, original_sz):
    """Window of SiameseTracker.get_subwindow (hdn/tracker/base_tracker.py:76-92) in FRAME coordinates: (x0, y0, n) = top-left
    pixel (may lie outside the frame: the reference pads with the channel means) and side in pixels."""
    half = (original_sz - 1) / 2
    x0 = np.floor(pos[0] - half + 0.5)
    y0 = np.floor(pos[1] - half + 0.5)
    return int(x0), int(y0), int(x0 + original_sz - 1 + 1) - int(x0)


def crop_resize(frame, pos, model_sz, original_sz, avg_chans):
    """get_subwindow without the tensor wrapping: mean-padded window -> cv2.resize -> float32 [1, 3, S, S] (restated on the
    functions above; the compat tracker's crop_window is pinned to the reference's own output in tests/test_compat_host.py)."""
    x0, y0, n = crop_geometry(pos, original_sz)
    H, W = frame.shape[:2]
    fill = np.asarray(avg_chans).astype(np.uint8)
    ys, xs = np.arange(y0, y0 + n), np.arange(x0, x0 + n)
    inside = ((ys >= 0) & (ys < H))[:, None] & ((xs >= 0) & (xs < W))[None, :]
    patch = np.where(inside[..., None], frame[np.clip(ys, 0, H - 1)[:, None], np.clip(xs, 0, W - 1)[None, :]], fill[None, None, :])
    if n != model_sz:
        patch = resize_linear_u8(patch, (model_sz, model_sz))
    return np.ascontiguousarray(patch.transpose(2, 0, 1)[None].astype(np.float32))


def gray_normalise(crop):
    """get_search_info / get_template_info (Oneline_DLTv1/tools/get_img_info.py:42-70): per-channel normalisation in float64, mean
    over the channels -> [1, S, S] float64."""
    mean = np.reshape(np.array([118.93, 113.97, 102.60]), (1, 1, 3))
    std = np.reshape(np.array([69.85, 68.81, 72.45]), (1, 1, 3))
    v = (np.asarray(crop)[0].transpose(1, 2, 0) - mean) / std
    return np.transpose(np.mean(v, axis=2, keepdims=True), [2, 0, 1])
