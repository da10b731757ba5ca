"""Mirror of homo_estimator/Deep_homography/Oneline_DLTv1/utils.py (lines 7-274: DLT_solve, transformer, transform),
bound to the sm_100a kernels.  The tensorboard / nvidia-smi / psutil helpers of lines 277-376 are not part of
the tracking path and are not mirrored."""
from hdn_b200.ops import DLT_solve, dlt_warp, homo_warp, transform  # noqa: F401


def transformer(U, theta, out_size, **kwargs):
    """utils.py:70-254.  `theta` is the already conjugated 3x3 (M^-1 H M); returns ([B,H,W,C] output, condition).
    Implemented by running the warp kernel with M = M^-1 = I."""
    import torch

    eye = [1.0, 0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0, 1.0]
    if tuple(out_size) != tuple(U.shape[2:]):
        raise RuntimeError("transformer: out_size must equal the input size (the only form the reference calls)")
    out = homo_warp(U, theta.reshape(-1, 3, 3), eye, eye)
    # `condition` (utils.py:240) counts |t| > 1e-7 after the epsilon fix-up; no caller reads it (utils.py:265 discards it)
    return out.permute(0, 2, 3, 1), torch.tensor(float(U.shape[0] * U.shape[2] * U.shape[3]), device=U.device)
