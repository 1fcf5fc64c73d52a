"""CPU restatement of the evaluation helpers of ModeT/utils.py (test infrastructure: imported by tests/ only, never
by smilecode_b200/).  numpy / torch-CPU, each function citing the reference lines it follows.  Pinned against
fixtures generated from the reference's own utils.py (oracle/make_golden_metrics.py -> tests/golden/metrics_*.npz).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

VOI_LBLS = list(range(1, 55))   # ModeT/utils.py:87-91


def warp_nearest(src: torch.Tensor, flow: torch.Tensor) -> torch.Tensor:
    """SpatialTransformer(mode='nearest').forward (ModeT/utils.py:49-72); register_model (74-83) wraps it."""
    shape = flow.shape[2:]
    grids = torch.meshgrid([torch.arange(0, s) for s in shape], indexing="ij")           # utils.py:39-42
    grid = torch.stack(grids).unsqueeze(0).to(torch.float32)
    new_locs = grid + flow                                                                # utils.py:55
    for i in range(len(shape)):                                                           # utils.py:59-61
        new_locs[:, i, ...] = 2 * (new_locs[:, i, ...] / (shape[i] - 1) - 0.5)
    new_locs = new_locs.permute(0, 2, 3, 4, 1)[..., [2, 1, 0]]                            # utils.py:68-70
    return F.grid_sample(src, new_locs, align_corners=True, mode="nearest")               # utils.py:72


def dice_val_VOI(pred: np.ndarray, true: np.ndarray, labels=VOI_LBLS) -> float:
    """ModeT/utils.py:86-106 on integer label volumes [D,H,W]."""
    dscs = np.zeros((len(labels), 1))
    for idx, i in enumerate(labels):
        pred_i, true_i = pred == i, true == i
        intersection = np.sum(pred_i * true_i)
        union = np.sum(pred_i) + np.sum(true_i)
        dscs[idx] = (2.0 * intersection) / (union + 1e-5)
    return float(np.mean(dscs))


def jacobian_determinant_vxm(disp: np.ndarray) -> np.ndarray:
    """ModeT/utils.py:108-150 for disp [3,D,H,W].  `nd.volsize2ndgrid` (pystrum, un-vendored and unpinned in the
    reference) is np.meshgrid(*[np.arange(e) for e in volshape], indexing='ij')."""
    disp = disp.transpose(1, 2, 3, 0)
    volshape = disp.shape[:-1]
    grid = np.stack(np.meshgrid(*[np.arange(e) for e in volshape], indexing="ij"), len(volshape))
    J = np.gradient(disp + grid)
    dx, dy, dz = J[0], J[1], J[2]
    jdet0 = dx[..., 0] * (dy[..., 1] * dz[..., 2] - dy[..., 2] * dz[..., 1])
    jdet1 = dx[..., 1] * (dy[..., 0] * dz[..., 2] - dy[..., 2] * dz[..., 0])
    jdet2 = dx[..., 2] * (dy[..., 0] * dz[..., 1] - dy[..., 1] * dz[..., 0])
    return jdet0 - jdet1 + jdet2
