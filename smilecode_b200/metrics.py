"""Evaluation path of ModeT/infer.py on the GPU: the reference's `utils.register_model`, `utils.dice_val_VOI` and
`utils.jacobian_determinant_vxm` (ModeT/utils.py:74-150) with the same names, arguments and results, but without
the 59 MB device->host copy of the flow and the numpy passes over it.

    reg_model = metrics.register_model(img_size, 'nearest')           # infer.py:66
    def_out = reg_model([x_seg.float(), flow])                        # infer.py:87
    dsc = metrics.dice_val_VOI(def_out.long(), y_seg.long())          # infer.py:91
    frac = metrics.nonpositive_jacobian_fraction(flow)                # infer.py:89-90 in one call
"""
from __future__ import annotations

from typing import Optional, Sequence

import numpy as np
import torch
from torch import Tensor, nn

from . import ops
from ._lib import SmileError, call

# ModeT/utils.py:87-91 (LPBA40: 54 labels)
VOI_LBLS = tuple(range(1, 55))


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _f32(x: Tensor, name: str) -> Tensor:
    if not x.is_cuda:
        raise SmileError(f"{name} must be a CUDA tensor (smilecode_b200 has no CPU path)")
    return x.detach().to(torch.float32).contiguous()


def warp3d_nearest(src: Tensor, flow: Tensor) -> Tensor:
    """SpatialTransformer(mode='nearest').forward (ModeT/utils.py:49-72)."""
    src, flow = _f32(src, "src"), _f32(flow, "flow")
    if src.dim() != 5 or flow.dim() != 5:
        raise SmileError("src must be [B,C,D,H,W] and flow [B,3,D,H,W]")
    B, C, D, H, W = src.shape
    if tuple(flow.shape) != (B, 3, D, H, W):
        raise SmileError(f"flow must be [{B},3,{D},{H},{W}], got {tuple(flow.shape)}")
    out = torch.empty_like(src)
    call("smile_warp3d_nearest_fwd", src.data_ptr(), flow.data_ptr(), out.data_ptr(), B, C, D, H, W, _stream(),
         label=f"[c{C} {D}x{H}x{W}]")
    return out


class register_model(nn.Module):
    """utils.register_model (ModeT/utils.py:74-83): forward([img, flow]) -> warped img."""

    def __init__(self, img_size=(64, 256, 256), mode: str = "bilinear"):
        super().__init__()
        if mode not in ("bilinear", "nearest"):
            raise SmileError(f"register_model: mode {mode!r} not supported")
        self.img_size, self.mode = tuple(img_size), mode

    def forward(self, x):
        img, flow = x[0].cuda(), x[1].cuda()
        if self.mode == "nearest":
            return warp3d_nearest(img, flow)
        return ops.warp3d(_f32(img, "img"), _f32(flow, "flow"))


def dice_counts(y_pred: Tensor, y_true: Tensor, labels: Sequence[int] = VOI_LBLS) -> Tensor:
    """[len(labels), 3] int64 on the device: |pred==l & true==l|, |pred==l|, |true==l| of sample 0, channel 0."""
    pred, true = _f32(y_pred, "y_pred")[0, 0].contiguous(), _f32(y_true, "y_true")[0, 0].contiguous()
    if pred.shape != true.shape:
        raise SmileError(f"dice: shapes differ: {tuple(pred.shape)} vs {tuple(true.shape)}")
    lab = torch.tensor(list(labels), dtype=torch.int32, device=pred.device)
    counts = torch.empty((len(labels), 3), dtype=torch.int64, device=pred.device)
    call("smile_dice_counts_fwd", pred.data_ptr(), true.data_ptr(), lab.data_ptr(), len(labels), counts.data_ptr(),
         pred.numel(), _stream(), label=f"[{len(labels)} labels]")
    return counts


def dice_val_VOI(y_pred: Tensor, y_true: Tensor, labels: Sequence[int] = VOI_LBLS) -> float:
    """utils.dice_val_VOI (ModeT/utils.py:86-106): mean over the labels of 2|A&B| / (|A| + |B| + 1e-5), in float64."""
    c = dice_counts(y_pred, y_true, labels).cpu().numpy()    # 54 x 3 integers cross the bus, not two volumes
    dscs = np.zeros((len(labels), 1))                        # same float64 arithmetic and np.mean as utils.py:97-106
    dscs[:, 0] = (2.0 * c[:, 0]) / (c[:, 1] + c[:, 2] + 1e-5)
    return float(np.mean(dscs))


def jacobian_determinant_vxm(disp: Tensor, want_det: bool = True):
    """utils.jacobian_determinant_vxm (ModeT/utils.py:108-150) for disp [3,D,H,W] (or [1,3,D,H,W]): returns
    (det float64 [D,H,W] on the device or None, number of voxels with det <= 0)."""
    flow = _f32(disp, "disp")
    if flow.dim() == 5:
        if flow.shape[0] != 1:
            raise SmileError("jacobian_determinant_vxm takes one displacement field")
        flow = flow[0]
    if flow.dim() != 4 or flow.shape[0] != 3:
        raise SmileError(f"disp must be [3,D,H,W], got {tuple(flow.shape)}")
    flow = flow.contiguous()
    _, D, H, W = flow.shape
    det: Optional[Tensor] = torch.empty((D, H, W), dtype=torch.float64, device=flow.device) if want_det else None
    nonpos = torch.empty(1, dtype=torch.int64, device=flow.device)
    call("smile_jacdet_fwd", flow.data_ptr(), det.data_ptr() if det is not None else None, nonpos.data_ptr(), D, H, W,
         _stream(), label=f"[{D}x{H}x{W}]")
    return det, nonpos


def nonpositive_jacobian_fraction(flow: Tensor) -> float:
    """np.sum(jac_det <= 0) / np.prod(shape) of infer.py:89-90 with 8 bytes crossing the bus."""
    _, nonpos = jacobian_determinant_vxm(flow, want_det=False)
    n = flow.shape[-1] * flow.shape[-2] * flow.shape[-3]
    return float(nonpos.item()) / float(n)
