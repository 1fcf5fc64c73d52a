"""Synthetic LPBA-like volume pairs.

The reference's preprocessing (`/makePklDataset.py:19-22,76`) yields fp32 volumes min-max
scaled to [0, 1] with a skull-stripped (exactly zero) background.  There is no dataset on the
box, so benchmarks and tests use this seeded generator: a smoothed random field inside an
ellipsoid mask for the fixed image, and the same field pushed through a smooth random +-3
voxel displacement plus sigma=0.01 noise for the moving image.  Generated on the CPU (data
preparation, not part of the timed path) so that every box builds identical inputs.
"""
from __future__ import annotations

from typing import Sequence, Tuple

import torch
import torch.nn.functional as F


def _smooth(x: torch.Tensor, passes: int = 3) -> torch.Tensor:
    for _ in range(passes):
        x = F.avg_pool3d(x, kernel_size=5, stride=1, padding=2)
    return x


def _minmax(x: torch.Tensor) -> torch.Tensor:
    lo, hi = x.amin(dim=(1, 2, 3, 4), keepdim=True), x.amax(dim=(1, 2, 3, 4), keepdim=True)
    return (x - lo) / (hi - lo).clamp_min(1e-12)


def make_pair(shape: Sequence[int], batch: int = 1, seed: int = 24,
              max_disp: float = 3.0) -> Tuple[torch.Tensor, torch.Tensor]:
    """Returns (moving, fixed), each [batch, 1, D, H, W] fp32 in [0, 1] on the CPU."""
    D, H, W = (int(s) for s in shape)
    g = torch.Generator().manual_seed(seed)
    fixed = _minmax(_smooth(torch.rand(batch, 1, D, H, W, generator=g)))
    zz, yy, xx = torch.meshgrid(torch.linspace(-1, 1, D), torch.linspace(-1, 1, H),
                                torch.linspace(-1, 1, W), indexing="ij")
    mask = ((zz / 0.9) ** 2 + (yy / 0.9) ** 2 + (xx / 0.9) ** 2 <= 1.0).to(torch.float32)
    fixed = fixed * mask

    cs = [max(2, s // 16) for s in (D, H, W)]
    coarse = (torch.rand(batch, 3, *cs, generator=g) * 2 - 1) * max_disp
    disp = F.interpolate(coarse, size=(D, H, W), mode="trilinear", align_corners=True)
    base = torch.stack([zz, yy, xx], 0).unsqueeze(0)                       # normalised coords
    scale = torch.tensor([2.0 / max(D - 1, 1), 2.0 / max(H - 1, 1), 2.0 / max(W - 1, 1)]).view(1, 3, 1, 1, 1)
    grid = (base + disp * scale).permute(0, 2, 3, 4, 1)[..., [2, 1, 0]]
    moving = F.grid_sample(fixed, grid, mode="bilinear", padding_mode="zeros", align_corners=True)
    moving = (moving + 0.01 * torch.randn(batch, 1, D, H, W, generator=g)).clamp_(0, 1) * mask
    return moving.contiguous(), fixed.contiguous()


def randomize_weights(model: torch.nn.Module, seed: int = 1234) -> None:
    """Non-degenerate decoder weights for benchmarks: the reference's default
    init (proj.weight ~ N(0, 1e-5), rpb = 0) makes q, k ~ 0 and the flow ~ 5e-3 voxels, which would
    exercise nothing.  proj.weight ~ N(0, 1/Cin), LayerNorm gain ~ U(0.25, 0.75), bias ~ N(0, 0.05),
    rpb ~ N(0, 0.5); conv weights keep their default initialisation."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in sorted(model.named_parameters()):
            if name.endswith("rpb"):
                p.copy_(torch.randn(p.shape, generator=g) * 0.5)
            elif name.endswith("proj.weight"):
                p.copy_(torch.randn(p.shape, generator=g) / p.shape[1] ** 0.5)
            elif name.endswith("norm.weight"):
                p.copy_(torch.rand(p.shape, generator=g) * 0.5 + 0.25)
            elif name.endswith("norm.bias"):
                p.copy_(torch.randn(p.shape, generator=g) * 0.05)
