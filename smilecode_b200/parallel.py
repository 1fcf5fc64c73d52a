"""Multi-GPU plumbing for the ModeT path: one process per GPU.

* Inference: volume pairs are independent (InstanceNorm / LayerNorm are per sample, the reference has no
  cross-sample op), so the batch of pairs is sharded across ranks with NO data-path collective --
  `shard_range` is the whole story, `bench.py --gpus N` times it.
* Training: weights are replicated (1,029,670 fp32 = 4.12 MB) and the only exchange step is one
  all-reduce of the flat gradient per step -- `FlatGradAllReduce`.  The reference itself is single-GPU
  batch-1 (ModeT/train.py:43,183-189); this is added capability, backend NCCL on the GPUs (NVLink 5 /
  NVSwitch; at 4 MB the collective is latency-bound, so it is ONE bucket) and gloo in the CPU tests.
"""
from __future__ import annotations

from typing import Iterable, List, Tuple

import torch
import torch.distributed as dist


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [begin, end) share of `n_items` for `rank`; shares differ by at most one item."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of size {world}")
    base, extra = divmod(n_items, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def flat_layout(params: Iterable[torch.nn.Parameter], align: int = 4) -> Tuple[List[int], int]:
    """Offsets (in elements) of `params` packed into one flat buffer with every parameter starting on a multiple of `align`
    elements (4 fp32 = 16 bytes: the C ABI's pointer alignment -- an unaligned view would be cloned on every kernel call),
    and the padded total.  The parameter buffer of `train.Trainer` and the gradient bucket below share this layout, so the
    fused Adam update runs over one contiguous range; the padding stays zero."""
    offs, off = [], 0
    for p in params:
        offs.append(off)
        off += (p.numel() + align - 1) // align * align
    return offs, off


class FlatGradAllReduce:
    """Averages the gradients of `params` across ranks with ONE all-reduce over a flat fp32 bucket.

    The bucket is allocated once; `.grad` tensors are re-pointed into it (views), so the backward pass writes
    straight into the communication buffer and no pack/unpack copies are needed after the first step."""

    def __init__(self, params: Iterable[torch.nn.Parameter], group=None):
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("no trainable parameters")
        dev, dt = self.params[0].device, torch.float32
        offs, self.numel = flat_layout(self.params)
        self.bucket = torch.zeros(self.numel, device=dev, dtype=dt)
        self.group = group
        for p, off in zip(self.params, offs):
            if p.dtype != dt or p.device != dev:
                raise ValueError("FlatGradAllReduce expects fp32 parameters on one device")
            view = self.bucket[off:off + p.numel()].view_as(p)
            if p.grad is not None:
                view.copy_(p.grad)
            p.grad = view

    def zero_(self) -> None:
        self.bucket.zero_()

    def allreduce_mean_(self, async_op: bool = False):
        """Sum across ranks and divide by the world size (in place).  Returns the work handle if async."""
        world = dist.get_world_size(self.group)
        if world == 1:
            return None
        self.bucket.mul_(1.0 / world)   # pre-scale: the reduction is then a plain SUM (NVLS-friendly)
        work = dist.all_reduce(self.bucket, op=dist.ReduceOp.SUM, group=self.group, async_op=async_op)
        return work if async_op else None
