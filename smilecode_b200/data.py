"""Data side of the hot path: the reference's pickle pair datasets (ModeT/data/datasets.py) and a
prefetch ring that keeps pinned host buffers and a copy stream ahead of the registration kernels.

A sample file is `pickle.dump((img float32 [D,H,W] in [0,1], seg uint16 [D,H,W]))` (ModeT/data/datasets.py:8-10,
makePklDataset.py).  `PklPairDataset` reproduces LPBABrainDatasetS2S / LPBABrainInferDatasetS2S and their `Half`
variants (datasets.py:12-185) together with the transforms train.py:92-95 / infer.py:68-70 compose around them
(`Seg_norm` + `NumpyType`, data/trans.py:27-55): same index -> (moving, fixed) pairing, same dtypes and shapes.
"""
from __future__ import annotations

import pickle
import queue
import threading
from typing import Iterable, Iterator, List, Sequence, Tuple

import numpy as np
import torch

# data/trans.py:30-32 -- LPBA40 label values in file order; Seg_norm maps table[i] -> i, everything else -> 0
SEG_TABLE = np.array([0, 21, 22, 23, 24, 25, 26, 27, 28, 29, 30, 31, 32, 33, 34, 41, 42, 43, 44, 45, 46, 47, 48, 49, 50, 61,
                      62, 63, 64, 65, 66, 67, 68, 81, 82, 83, 84, 85, 86, 87, 88, 89, 90, 91, 92, 101, 102, 121, 122, 161,
                      162, 163, 164, 165, 166])


def pkload(fname: str):
    """datasets.py:8-10."""
    with open(fname, "rb") as f:
        return pickle.load(f)


def pair_index(index: int, n: int) -> Tuple[int, int]:
    """index -> (moving subject, fixed subject) over all ordered pairs of n subjects (datasets.py:25-27)."""
    x_index = index // (n - 1)
    s = index % (n - 1)
    y_index = s + 1 if s >= x_index else s
    return x_index, y_index


def seg_norm(seg: np.ndarray) -> np.ndarray:
    """trans.Seg_norm.tf for the label volume (data/trans.py:33-39) as one table lookup."""
    lut = np.zeros(int(max(int(seg.max(initial=0)), int(SEG_TABLE.max()))) + 1, dtype=seg.dtype)
    lut[SEG_TABLE] = np.arange(len(SEG_TABLE), dtype=seg.dtype)
    return lut[seg]


class PklPairDataset(torch.utils.data.Dataset):
    """LPBABrain[Half][Infer]DatasetS2S.  infer=False -> (x, y) float32 [1,D,H,W];
    infer=True -> (x, y, x_seg, y_seg) with segmentations normalised by SEG_TABLE and cast to int16."""

    def __init__(self, paths: Sequence[str], infer: bool = False, half: bool = False):
        self.paths, self.infer, self.half = list(paths), infer, half

    def __len__(self) -> int:
        return len(self.paths) * (len(self.paths) - 1)

    def _load(self, path: str):
        img, seg = pkload(path)
        if self.half:                                     # datasets.py:105-106
            img, seg = img[::2, ::2, ::2], seg[::2, ::2, ::2]
        return img, seg

    def __getitem__(self, index: int):
        xi, yi = pair_index(index, len(self.paths))
        x, x_seg = self._load(self.paths[xi])
        y, y_seg = self._load(self.paths[yi])
        x = np.ascontiguousarray(x[None, ...].astype(np.float32))
        y = np.ascontiguousarray(y[None, ...].astype(np.float32))
        if not self.infer:
            return torch.from_numpy(x), torch.from_numpy(y)
        x_seg = np.ascontiguousarray(seg_norm(x_seg[None, ...]).astype(np.int16))
        y_seg = np.ascontiguousarray(seg_norm(y_seg[None, ...]).astype(np.int16))
        return torch.from_numpy(x), torch.from_numpy(y), torch.from_numpy(x_seg), torch.from_numpy(y_seg)


class PrefetchRing:
    """Iterates a dataset `depth` samples ahead.  With a device: a fixed RING of `depth + workers` pinned host slots is
    allocated once (sized by the first sample); worker threads unpickle and copy into a free slot, a copy stream uploads
    the slot, and the slot returns to the free list guarded by the CUDA event of its upload -- no cudaHostAlloc per sample
    (round-1 `t.pin_memory()` per sample allocated and copied 2 x 19.7 MB of fresh pinned memory per pair).  The consumer
    receives device tensors whose upload has been ordered before its current stream.
    With device=None it only prefetches on the host (usable without a GPU; this is what the CPU tests exercise)."""

    def __init__(self, dataset, indices: Iterable[int], depth: int = 3, workers: int = 2, device=None):
        self.dataset, self.indices, self.depth, self.workers = dataset, list(indices), max(1, depth), max(1, workers)
        self.device = torch.device(device) if device is not None else None
        self._copy_stream = torch.cuda.Stream(self.device) if self.device is not None else None
        self.nslots = self.depth + self.workers
        self._slots: List[dict] = [{"tensors": None, "event": None} for _ in range(self.nslots)]
        self.pinned_allocations = 0          # how many times pinned memory was (re)allocated: nslots for a uniform dataset

    def _fill_slot(self, slot: dict, sample):
        if slot["event"] is not None:        # the upload that last read this slot must be complete
            slot["event"].synchronize()
            slot["event"] = None
        ts = slot["tensors"]
        if ts is None or len(ts) != len(sample) or any(a.shape != b.shape or a.dtype != b.dtype for a, b in zip(ts, sample)):
            ts = tuple(torch.empty(t.shape, dtype=t.dtype).pin_memory() for t in sample)
            slot["tensors"] = ts
            self.pinned_allocations += 1
        for dst, src in zip(ts, sample):
            dst.copy_(src)
        return ts

    def _produce(self, job_q: "queue.Queue", free_q: "queue.Queue", out: dict, cv: threading.Condition):
        while True:
            job = job_q.get()
            if job is None:
                return
            pos, idx = job
            slot_id = None
            try:
                sample = self.dataset[idx]
                if self.device is not None:
                    slot_id = free_q.get()
                    sample = self._fill_slot(self._slots[slot_id], sample)
            except BaseException as e:  # surfaced to the consumer, never swallowed
                if slot_id is not None:
                    free_q.put(slot_id)
                    slot_id = None
                sample = e
            with cv:
                out[pos] = (slot_id, sample)
                cv.notify_all()

    def __iter__(self) -> Iterator[Tuple[torch.Tensor, ...]]:
        job_q: "queue.Queue" = queue.Queue()
        free_q: "queue.Queue" = queue.Queue()
        for i in range(self.nslots):
            free_q.put(i)
        out: dict = {}
        cv = threading.Condition()
        threads = [threading.Thread(target=self._produce, args=(job_q, free_q, out, cv), daemon=True)
                   for _ in range(self.workers)]
        for t in threads:
            t.start()
        issued = 0
        try:
            for pos in range(len(self.indices)):
                while issued < len(self.indices) and issued < pos + self.depth:
                    job_q.put((issued, self.indices[issued]))
                    issued += 1
                with cv:
                    cv.wait_for(lambda: pos in out)
                    slot_id, sample = out.pop(pos)
                if isinstance(sample, BaseException):
                    raise sample
                if self.device is None:
                    yield sample
                    continue
                with torch.cuda.stream(self._copy_stream):
                    dev = tuple(t.to(self.device, non_blocking=True) for t in sample)
                done = torch.cuda.Event()
                done.record(self._copy_stream)
                self._slots[slot_id]["event"] = done     # the slot may be refilled once its upload has completed
                free_q.put(slot_id)
                torch.cuda.current_stream(self.device).wait_event(done)
                for t in dev:
                    t.record_stream(torch.cuda.current_stream(self.device))
                yield dev
        finally:
            for _ in threads:
                job_q.put(None)
