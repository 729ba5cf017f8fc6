"""Input contract of the hot path: the feature-column dictionaries every model
constructor consumes, Criteo-shaped synthetic batches in the reference's array format, and a
prefetching host -> device batch feeder.

The reference builds `feature_columns = [[denseFeature...], [sparseFeature...]]` and feeds
`[float32 (N, 13), int32 (N, 26)]` + int32 labels (src/ctr/utils/data_process.py:13-30,80-91);
the matching models use the `feat_len` / `maxlen` variants (src/match/utils/feature_util.py:1-29).
Its pandas/sklearn loaders are host code and out of scope; only their output format is kept.
"""
from __future__ import annotations

from typing import Iterable, Iterator, List, Optional, Sequence, Tuple

import numpy as np
import torch

# Criteo-Kaggle sparse cardinalities (C1..C26), SURVEY.md §8d
CRITEO_KAGGLE_ROWS = (1460, 583, 10131227, 2202608, 305, 24, 12517, 633, 3, 93145, 5683, 8351593,
                      3194, 27, 14992, 5461306, 10, 5652, 2173, 4, 7046547, 18, 15, 286181, 105,
                      142572)
N_DENSE, N_SPARSE = 13, 26


def sparseFeature(feat, feat_num, embed_dim=4, feat_len=None):
    """src/ctr/utils/data_process.py:13-21 (and, with feat_len, src/match/utils/feature_util.py:1-10)."""
    d = {"feat": feat, "feat_num": feat_num, "embed_dim": embed_dim}
    if feat_len is not None:
        d["feat_len"] = feat_len
    return d


def denseFeature(feat):
    """src/ctr/utils/data_process.py:24-30."""
    return {"feat": feat}


def varLenSparseFeat(feat, feat_num, maxlen, embed_dim=4):
    """src/match/utils/feature_util.py:21-29."""
    return {"feat": feat, "feat_num": feat_num, "maxlen": maxlen, "embed_dim": embed_dim}


def criteo_feature_columns(embed_dim: int = 8, rows: Optional[Sequence[int]] = None,
                           row_cap: Optional[int] = None):
    """The `feature_columns` list create_criteo_dataset returns (data_process.py:80-82): dense
    I1..I13, sparse C1..C26 with `feat_num` = cardinality (optionally capped)."""
    rows = list(CRITEO_KAGGLE_ROWS if rows is None else rows)
    if row_cap is not None:
        rows = [min(r, row_cap) for r in rows]
    return [[denseFeature(f"I{i}") for i in range(1, N_DENSE + 1)],
            [sparseFeature(f"C{i}", r, embed_dim=embed_dim) for i, r in enumerate(rows, 1)]]


def synthetic_criteo_batch(rng: np.random.Generator, batch: int, rows: Sequence[int],
                           ids: str = "uniform", pos_rate: float = 0.25):
    """One batch in the reference's array format: dense float32 (B, 13) in [0, 1) (the loader
    min-max scales, data_process.py:76-78), sparse int32 (B, 26) label-encoded ids in
    [0, feat_num), labels float32 (B, 1).  ids: 'uniform' (worst case for HBM) or 'zipf'
    (Zipf(1.05) mod N_t: realistic skew)."""
    dense = rng.random((batch, N_DENSE), dtype=np.float32)
    if ids == "uniform":
        sparse = np.stack([rng.integers(0, r, batch, dtype=np.int64) for r in rows], 1)
    elif ids == "zipf":
        sparse = np.stack([(rng.zipf(1.05, batch) - 1) % r for r in rows], 1)
    else:
        raise ValueError(ids)
    y = (rng.random((batch, 1)) < pos_rate).astype(np.float32)
    return dense, sparse.astype(np.int32), y


class DeviceFeeder:
    """Feeds batches to the GPU ahead of the consumer: batch i+depth-1 is copied from PINNED host
    memory on a dedicated copy stream while step i computes, so the H2D transfer (10 MB per
    65 536 Criteo samples) leaves the critical path without leaving the measured step.

        for dense, sparse, y in DeviceFeeder(host_batches):
            trainer.step(dense, sparse, y)

    Device staging is a ring of `depth` preallocated slots (no allocation per step: a fresh
    block per batch on the copy stream's pool costs a synchronising cudaMalloc every few steps,
    measured +1.4 ms/step).  A slot is refilled only after the consumer stream has passed the
    step that read it (event), so the yielded tensors stay valid until the next batch is
    requested — consume them inside the loop body.  Host tensors that are not pinned are pinned
    once (a page-locked staging copy)."""

    def __init__(self, batches: Iterable[Tuple[torch.Tensor, ...]], device=None, depth: int = 2):
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else device
        self.it = iter(batches)
        self.depth = max(1, depth)
        self.stream = torch.cuda.Stream(device=self.device)
        self.slots = [{"dev": None, "ready": torch.cuda.Event(), "free": None, "host": None}
                      for _ in range(self.depth)]
        self.h2d_bytes = 0

    def _fill(self, slot) -> bool:
        try:
            host = next(self.it)
        except StopIteration:
            return False
        host = tuple(t if t.is_pinned() else t.pin_memory() for t in host)
        dev = slot["dev"]
        if dev is None or any(d.shape != h.shape or d.dtype != h.dtype for d, h in zip(dev, host)):
            dev = tuple(torch.empty(h.shape, dtype=h.dtype, device=self.device) for h in host)
            slot["dev"] = dev
        if slot["free"] is not None:
            self.stream.wait_event(slot["free"])       # the step that read this slot is done
        with torch.cuda.stream(self.stream):
            for d, h in zip(dev, host):
                d.copy_(h, non_blocking=True)
            slot["ready"].record(self.stream)
        slot["host"] = host                            # keep the pinned source alive until copied
        self.h2d_bytes += sum(t.numel() * t.element_size() for t in host)
        return True

    def __iter__(self) -> Iterator[Tuple[torch.Tensor, ...]]:
        queue = self._queue = []
        for slot in self.slots:
            if not self._fill(slot):
                break
            queue.append(slot)
        while queue:
            slot = queue.pop(0)
            cur = torch.cuda.current_stream(self.device)
            cur.wait_event(slot["ready"])
            yield slot["dev"]
            # resumed: the consumer has enqueued its work on this batch
            slot["free"] = torch.cuda.Event()
            slot["free"].record(torch.cuda.current_stream(self.device))
            if self._fill(slot):
                queue.append(slot)

    def peek_next(self) -> Optional[Tuple[torch.Tensor, ...]]:
        """Inside the loop body: the device tensors of the NEXT batch (already on their way over
        the copy stream), or None after the last one.  The current stream is made to wait for
        that copy, so the tensors may be read by work enqueued from now on — e.g. to start the
        next batch's embedding exchange behind the current step (ShardedDLRMTrainer.step)."""
        queue = getattr(self, "_queue", None)
        if not queue:
            return None
        slot = queue[0]
        torch.cuda.current_stream(self.device).wait_event(slot["ready"])
        return slot["dev"]
