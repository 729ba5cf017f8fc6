"""Exact inner-product retrieval after the towers: the drop-in for the
`faiss.IndexFlatIP` + `index.search(user_embs, 10)` step of the matching scripts
(src/match/fm/train.py:71-75, src/match/dssm/dssm_train.py:74-78).  CUDA only."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib as L


def topk_ip(user_embs: torch.Tensor, item_embs: torch.Tensor, k: int = 10, check: bool = True):
    """-> (scores (B, k) fp32, indices (B, k) int64), best first, ties by lower index.
    Equal to ranking the fp64 inner products (rtf_topk_ip proves it per row; with check=True a
    result that could not be proven — more than 32 - k near-ties in a row — raises)."""
    L.require_cuda(user_embs, "topk_ip(user_embs)")
    L.require_cuda(item_embs, "topk_ip(item_embs)")
    u = user_embs.to(torch.float32).contiguous()
    it = item_embs.to(torch.float32).contiguous()
    B, D = u.shape
    N = it.shape[0]
    if it.shape[1] != D:
        raise ValueError("user and item embeddings must have the same width")
    idx = torch.empty((B, k), dtype=torch.int64, device=u.device)
    sc = torch.empty((B, k), dtype=torch.float32, device=u.device)
    flag = torch.zeros(1, dtype=torch.int32, device=u.device)
    nb = C.c_size_t(0)
    L.check(L.lib().rtf_topk_ip_workspace(B, N, D, k, C.byref(nb)), "rtf_topk_ip_workspace")
    ws = torch.empty(nb.value, dtype=torch.uint8, device=u.device)
    L.check(L.lib().rtf_topk_ip(u.data_ptr(), u.stride(0), B, it.data_ptr(), it.stride(0), N, D, k,
                                idx.data_ptr(), sc.data_ptr(), flag.data_ptr(), ws.data_ptr(),
                                ws.numel(), L.current_stream_ptr()), "rtf_topk_ip")
    if check and int(flag.item()) != 0:
        raise L.RtfError("topk_ip: a row has more than 32 - k near-tied candidates; exactness "
                         "against the fp64 ranking could not be proven")
    return sc, idx


class IndexFlatIP:
    """faiss.IndexFlatIP's three calls as the reference uses them: IndexFlatIP(d); add(x);
    search(q, k) -> (D, I).  Vectors stay on the GPU; `add` appends."""

    def __init__(self, d: int):
        self.d = d
        self._items = None

    @property
    def ntotal(self) -> int:
        return 0 if self._items is None else int(self._items.shape[0])

    def add(self, x):
        x = torch.as_tensor(x, dtype=torch.float32)
        if not x.is_cuda:
            x = x.cuda()
        if x.shape[1] != self.d:
            raise ValueError(f"expected vectors of width {self.d}")
        self._items = x if self._items is None else torch.cat([self._items, x], 0)

    def search(self, q, k: int):
        q = torch.as_tensor(q, dtype=torch.float32)
        if not q.is_cuda:
            q = q.cuda()
        # faiss leaves the order of exact ties unspecified; here: lower index first, no exception
        return topk_ip(q, self._items, k, check=False)
