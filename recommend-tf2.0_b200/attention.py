"""Host side of the fused short-sequence attention core (K7; also the core of K6).

q/k/v are the (B, L, H*hs) projection outputs; the kernel reads head h as a column slice, so
the reference's split_heads / merge transposes (src/match/layers/modules.py:63-74,130;
src/ctr/layers/modules.py:198-220,272-283) cost nothing."""
from __future__ import annotations

from typing import Optional

import torch

from . import _lib as L


def _mask2d(mask: Optional[torch.Tensor], B: int, n: int, what: str) -> Optional[torch.Tensor]:
    if mask is None:
        return None
    m = mask.reshape(B, n).to(torch.float32).contiguous()
    L.require_cuda(m, what)
    return m


class _AttnFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, q, k, v, H, scale, row_mask, key_mask, causal):
        for t, n in ((q, "q"), (k, "k"), (v, "v")):
            L.require_cuda(t, f"attention({n})")
        q, k, v = q.contiguous(), k.contiguous(), v.contiguous()
        B, Lq, HS = q.shape
        Lk = k.shape[1]
        hs = HS // H
        out = torch.empty((B, Lq, HS), dtype=torch.float32, device=q.device)
        sm = torch.empty((B, H, Lq), dtype=torch.float32, device=q.device)
        sil = torch.empty_like(sm)
        rc = L.lib().rtf_attn_fwd(
            q.data_ptr(), q.stride(0), q.stride(1), k.data_ptr(), k.stride(0), k.stride(1),
            v.data_ptr(), v.stride(0), v.stride(1),
            None if row_mask is None else row_mask.data_ptr(), Lq,
            None if key_mask is None else key_mask.data_ptr(), Lk, int(causal), B, H, Lq, Lk, hs,
            float(scale), out.data_ptr(), out.stride(0), out.stride(1), sm.data_ptr(),
            sil.data_ptr(), L.current_stream_ptr())
        L.check(rc, "rtf_attn_fwd")
        ctx.save_for_backward(q, k, v, out, sm, sil)
        ctx.cfg = (H, scale, row_mask, key_mask, causal)
        return out

    @staticmethod
    def backward(ctx, dout):
        q, k, v, out, sm, sil = ctx.saved_tensors
        H, scale, row_mask, key_mask, causal = ctx.cfg
        dout = dout.contiguous()
        B, Lq, HS = q.shape
        Lk = k.shape[1]
        hs = HS // H
        dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
        delta = torch.empty_like(sm)
        rc = L.lib().rtf_attn_bwd(
            q.data_ptr(), q.stride(0), q.stride(1), k.data_ptr(), k.stride(0), k.stride(1),
            v.data_ptr(), v.stride(0), v.stride(1),
            None if row_mask is None else row_mask.data_ptr(), Lq,
            None if key_mask is None else key_mask.data_ptr(), Lk, int(causal), B, H, Lq, Lk, hs,
            float(scale), out.data_ptr(), out.stride(0), out.stride(1), sm.data_ptr(),
            sil.data_ptr(), dout.data_ptr(), dout.stride(0), dout.stride(1), delta.data_ptr(),
            dq.data_ptr(), dq.stride(0), dq.stride(1), dk.data_ptr(), dk.stride(0), dk.stride(1),
            dv.data_ptr(), dv.stride(0), dv.stride(1), L.current_stream_ptr())
        L.check(rc, "rtf_attn_bwd")
        return dq, dk, dv, None, None, None, None, None


def attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, num_heads: int, scale: float,
              row_mask: Optional[torch.Tensor] = None, key_mask: Optional[torch.Tensor] = None,
              causal: bool = False) -> torch.Tensor:
    """softmax(mask(q k^T * scale)) v per head; q (B,Lq,H*hs), k/v (B,Lk,H*hs) -> (B,Lq,H*hs).
    row_mask (B,Lq[,1]) blanks whole query rows (match-side quirk); key_mask (B,Lk)."""
    B, Lq, HS = q.shape
    if HS % num_heads:
        raise ValueError("last dim must be divisible by num_heads")
    rm = _mask2d(row_mask, B, Lq, "attention(row_mask)")
    km = _mask2d(key_mask, B, k.shape[1], "attention(key_mask)")
    return _AttnFn.apply(q, k, v, num_heads, scale, rm, km, causal)
