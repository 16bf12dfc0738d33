"""Fused pixel loss of the training step: bin_model.get_loss, bin_model.py:395-425.

`pixel_loss(outs, gts, kind)` returns (loss, loss_list) exactly like `get_loss(ret=1)` for the shipped configuration
(nframes == 6, version == 2: 14 output-vs-GT terms + 3 cycle terms on output pairs (1,7), (5,9), (2,8), averaged over
the 17 terms).  One reduction launch and one gradient launch replace ~50 element-wise / reduce kernels."""
from __future__ import annotations

import ctypes as C
from typing import List, Sequence, Tuple

import torch

from ._lib import BinB200Error, check, lib

KINDS = {"l1": 0, "l2": 1, "cb": 2}                   # bin_model.py:54-59
CYCLE_PAIRS_6V2 = ((1, 7), (5, 9), (2, 8))            # bin_model.py:409-416


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


class PixelLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, kind: int, eps: float, npairs: int, *tensors):
        a = [t.detach().contiguous().float() for t in tensors[:npairs]]
        b = [t.detach().contiguous().float() for t in tensors[npairs:]]
        n = a[0].numel()
        dev = a[0].device
        if not a[0].is_cuda:
            raise BinB200Error("bin_b200.loss runs on CUDA tensors only (no CPU fallback)")
        if any(t.numel() != n or t.device != dev for t in a + b):
            raise BinB200Error("pixel_loss: all tensors must share size and device")
        with torch.cuda.device(dev):
            pair = torch.empty(npairs, device=dev)
            ap = (C.c_void_p * npairs)(*[t.data_ptr() for t in a])
            bp = (C.c_void_p * npairs)(*[t.data_ptr() for t in b])
            check(lib().bin_pixel_loss_fwd(ap, bp, npairs, n, kind, eps, pair.data_ptr(), _stream()))
        ctx.kind, ctx.eps, ctx.npairs, ctx.n = kind, eps, npairs, n
        ctx.need_b = [t.requires_grad for t in tensors[npairs:]]
        ctx.save_for_backward(*a, *b)
        ctx.mark_non_differentiable(pair)
        return pair.sum() / npairs, pair

    @staticmethod
    def backward(ctx, gtotal, _gpair):
        npairs = ctx.npairs
        saved = ctx.saved_tensors
        a, b = saved[:npairs], saved[npairs:]
        dev = a[0].device
        with torch.cuda.device(dev):
            up = gtotal.reshape(1).float().contiguous()
            da = [torch.empty_like(t) for t in a]
            db = [torch.empty_like(t) if need else None for t, need in zip(b, ctx.need_b)]
            P = lambda ts: (C.c_void_p * npairs)(*[None if t is None else t.data_ptr() for t in ts])
            check(lib().bin_pixel_loss_bwd(P(a), P(b), P(da), P(db), npairs, ctx.n, ctx.kind, ctx.eps, up.data_ptr(), _stream()))
        return (None, None, None, *da, *db)


def pixel_loss(outs: Sequence[torch.Tensor], gts: Sequence[torch.Tensor], kind: str = "l1", eps: float = 1e-6,
               cycle_pairs: Sequence[Tuple[int, int]] = CYCLE_PAIRS_6V2) -> Tuple[torch.Tensor, List[torch.Tensor]]:
    if kind not in KINDS:
        raise BinB200Error(f"unknown pixel criterion {kind!r} (bin_model.py:54-61 knows l1, l2, cb)")
    if len(outs) != len(gts):
        raise BinB200Error("pixel_loss: one target per output (bin_model.get_info)")
    a = list(outs) + [outs[i] for i, _ in cycle_pairs]
    b = list(gts) + [outs[j] for _, j in cycle_pairs]
    total, pair = PixelLossFn.apply(KINDS[kind], float(eps), len(a), *a, *b)
    return total, [pair[k] for k in range(len(outs))]
