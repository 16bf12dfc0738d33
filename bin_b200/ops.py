"""Thin torch-tensor wrappers over the C ABI (device pointers + current stream).

These are the unit-level entry points used by the tests and by bin_b200.rdn; none of them
computes anything in Python/PyTorch.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence

import torch

from . import _lib
from ._lib import Act, ConvArgs, Frames, check, lib


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _req(t: torch.Tensor, dtype, name: str) -> torch.Tensor:
    if not t.is_cuda:
        raise _lib.BinB200Error(f"{name}: expected a CUDA tensor (bin_b200 has no CPU path)")
    if t.dtype != dtype:
        raise _lib.BinB200Error(f"{name}: expected dtype {dtype}, got {t.dtype}")
    return t.contiguous()


def act_view(t: torch.Tensor) -> Act:
    """P8 tensor [B, planes, H, W, 8] fp16 -> bin_act_t."""
    assert t.dtype == torch.float16 and t.dim() == 5 and t.shape[-1] == 8 and t.is_contiguous()
    return Act(t.data_ptr(), t.shape[0], t.shape[1], t.shape[2], t.shape[3])


def empty_p8(B: int, planes: int, H: int, W: int, device) -> torch.Tensor:
    return torch.empty((B, planes, H, W, 8), dtype=torch.float16, device=device)


def nchw_to_p8(x: torch.Tensor, dst: Optional[torch.Tensor] = None, plane0: int = 0, pad_to: int = 32) -> torch.Tensor:
    x = _req(x, torch.float32, "x")
    B, Cc, H, W = x.shape
    if dst is None:
        planes = ((Cc + pad_to - 1) // pad_to * pad_to) // 8
        dst = torch.zeros((B, planes, H, W, 8), dtype=torch.float16, device=x.device)
    check(lib().bin_nchw_to_p8(x.data_ptr(), Cc, act_view(dst), plane0, _stream()))
    return dst


def p8_to_nchw(src: torch.Tensor, C_: int, plane0: int = 0) -> torch.Tensor:
    B, _, H, W, _ = src.shape
    y = torch.empty((B, C_, H, W), dtype=torch.float32, device=src.device)
    check(lib().bin_p8_to_nchw(act_view(src), plane0, C_, y.data_ptr(), _stream()))
    return y


def _req_contig(t: torch.Tensor, dtype, name: str) -> torch.Tensor:
    """Like _req but never copies: the caller stores a raw pointer, and a temporary `.contiguous()` copy would be freed
    (and its block handed to the next allocation on this stream) before the kernel that reads it is launched."""
    if not t.is_cuda:
        raise _lib.BinB200Error(f"{name}: expected a CUDA tensor (bin_b200 has no CPU path)")
    if t.dtype != dtype:
        raise _lib.BinB200Error(f"{name}: expected dtype {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise _lib.BinB200Error(f"{name}: expected a contiguous tensor (call .contiguous() and keep the result alive)")
    return t


def make_frames(calls: Sequence[Sequence[torch.Tensor]], outs: Sequence[Optional[torch.Tensor]]) -> Frames:
    """Pointer table of a batched backbone launch.  Holds raw pointers: every tensor must be contiguous and must stay
    alive until the launch that consumes the table has been enqueued."""
    fr = Frames()
    fr.ncalls = len(calls)
    fr.nframes = len(calls[0])
    fr.Bc = calls[0][0].shape[0]
    for k, frames in enumerate(calls):
        assert len(frames) == fr.nframes
        for f, t in enumerate(frames):
            fr.frame[k][f] = _req_contig(t, torch.float32, "frame").data_ptr()
        fr.out[k] = _ptr(outs[k])
    return fr


def pack_frames(calls: Sequence[Sequence[torch.Tensor]]) -> torch.Tensor:
    """RDN.py:211 + 107-132: concat + space-to-depth + fp16 cast, batched over calls."""
    fr = make_frames(calls, [None] * len(calls))
    B, _, H, W = calls[0][0].shape
    cin_pad = (12 * fr.nframes + 31) // 32 * 32
    dst = empty_p8(fr.ncalls * B, cin_pad // 8, H // 2, W // 2, calls[0][0].device)
    check(lib().bin_pack_frames(C.byref(fr), H, W, act_view(dst), _stream()))
    return dst


def pack_conv_weight(w: torch.Tensor, cout_pad: int, cin_pad: int, variant: int = 0) -> torch.Tensor:
    w = _req(w, torch.float32, "weight")
    cout, cin, k, _ = w.shape
    out = torch.empty(lib().bin_packed_weight_bytes(cout_pad, cin_pad, k) // 2, dtype=torch.float16, device=w.device)
    check(lib().bin_pack_conv_weight(w.data_ptr(), cout, cin, k, cout_pad, cin_pad, variant, out.data_ptr(), _stream()))
    return out


def pad_bias(b: torch.Tensor, cout_pad: int) -> torch.Tensor:
    out = torch.zeros(cout_pad, dtype=torch.float32, device=b.device)
    out[: b.numel()] = b
    return out


def conv_fwd(in0: torch.Tensor, w_packed: torch.Tensor, bias_pad: torch.Tensor, ksize: int, cout_pad: int, *,
             in0_plane0: int = 0, in0_planes: Optional[int] = None,
             in1: Optional[torch.Tensor] = None, in1_plane0: int = 0, in1_planes: int = 0,
             relu: bool = False, epilogue: int = _lib.EPI_P8,
             out: Optional[torch.Tensor] = None, out_plane0: int = 0,
             res: Optional[torch.Tensor] = None, res_plane0: int = 0,
             frames: Optional[Frames] = None, variant: int = 0, sub=None, store_planes: int = 0) -> None:
    a = ConvArgs()
    a.in0 = act_view(in0)
    a.in0_plane0 = in0_plane0
    a.in0_planes = in0.shape[1] - in0_plane0 if in0_planes is None else in0_planes
    if in1 is not None and in1_planes > 0:
        a.in1 = act_view(in1)
        a.in1_plane0, a.in1_planes = in1_plane0, in1_planes
    a.w_packed = w_packed.data_ptr()
    a.bias = bias_pad.data_ptr()
    a.ksize, a.cout_pad, a.relu, a.epilogue, a.variant = ksize, cout_pad, int(relu), epilogue, variant
    if out is not None:
        a.out = act_view(out)
        a.out_plane0 = out_plane0
    if res is not None:
        a.res = act_view(res)
        a.res_plane0 = res_plane0
    if frames is not None:
        a.fr = frames
    if sub is not None:
        a.b_begin, a.b_count, a.y_begin, a.y_count = sub
    a.store_planes = store_planes
    check(lib().bin_conv_fwd(C.byref(a), _stream()))


def rdb_tail_fwd(x: torch.Tensor, g: torch.Tensor, w_conv: torch.Tensor, b_conv: torch.Tensor, w_lff: torch.Tensor,
                 b_lff: torch.Tensor, out: torch.Tensor, *, x_plane0: int = 0, g_plane0: int = 0, out_plane0: int = 0,
                 sub=(0, 0, 0, 0)) -> None:
    """Fused conv3 + LFF + residual of one RDB (RDN.py:141-147, 162-165) on P8 tensors; see bin_rdb_tail_fwd."""
    ax, ag, ao = act_view(x), act_view(g), act_view(out)
    check(lib().bin_rdb_tail_fwd(C.byref(ax), x_plane0, C.byref(ag), g_plane0, w_conv.data_ptr(), b_conv.data_ptr(),
                                 w_lff.data_ptr(), b_lff.data_ptr(), C.byref(ao), out_plane0, *sub, _stream()))

def convlstm_fwd(x, w, b, state=None):
    """ConvLSTMCell.forward (RDN.py:50-95) -> (h, c)."""
    x = _req(x, torch.float32, "x")
    B, _, H, W = x.shape
    h = torch.empty_like(x)
    c = torch.empty_like(x)
    cp = hp = None
    if state is not None:
        cp, hp = (_req(t, torch.float32, "state") for t in state)
    check(lib().bin_convlstm_fwd(x.data_ptr(), _ptr(cp), _ptr(hp), _req(w, torch.float32, "w").data_ptr(),
                                 _req(b, torch.float32, "b").data_ptr(), h.data_ptr(), c.data_ptr(), B, H, W, _stream()))
    return h, c
