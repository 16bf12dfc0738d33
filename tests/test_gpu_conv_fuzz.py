"""Seeded shape fuzzing of bin_conv_fwd (every kernel instantiation the backbone uses) against F.conv2d on the
same fp16-rounded operands: odd tile remainders in H and W, batch > 1, two-segment inputs, sub-range launches,
residual / ReLU / PixelShuffle / final epilogues.  Reference on CPU in fp64 (cuDNN fp32 is not trustworthy at 1e-3)."""
import random

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

LAYERS = [  # (cin, cout, k, epilogue, relu, residual, split)
    (24, 96, 5, 0, False, False, None), (36, 96, 5, 0, False, False, None), (60, 96, 5, 0, False, False, None),
    (96, 96, 3, 0, False, True, None), (96, 32, 3, 0, True, False, None), (128, 32, 3, 0, True, False, 96),
    (160, 32, 3, 0, True, False, 96), (192, 32, 3, 0, True, False, 96), (224, 96, 1, 0, False, True, 96),
    (1152, 96, 1, 0, False, False, None), (96, 256, 3, 1, False, False, None), (64, 3, 3, 2, False, False, None),
]


def _case(seed):
    rnd = random.Random(seed)
    layer = LAYERS[seed % len(LAYERS)]
    B = rnd.choice([1, 2, 3])
    H, W = rnd.randint(1, 70), rnd.randint(1, 100)
    sub = None
    if layer[3] == 0 and rnd.random() < 0.4:
        b0 = rnd.randrange(B)
        y0 = rnd.randrange(H)
        sub = (b0, rnd.randint(1, B - b0), y0, rnd.randint(1, H - y0))
    return layer, B, H, W, sub


@pytest.mark.parametrize("seed", range(36))
def test_conv_fuzz(seed):
    from bin_b200 import ops
    (cin, cout, k, epi, relu, res, split), B, H, W, sub = _case(seed)
    g = torch.Generator().manual_seed(1000 + seed)
    dev = "cuda"
    x = torch.randn((B, cin, H, W), generator=g).half().float()
    w = (torch.randn((cout, cin, k, k), generator=g) / (cin * k * k) ** 0.5).half().float()
    b = torch.randn((cout,), generator=g) * 0.1
    ref = F.conv2d(x.double(), w.double(), b.double(), padding=k // 2)
    if relu:
        ref = ref.relu()
    cin_pad, cout_pad = (cin + 31) // 32 * 32, (16 if cout == 3 else cout)
    wp, bp = ops.pack_conv_weight(w.to(dev), cout_pad, cin_pad), ops.pad_bias(b.to(dev), cout_pad)
    if split is None:
        in0, in1, p0, p1 = ops.nchw_to_p8(x.to(dev)), None, cin_pad // 8, 0
    else:
        in0, in1 = ops.nchw_to_p8(x[:, :split].contiguous().to(dev)), ops.nchw_to_p8(x[:, split:].contiguous().to(dev))
        p0, p1 = split // 8, (cin - split) // 8
    if epi == 0:
        out = torch.full((B, cout_pad // 8, H, W, 8), 3.0, dtype=torch.float16, device=dev)
        r = None
        if res:
            rx = torch.randn((B, cout, H, W), generator=g).half().float()
            r = ops.nchw_to_p8(rx.to(dev))
            ref = ref + rx.double()
        ops.conv_fwd(in0, wp, bp, k, cout_pad, in0_planes=p0, in1=in1, in1_planes=p1, relu=relu, out=out, res=r, sub=sub)
        got = ops.p8_to_nchw(out, cout).cpu().double()
        if sub is not None:
            b0, nb, y0, ny = sub
            mask = torch.zeros_like(ref)
            mask[b0:b0 + nb, :, y0:y0 + ny] = 1
            assert ((got - 3.0) * (1 - mask)).abs().max().item() == 0.0          # nothing outside the sub-range is touched
            got, ref = got * mask, ref * mask
    elif epi == 1:
        out = torch.zeros((B, 8, 2 * H, 2 * W, 8), dtype=torch.float16, device=dev)
        ops.conv_fwd(in0, wp, bp, k, cout_pad, epilogue=1, out=out)
        got, ref = ops.p8_to_nchw(out, 64).cpu().double(), F.pixel_shuffle(ref, 2)
    else:
        nf = 2 + seed % 4
        frames = [[torch.rand((1, 3, H, W), generator=g).to(dev) for _ in range(nf)] for _ in range(B)]
        outs = [torch.zeros(1, 3, H, W, device=dev) for _ in range(B)]
        ops.conv_fwd(in0, wp, bp, k, cout_pad, epilogue=2, frames=ops.make_frames(frames, outs))
        got = torch.cat(outs, 0).cpu().double()
        ref = ref + torch.cat([sum(f).cpu().double() / nf for f in frames], 0)
    torch.cuda.synchronize()
    tol = (1e-5 if epi == 2 else 1.5e-3) * max(1.0, ref.abs().max().item())   # fp16 output rounding (fp32 for the final layer)
    assert (got - ref).abs().max().item() <= tol, (cin, cout, k, epi, B, H, W, sub)
