"""CPU oracle for the BIN hot path -- TEST INFRASTRUCTURE ONLY.

This file is a plain fp32 PyTorch restatement of the arithmetic the reference
performs in ``models/archs/RDN.py`` (laomao0/BIN).  It exists so that the CUDA
path in ``bin_b200`` can be checked on a box where ``/root/reference`` is not
mounted.  Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline``
/ ``--impl reference`` legs of ``bench.py`` may import it; the product package
``bin_b200`` never does (and fails loudly without its CUDA library).

Parity pinning: the reference ships no tests or golden vectors,
so this restatement is pinned by *executing the reference's own RDN.py* in the
authoring container: ``oracle/make_golden.py`` imports the unmodified reference,
loads ``synth_state_dict`` weights, runs seeded inputs and commits the outputs
under ``tests/golden/``.  ``tests/test_oracle_golden.py`` checks this file
against those fixtures bit-for-bit-close (fp32, <=2e-6).

Every function cites the reference lines it follows.  Weights are passed as a
flat ``state_dict`` using the reference's key schema (1 332 keys).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
SD = Dict[str, Tensor]

G0, D, C, G = 96, 12, 4, 32           # RDN.py:418 (GO=96, D=12), :171-172 (C=4, G=32)
LSTM_NAMES = ["clstm_4_prime", "clstm_6_prime", "clstm_8_prime",
              "clstm_5_prime_prime", "clstm_7_prime_prime",
              "clstm_6_prime_prime_prime"]                      # RDN.py:412-417
# canonical backbone -> aliases that re-emit the same tensors (RDN.py:342-365)
BACKBONE_ALIASES = {
    "model1_1": ["model1_1", "model1_2", "model1_3", "model1_4"],
    "model2_1": ["model2_1", "model2_2", "model2_3"],
    "model3_1": ["model3_1", "model3_2"],
    "model4_1": ["model4_1"],
}
BACKBONE_NFRAMES = {"model1_1": 2, "model2_1": 3, "model3_1": 5, "model4_1": 5}


# --------------------------------------------------------------------------- #
# schema + synthetic weights
# --------------------------------------------------------------------------- #
def backbone_param_shapes(nframes: int) -> List[Tuple[str, Tuple[int, ...]]]:
    """(suffix, shape) for one backbone, in the reference's registration order
    (RDN.py:187-208 / 245-266 / 299-320)."""
    cin0 = 12 * nframes                                     # 3 ch * 4 (s2d) per frame
    out: List[Tuple[str, Tuple[int, ...]]] = []
    out += [("SFENet1.weight", (G0, cin0, 5, 5)), ("SFENet1.bias", (G0,))]
    out += [("SFENet2.weight", (G0, G0, 3, 3)), ("SFENet2.bias", (G0,))]
    for i in range(D):
        for c in range(C):
            out += [(f"RDBs.{i}.convs.{c}.conv.0.weight", (G, G0 + c * G, 3, 3)),
                    (f"RDBs.{i}.convs.{c}.conv.0.bias", (G,))]
        out += [(f"RDBs.{i}.LFF.weight", (G0, G0 + C * G, 1, 1)),
                (f"RDBs.{i}.LFF.bias", (G0,))]
    out += [("GFF.0.weight", (G0, D * G0, 1, 1)), ("GFF.0.bias", (G0,))]
    out += [("GFF.1.weight", (G0, G0, 3, 3)), ("GFF.1.bias", (G0,))]
    out += [("UPNet.0.weight", (256, G0, 3, 3)), ("UPNet.0.bias", (256,))]
    out += [("UPNet.2.weight", (3, 64, 3, 3)), ("UPNet.2.bias", (3,))]
    return out


def _uniform(shape, bound, gen):
    return (torch.rand(shape, generator=gen, dtype=torch.float32) * 2 - 1) * bound


def synth_backbone_sd(nframes: int, seed: int) -> SD:
    """Deterministic weights with PyTorch-default-like scale U(+-1/sqrt(fan_in))."""
    gen = torch.Generator().manual_seed(seed)
    sd: SD = {}
    fan_in = 1
    for name, shape in backbone_param_shapes(nframes):
        if name.endswith("weight"):
            fan_in = shape[1] * shape[2] * shape[3]
        sd[name] = _uniform(shape, 1.0 / math.sqrt(fan_in), gen)
    return sd


def synth_state_dict(seed: int = 0) -> SD:
    """Full ``bin_stage4_lstm`` state_dict (1 332 keys; aliases share storage)."""
    sd: SD = {}
    gen = torch.Generator().manual_seed(seed * 1000 + 7)
    for n in LSTM_NAMES:                                    # xavier-uniform, bias made non-zero on purpose
        bound = math.sqrt(6.0 / (6 * 9 + 12 * 9))
        sd[f"{n}.Gates.weight"] = _uniform((12, 6, 3, 3), bound, gen)
        sd[f"{n}.Gates.bias"] = _uniform((12,), 0.1, gen)
    for k, (canon, aliases) in enumerate(BACKBONE_ALIASES.items()):
        bsd = synth_backbone_sd(BACKBONE_NFRAMES[canon], seed * 1000 + 100 + k)
        for a in aliases:
            for name, t in bsd.items():
                sd[f"model.{a}.{name}"] = t
    return sd


def sub_sd(sd: SD, prefix: str) -> SD:
    p = prefix + "."
    return {k[len(p):]: v for k, v in sd.items() if k.startswith(p)}


def synth_frames(n: int, B: int, H: int, W: int, seed: int = 1234, smooth: bool = False) -> List[Tensor]:
    """Frames ~U[0,1) drawn in frame order from one generator."""
    gen = torch.Generator().manual_seed(seed)
    frames = [torch.rand((B, 3, H, W), generator=gen, dtype=torch.float32) for _ in range(n)]
    if smooth:                                              # "natural-ish": box-blurred noise
        k = torch.ones(3, 1, 5, 5) / 25.0
        frames = [F.conv2d(F.pad(f, (2, 2, 2, 2), mode="replicate"), k, groups=3) for f in frames]
    return frames


# --------------------------------------------------------------------------- #
# arithmetic
# --------------------------------------------------------------------------- #
def space_to_depth2(x: Tensor) -> Tensor:
    """RDN.py:107-132 pixel_reshuffle(x, 2): out[c*4+i*2+j, y, x] = in[c, 2y+i, 2x+j]."""
    B, Cc, H, W = x.shape
    v = x.reshape(B, Cc, H // 2, 2, W // 2, 2).permute(0, 1, 3, 5, 2, 4)
    return v.reshape(B, Cc * 4, H // 2, W // 2)


_EMULATE_FP16 = False      # tests only: round weights / stored activations to fp16 like the CUDA path stores them
_EMULATE_FP16_GRADS = False


class emulate_fp16_storage:
    """Context manager: the oracle rounds conv weights and every tensor the CUDA path stores in fp16
    (packed input, each conv's stored output) to fp16 precision (straight-through for autograd).  Used by the
    backward tests to separate kernel errors from ReLU sign flips caused by fp16-vs-fp32 forward differences.
    grads=True additionally rounds the GRADIENT of every such tensor to fp16 after power-of-two loss scaling, which
    is how the CUDA backward stores dY between layers (bin_b200/autograd.py: scale = 2^floor(log2(2048 / max|g|)))."""

    def __init__(self, grads: bool = False):
        self.grads = grads

    def __enter__(self):
        global _EMULATE_FP16, _EMULATE_FP16_GRADS
        self.prev = (_EMULATE_FP16, _EMULATE_FP16_GRADS)
        _EMULATE_FP16, _EMULATE_FP16_GRADS = True, self.grads

    def __exit__(self, *a):
        global _EMULATE_FP16, _EMULATE_FP16_GRADS
        _EMULATE_FP16, _EMULATE_FP16_GRADS = self.prev


def _round_grad_fp16(g: Tensor) -> Tensor:
    m = g.abs().max()
    if not torch.isfinite(m) or m == 0:
        return g
    s = torch.exp2(torch.floor(torch.log2(2048.0 / m)))
    return (g * s).half().to(g.dtype) / s


def _q(t: Tensor) -> Tensor:
    if not _EMULATE_FP16:
        return t
    out = t + (t.half().to(t.dtype) - t).detach()
    if _EMULATE_FP16_GRADS and out.requires_grad:
        out.register_hook(_round_grad_fp16)
    return out


def conv(x: Tensor, sd: SD, name: str) -> Tensor:
    w = _q(sd[name + ".weight"])
    return F.conv2d(x, w, sd[name + ".bias"], stride=1, padding=w.shape[-1] // 2)


def rdb(x: Tensor, sd: SD, prefix: str) -> Tensor:
    """RDN.py:135-165: 4 x (conv3x3 -> ReLU -> concat) -> LFF 1x1 -> + x."""
    feat = x
    for c in range(C):
        g = _q(F.relu(conv(feat, sd, f"{prefix}.convs.{c}.conv.0")))      # :140-146
        feat = torch.cat((feat, g), 1)                                     # :147
    return _q(conv(feat, sd, f"{prefix}.LFF") + x)                         # :165


def backbone(frames: Sequence[Tensor], sd: SD, return_feats: bool = False):
    """RDN.py:210-222 / 268-280 / 322-334 (identical up to the frame count)."""
    x0 = _q(space_to_depth2(torch.cat(list(frames), 1)))                   # :211
    f1 = _q(conv(x0, sd, "SFENet1"))                                       # :212
    x = _q(conv(f1, sd, "SFENet2"))                                        # :213
    outs = []
    for i in range(D):                                                     # :215-217
        x = rdb(x, sd, f"RDBs.{i}")
        outs.append(x)
    x = _q(conv(_q(conv(torch.cat(outs, 1), sd, "GFF.0")), sd, "GFF.1") + f1)   # :218-219
    up = _q(F.pixel_shuffle(conv(x, sd, "UPNet.0"), 2))                    # :205-206
    y = conv(up, sd, "UPNet.2")                                            # :207
    mean = sum(frames) / float(len(frames))                                # :221 / :279 / :333
    out = y + mean
    if return_feats:
        return out, {"x0": x0, "f1": f1, "rdb": outs, "gff": x, "up": up}
    return out


def convlstm(x: Tensor, sd: SD, prefix: str, state: Optional[Tuple[Tensor, Tensor]] = None):
    """RDN.py:50-95.  state=(c,h); None -> zeros.  Returns h', (c', h')."""
    if state is None:
        z = torch.zeros_like(x)
        state = (z, z)
    c, h = state                                                           # :71
    g = conv(torch.cat((x, h), 1), sd, prefix + ".Gates")                  # :74-75
    i, j, f, o = g.chunk(4, 1)                                             # :79
    c2 = c * torch.sigmoid(f + 1.0) + torch.sigmoid(i) * torch.tanh(j)     # :81 (forget_bias=1.0, :16)
    h2 = torch.tanh(c2) * torch.sigmoid(o)                                 # :82
    return h2, (c2, h2)


def pyramid(fr: Sequence[Tensor], prev: Sequence[Optional[Tensor]], sd: SD) -> List[Tensor]:
    """RDN.py:367-405, lstm branch.  sd = state_dict of the '...5_input' module."""
    B1, B3, B5, B7, B9 = fr
    m1, m2, m3, m4 = (sub_sd(sd, k) for k in ("model1_1", "model2_1", "model3_1", "model4_1"))
    I2 = backbone((B1, B3), m1); I4 = backbone((B3, B5), m1)               # :371-374
    I6 = backbone((B5, B7), m1); I8 = backbone((B7, B9), m1)
    if prev[0] is not None:                                                # :375-382
        p4, p6, p8, p5, p7, p6b = prev
        I3 = backbone((p4, I2, I4), m2); I5 = backbone((p6, I4, I6), m2); I7 = backbone((p8, I6, I8), m2)
        I4b = backbone((p5, B3, I3, I5, B5), m3); I6b = backbone((p7, B5, I5, I7, B7), m3)
        I5c = backbone((p6b, I4, I4b, I6b, I6), m4)
    else:                                                                  # :383-389
        I3 = backbone((I2, I2, I4), m2); I5 = backbone((I4, I4, I6), m2); I7 = backbone((I6, I6, I8), m2)
        I4b = backbone((I3, B3, I3, I5, B5), m3); I6b = backbone((I5, B5, I5, I7, B7), m3)
        I5c = backbone((I4, I4, I4b, I6b, I6), m4)
    return [I2, I4, I6, I8, I3, I5, I7, I4b, I6b, I5c]                    # :403-405


def window_forward(frames: Sequence[Tensor], sd: SD) -> List[Tensor]:
    """RDN.py:422-465: two pyramid steps linked by six ConvLSTM cells; 14 outputs."""
    assert len(frames) == 6
    msd = sub_sd(sd, "model")
    prev: List[Optional[Tensor]] = [None] * 6
    res = []
    for step in range(2):                                                  # :438
        out = pyramid(frames[step:step + 5], prev, msd)                    # :442
        hid = [out[1], out[2], out[3], out[5], out[6], out[8]]             # :443-448
        # every cell is called with prev_state=None at step 0; the step-1 calls are dead (:449-465)
        prev = [convlstm(hid[k], sd, LSTM_NAMES[k], None)[0] for k in range(6)] if step == 0 else prev
        res.append(out)
    r0, r1 = res
    return r0[:10] + [r1[3], r1[6], r1[8], r1[9]]                          # :461-465


def pyramid3_4frames(frames: Sequence[Tensor], sd: SD) -> List[Tensor]:
    """BASELINE config 2a/3a: stages 1-3 over 4 frames, first-window semantics
    (pattern of RDN.py:383-387).  6 outputs."""
    B1, B3, B5, B7 = frames
    msd = sub_sd(sd, "model")
    m1, m2, m3 = (sub_sd(msd, k) for k in ("model1_1", "model2_1", "model3_1"))
    I2 = backbone((B1, B3), m1); I4 = backbone((B3, B5), m1); I6 = backbone((B5, B7), m1)
    I3 = backbone((I2, I2, I4), m2); I5 = backbone((I4, I4, I6), m2)
    I4b = backbone((I3, B3, I3, I5, B5), m3)
    return [I2, I4, I6, I3, I5, I4b]


# --------------------------------------------------------------------------- #
# metrics (reference utils/util.py:113-137, 201-208)
# --------------------------------------------------------------------------- #
def tensor2img_u8(t: Tensor) -> Tensor:
    """clamp to [0,1], *255, round -> uint8 (utils/util.py:113-137, min_max=(0,1))."""
    return (t.clamp(0, 1) * 255.0).round().to(torch.uint8)


def psnr_u8(a: Tensor, b: Tensor) -> float:
    """utils/util.py:201-208 on uint8-range images."""
    mse = ((a.double() - b.double()) ** 2).mean().item()
    if mse == 0:
        return float("inf")
    return 20.0 * math.log10(255.0 / math.sqrt(mse))


def get_loss_6v2(outs: Sequence[Tensor], gts: Sequence[Tensor], kind: str = "l1", eps: float = 1e-6):
    """bin_model.get_loss(ret=1) for nframes == 6, version == 2 (bin_model.py:395-425) with cri_pix = L1 sum (:55),
    MSE sum (:57) or CharbonnierLoss (loss.py:130-140).  Returns (loss, loss_list[:num])."""
    def cri(x, y):
        d = x - y
        if kind == "l1":
            return d.abs().sum()
        if kind == "l2":
            return (d * d).sum()
        return torch.sqrt(d * d + eps).mean()
    terms = [cri(o, g) for o, g in zip(outs, gts)]                                   # :400-402
    terms += [cri(outs[1], outs[7]), cri(outs[5], outs[9]), cri(outs[2], outs[8])]   # :409-415
    return sum(terms) / len(terms), terms[:len(outs)]                                # :416, :418


def read_image_u8(img_bgr_u8):
    """test.py:44-56 read_image on an already decoded uint8 HWC BGR array -> (3,H,W) fp32 RGB in [0,1]."""
    import numpy as np
    img = np.asarray(img_bgr_u8).astype(np.float32) / 255.
    img = img[:, :, [2, 1, 0]]
    return torch.from_numpy(np.ascontiguousarray(np.transpose(img, (2, 0, 1)))).float()


def tensor2img_bgr_u8(t: Tensor):
    """utils/util.py:113-137 tensor2img(tensor, np.uint8, (0,1)) for one (3,H,W) image -> uint8 HWC BGR numpy array."""
    import numpy as np
    t = t.squeeze().float().cpu().clamp(0, 1)
    img = t.numpy()
    img = np.transpose(img[[2, 1, 0], :, :], (1, 2, 0))
    return (img * 255.0).round().astype(np.uint8)


def adam_step(p, g, m, v, step: int, lr: float, beta1: float, beta2: float, eps: float, weight_decay: float):
    """One torch.optim.Adam step (the optimizer of bin_model.py:97-100,141; torch/optim/adam.py _single_tensor_adam,
    amsgrad=False) restated on fp32 numpy arrays; returns new (p, m, v).  `step` is the 1-based step number."""
    import numpy as np
    f = np.float32
    if weight_decay != 0:
        g = g + f(weight_decay) * p
    m = m + (g - m) * f(1 - beta1)
    v = v * f(beta2) + f(1 - beta2) * g * g
    bc1, bc2 = 1 - beta1 ** step, 1 - beta2 ** step
    denom = np.sqrt(v) / f(math.sqrt(bc2)) + f(eps)
    p = p - f(lr / bc1) * (m / denom)
    return p.astype(f), m.astype(f), v.astype(f)


def blur_average(frames_u8, window_size: int = 11, first_mid: int = 16, stride: int = 8, nwin: Optional[int] = None):
    """create_dataset_blur_N_frames_average.py:99-131 on a (T, ...) uint8 numpy array: blurry frame w is the float32
    sum of the 2r+1 frames around first_mid + w*stride (r = int((window_size-1)/2), :101), divided by their count
    (:129) and truncated by .astype("uint8") (:130); nwin defaults to floor(T/8) - 2 (:104)."""
    import numpy as np
    T = frames_u8.shape[0]
    r = int((window_size - 1) / 2)
    if nwin is None:
        nwin = math.floor(T / stride) - 2
    out = []
    for w in range(nwin):
        mid = first_mid + w * stride
        acc = np.zeros(frames_u8.shape[1:], dtype=np.float32)
        for loc in range(mid - r, mid + r + 1):
            acc = acc + frames_u8[loc].astype("float32")
        out.append((acc / np.float32(2 * r + 1)).astype("uint8"))
    return np.stack(out)

CONV_MACS_PER_PX_WINDOW = 14_234_976
def window_flops(H: int, W: int, B: int = 1) -> float:
    return 2.0 * CONV_MACS_PER_PX_WINDOW * H * W * B
