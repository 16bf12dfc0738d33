"""Generate tests/golden/*.npz by EXECUTING THE UNMODIFIED REFERENCE (laomao0/BIN).

Run in the authoring container only (needs /root/reference):

    python oracle/make_golden.py

The fixtures pin ``oracle/bin_oracle.py`` (and through it the CUDA path) to the
reference's own ``models/archs/RDN.py`` because the reference ships no tests or
golden vectors of its own.  Weights are
``bin_oracle.synth_state_dict(seed)`` loaded with ``strict=True`` -- which also
pins the 1 332-key state_dict schema.  Inputs are ``bin_oracle.synth_frames``.
Nothing under /root/reference is copied; only tensors it computes are stored.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, "/root/reference")

import models.archs.RDN as R          # noqa: E402  (the reference itself)
from oracle import bin_oracle as O    # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
os.makedirs(OUT, exist_ok=True)
torch.set_num_threads(os.cpu_count())


def save(name, **arrs):
    np.savez_compressed(os.path.join(OUT, name), **{k: np.asarray(v) for k, v in arrs.items()})
    print("wrote", name, {k: tuple(np.asarray(v).shape) for k, v in arrs.items()})


def main():
    sd = O.synth_state_dict(0)
    net = R.bin_stage4_lstm().eval()
    missing = net.load_state_dict(sd, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    assert list(net.state_dict().keys()) == list(sd.keys()), "key ORDER differs from reference"

    # -- full window, 14 outputs (RDN.py:422-465) ------------------------------------
    for tag, (B, H, W, smooth) in {"window_a": (1, 32, 48, False), "window_b": (2, 24, 16, True)}.items():
        fr = O.synth_frames(6, B, H, W, seed=1234, smooth=smooth)
        with torch.no_grad():
            outs = net(*fr)
        save(f"{tag}.npz", meta=np.array([B, H, W, int(smooth), 1234, 0]),
             **{f"out{k}": o.numpy() for k, o in enumerate(outs)})

    # -- pyramid with and without previous_input (RDN.py:367-405) ---------------------
    fr = O.synth_frames(5, 1, 16, 24, seed=77)
    prev = [t * 2 - 1 for t in O.synth_frames(6, 1, 16, 24, seed=78)]
    with torch.no_grad():
        o0 = net.model(*fr, [None] * 6)
        o1 = net.model(*fr, prev)
    save("pyramid.npz", meta=np.array([1, 16, 24, 77, 78]),
         **{f"none{k}": o.numpy() for k, o in enumerate(o0)},
         **{f"prev{k}": o.numpy() for k, o in enumerate(o1)})

    # -- the three backbones, odd tile remainders (RDN.py:167-334) --------------------
    for name, n in (("model1_1", 2), ("model2_1", 3), ("model3_1", 5), ("model4_1", 5)):
        m = getattr(net.model, name)
        fr = O.synth_frames(n, 2, 20, 36, seed=100 + n)
        with torch.no_grad():
            y = m(*fr)
        save(f"backbone_{name}.npz", meta=np.array([2, 20, 36, 100 + n]), out=y.numpy())

    # -- RDB / RDB_Conv (RDN.py:135-165) ----------------------------------------------
    g = torch.Generator().manual_seed(5)
    x = torch.randn((2, 96, 11, 13), generator=g)
    blk = net.model.model2_1.RDBs[3]
    with torch.no_grad():
        y = blk(x)
        y0 = blk.convs[0](x)
    save("rdb.npz", x=x.numpy(), out=y.numpy(), conv0=y0.numpy())

    # -- pixel_reshuffle (RDN.py:107-132) ---------------------------------------------
    x = torch.arange(2 * 6 * 4 * 6, dtype=torch.float32).reshape(2, 6, 4, 6)
    save("pixel_reshuffle.npz", x=x.numpy(), out=R.pixel_reshuffle(x, 2).numpy())

    # -- ConvLSTMCell with and without state (RDN.py:50-95) ----------------------------
    cell = net.clstm_7_prime_prime
    g = torch.Generator().manual_seed(6)
    x = torch.rand((2, 3, 10, 14), generator=g)
    c0 = torch.randn((2, 3, 10, 14), generator=g)
    h0 = torch.randn((2, 3, 10, 14), generator=g).tanh()
    with torch.no_grad():
        h1, (c1n, _) = cell(x, None)
        h2, (c2n, _) = cell(x, [c0, h0])
    save("convlstm.npz", x=x.numpy(), c0=c0.numpy(), h0=h0.numpy(),
         h_none=h1.numpy(), c_none=c1n.numpy(), h_state=h2.numpy(), c_state=c2n.numpy())

    # -- tensor2img (utils/util.py:113-137), the image boundary of the caller loop --------
    import utils.util as RU                    # the reference's own helper
    g = torch.Generator().manual_seed(8)
    t = torch.rand((3, 21, 34), generator=g) * 1.4 - 0.2
    t.view(-1)[:6] = torch.tensor([0.5 / 255, 1.5 / 255, 2.5 / 255, 254.5 / 255, 1.0, 0.0])
    save("tensor2img.npz", x=t.numpy(), out=RU.tensor2img(t.clone()))

    # -- bin_model.get_loss (bin_model.py:395-425) called on a stand-in object with the reference's own criteria --------
    import models.bin_model as BM
    from models.loss import CharbonnierLoss

    class _Stub:
        pass
    g = torch.Generator().manual_seed(3)
    louts = [torch.rand((1, 3, 8, 10), generator=g) for _ in range(14)]
    lgts = [torch.rand((1, 3, 8, 10), generator=g) for _ in range(14)]
    stub = _Stub()
    stub.Ft_p, stub.nframes, stub.version, stub.get_info = louts, 6, 2, (lambda mode=1: (14, lgts))
    rec = {}
    for name, cri in (("l1", torch.nn.L1Loss(reduction="sum")), ("l2", torch.nn.MSELoss(reduction="sum")), ("cb", CharbonnierLoss())):
        stub.cri_pix = cri
        loss, ll = BM.bin_model.get_loss(stub, ret=1)
        rec[name] = loss.numpy()
        rec[name + "_list"] = torch.stack(ll).numpy()
    save("get_loss.npz", outs=torch.stack(louts).numpy(), gts=torch.stack(lgts).numpy(), **rec)

    # -- backward: d(sum_k <out_k, cot_k>)/d(frames, a few params) (config 3) ----------
    net.train()
    fr = [f.requires_grad_(True) for f in O.synth_frames(6, 1, 16, 16, seed=9)]
    outs = net(*fr)
    cots = O.synth_frames(14, 1, 16, 16, seed=10)
    loss = sum((o * (c - 0.5)).sum() for o, c in zip(outs, cots))
    names = ["model.model1_1.SFENet1.weight", "model.model1_1.RDBs.0.convs.0.conv.0.weight",
             "model.model2_1.RDBs.5.LFF.weight", "model.model3_1.GFF.0.weight",
             "model.model4_1.UPNet.2.weight", "model.model4_1.UPNet.0.bias",
             "clstm_4_prime.Gates.weight", "clstm_6_prime_prime_prime.Gates.bias"]
    params = dict(net.named_parameters())
    grads = torch.autograd.grad(loss, fr + [params[n] for n in names])
    save("window_grad.npz", meta=np.array([1, 16, 16, 9, 10]), loss=loss.detach().numpy(),
         **{f"dframe{k}": g_.numpy() for k, g_ in enumerate(grads[:6])},
         **{"d:" + n: g_.numpy() for n, g_ in zip(names, grads[6:])})


if __name__ == "__main__":
    main()
