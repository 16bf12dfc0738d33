"""Pins oracle/bin_oracle.py to the fixtures produced by the unmodified reference
(oracle/make_golden.py).  fp32 CPU both sides -> tolerance 2e-6 (accumulation-order noise)."""
import os

import numpy as np
import pytest
import torch

from oracle import bin_oracle as O

TOL = 2e-6


def _load(golden_dir, name):
    return {k: torch.from_numpy(v) for k, v in np.load(os.path.join(golden_dir, name)).items()}


@pytest.fixture(scope="module")
def sd():
    return O.synth_state_dict(0)


def test_schema(sd):
    assert len(sd) == 1332
    uniq = {t.data_ptr(): t for t in sd.values()}
    assert len(uniq) == 540
    assert sum(t.numel() for t in uniq.values()) == 11_441_668


@pytest.mark.parametrize("tag", ["window_a", "window_b"])
def test_window(golden_dir, sd, tag):
    g = _load(golden_dir, tag + ".npz")
    B, H, W, smooth, seed, _ = [int(v) for v in g["meta"]]
    fr = O.synth_frames(6, B, H, W, seed=seed, smooth=bool(smooth))
    outs = O.window_forward(fr, sd)
    assert len(outs) == 14
    for k, o in enumerate(outs):
        assert (o - g[f"out{k}"]).abs().max().item() <= TOL, k


def test_pyramid(golden_dir, sd):
    g = _load(golden_dir, "pyramid.npz")
    fr = O.synth_frames(5, 1, 16, 24, seed=77)
    prev = [t * 2 - 1 for t in O.synth_frames(6, 1, 16, 24, seed=78)]
    msd = O.sub_sd(sd, "model")
    for tag, p in (("none", [None] * 6), ("prev", prev)):
        outs = O.pyramid(fr, p, msd)
        for k, o in enumerate(outs):
            assert (o - g[f"{tag}{k}"]).abs().max().item() <= TOL, (tag, k)


@pytest.mark.parametrize("name,n", [("model1_1", 2), ("model2_1", 3), ("model3_1", 5), ("model4_1", 5)])
def test_backbone(golden_dir, sd, name, n):
    g = _load(golden_dir, f"backbone_{name}.npz")
    fr = O.synth_frames(n, 2, 20, 36, seed=100 + n)
    y = O.backbone(fr, O.sub_sd(sd, "model." + name))
    assert (y - g["out"]).abs().max().item() <= TOL


def test_rdb(golden_dir, sd):
    g = _load(golden_dir, "rdb.npz")
    bsd = O.sub_sd(sd, "model.model2_1")
    y = O.rdb(g["x"], bsd, "RDBs.3")
    assert (y - g["out"]).abs().max().item() <= 1e-5      # inputs are N(0,1): larger magnitudes
    c0 = torch.cat((g["x"], torch.relu(O.conv(g["x"], bsd, "RDBs.3.convs.0.conv.0"))), 1)
    assert (c0 - g["conv0"]).abs().max().item() <= 1e-5


def test_pixel_reshuffle(golden_dir):
    g = _load(golden_dir, "pixel_reshuffle.npz")
    assert torch.equal(O.space_to_depth2(g["x"]), g["out"])


def test_convlstm(golden_dir, sd):
    g = _load(golden_dir, "convlstm.npz")
    h, (c, _) = O.convlstm(g["x"], sd, "clstm_7_prime_prime", None)
    assert (h - g["h_none"]).abs().max().item() <= TOL and (c - g["c_none"]).abs().max().item() <= TOL
    h, (c, _) = O.convlstm(g["x"], sd, "clstm_7_prime_prime", (g["c0"], g["h0"]))
    assert (h - g["h_state"]).abs().max().item() <= TOL and (c - g["c_state"]).abs().max().item() <= TOL


def test_window_grad(golden_dir, sd):
    """Backward of the oracle (autograd through the restatement) against the reference's autograd."""
    g = _load(golden_dir, "window_grad.npz")
    names = [k[2:] for k in g if k.startswith("d:")]
    sd2 = dict(sd)
    leaves = {}
    for n in names:                       # aliases share storage: make the canonical tensor a leaf everywhere
        leaf = sd[n].clone().requires_grad_(True)
        leaves[n] = leaf
        for k, v in sd.items():
            if v.data_ptr() == sd[n].data_ptr():
                sd2[k] = leaf
    fr = [f.requires_grad_(True) for f in O.synth_frames(6, 1, 16, 16, seed=9)]
    outs = O.window_forward(fr, sd2)
    cots = O.synth_frames(14, 1, 16, 16, seed=10)
    loss = sum((o * (c - 0.5)).sum() for o, c in zip(outs, cots))
    assert abs(loss.item() - g["loss"].item()) <= 1e-3 * max(1.0, abs(g["loss"].item()))
    grads = torch.autograd.grad(loss, fr + [leaves[n] for n in names])
    for k in range(6):
        ref = g[f"dframe{k}"]
        assert (grads[k] - ref).abs().max().item() <= 1e-4 * max(1.0, ref.abs().max().item()), k
    for n, gr in zip(names, grads[6:]):
        ref = g["d:" + n]
        assert (gr - ref).abs().max().item() <= 1e-4 * max(1.0, ref.abs().max().item()), n


def test_tensor2img(golden_dir):
    g = np.load(os.path.join(golden_dir, "tensor2img.npz"))
    assert np.array_equal(O.tensor2img_bgr_u8(torch.from_numpy(g["x"])), g["out"])


@pytest.mark.parametrize("kind", ["l1", "l2", "cb"])
def test_get_loss(golden_dir, kind):
    g = _load(golden_dir, "get_loss.npz")
    loss, ll = O.get_loss_6v2(list(g["outs"]), list(g["gts"]), kind)
    assert abs(loss.item() - g[kind].item()) <= 1e-5 * max(1.0, abs(g[kind].item()))
    assert (torch.stack(ll) - g[kind + "_list"]).abs().max().item() <= 1e-4 * max(1.0, g[kind + "_list"].abs().max().item())


@pytest.mark.parametrize("tag", ["wd0", "wd"])
def test_adam(golden_dir, tag):
    """oracle.adam_step vs three steps of torch.optim.Adam (oracle/make_golden_train.py)."""
    d = np.load(os.path.join(golden_dir, "adam.npz"))
    lr, b1, b2, eps, wd = (float(x) for x in d["hyper"])
    wd = wd if tag == "wd" else 0.0
    p = d[f"{tag}_p0"].copy()
    m, v = np.zeros_like(p), np.zeros_like(p)
    for k in range(3):
        p, m, v = O.adam_step(p, d[f"{tag}_grads"][k], m, v, k + 1, lr, b1, b2, eps, wd)
    np.testing.assert_allclose(m, d[f"{tag}_m"], rtol=2e-6, atol=1e-9)
    np.testing.assert_allclose(v, d[f"{tag}_v"], rtol=2e-6, atol=1e-12)
    np.testing.assert_allclose(p, d[f"{tag}_p"], rtol=0, atol=2e-7)      # 3 updates of magnitude lr = 1e-4


@pytest.mark.parametrize("ws", [7, 11])
def test_blur_average(golden_dir, ws):
    """oracle.blur_average vs the reference script's create_clips_overlap, bit-exact."""
    d = np.load(os.path.join(golden_dir, "blur_average.npz"))
    assert list(d[f"ws{ws}_mid"]) == [16, 24, 32, 40]
    out = O.blur_average(d["frames"], window_size=ws)
    assert out.dtype == np.uint8 and np.array_equal(out, d[f"ws{ws}_out"])


@pytest.mark.parametrize("ws", [1, 7, 8, 11, 255])
def test_blur_average_float_formula_equals_integer_floor(ws):
    """The script's float32 sum / float32 count / astype(uint8) (:117-130) equals integer floor division for every
    reachable sum, so the CUDA kernel (integer accumulation + one IEEE division) and the oracle agree for any input;
    even window sizes average 2*int((ws-1)/2)+1 frames (:101, :129)."""
    n = 2 * int((ws - 1) / 2) + 1
    sums = np.arange(0, 255 * n + 1, dtype=np.int64)
    got = (sums.astype(np.float32) / np.float32(n)).astype("uint8")
    assert np.array_equal(got, (sums // n).astype(np.uint8))
    frames = np.random.default_rng(ws).integers(0, 256, size=(2 * (n // 2) + 1, 3, 5), dtype=np.uint8)
    out = O.blur_average(frames, window_size=ws, first_mid=n // 2, stride=1, nwin=1)
    assert np.array_equal(out[0], (frames.astype(np.int64).sum(0) // n).astype(np.uint8))
