"""SURVEY 8f ranks 3-4 on the GPU: the multi-tensor Adam launch and the blur-synthesis kernel, against the fixtures the
reference's own code produced (oracle/make_golden_train.py), against the oracle on other shapes, and -- for Adam --
against torch.optim.Adam stepping the real 540-tensor network."""
import os

import numpy as np
import pytest
import torch

from oracle import bin_oracle as O

pytestmark = pytest.mark.gpu


def _split(flat, sizes):
    out, o = [], 0
    for n in sizes:
        out.append(flat[o:o + n])
        o += n
    return out


@pytest.mark.parametrize("tag", ["wd0", "wd"])
def test_adam_matches_torch_fixture(golden_dir, tag):
    from bin_b200.optim import Adam
    d = np.load(os.path.join(golden_dir, "adam.npz"))
    lr, b1, b2, eps, wd = (float(x) for x in d["hyper"])
    wd = wd if tag == "wd" else 0.0
    sizes = [int(x) for x in d["sizes"]]
    # one flat buffer: the views start at odd element offsets (3, 651, ...) -> exercises the unaligned path too
    flat = torch.from_numpy(d[f"{tag}_p0"].copy()).cuda()
    ps = [torch.nn.Parameter(t) for t in _split(flat, sizes)]
    opt = Adam(ps, lr=lr, betas=(b1, b2), eps=eps, weight_decay=wd)
    for k in range(3):
        for p, g in zip(ps, _split(torch.from_numpy(d[f"{tag}_grads"][k]).cuda(), sizes)):
            p.grad = g
        opt.step()
    torch.cuda.synchronize()
    cat = lambda key: torch.cat([opt.state[p][key].flatten() for p in ps]).cpu().numpy()
    np.testing.assert_allclose(cat("exp_avg"), d[f"{tag}_m"], rtol=2e-6, atol=1e-9)
    np.testing.assert_allclose(cat("exp_avg_sq"), d[f"{tag}_v"], rtol=2e-6, atol=1e-12)
    np.testing.assert_allclose(flat.cpu().numpy(), d[f"{tag}_p"], rtol=0, atol=2e-7)
    assert all(float(opt.state[p]["step"]) == 3.0 for p in ps)


def test_adam_on_the_network_matches_torch_adam():
    """All 540 tensors, two steps, schedulers' lr write and a torch.optim.Adam state_dict round trip."""
    import bin_b200.rdn as RDN
    from bin_b200.optim import Adam
    net_a = RDN.bin_stage4_lstm().cuda()
    net_b = RDN.bin_stage4_lstm().cuda()
    net_b.load_state_dict(net_a.state_dict())
    pa, pb = list(net_a.parameters()), list(net_b.parameters())
    assert len(pa) == 540
    ours = Adam(pa, lr=1e-4, betas=(0.9, 0.99), weight_decay=1e-5)
    ref = torch.optim.Adam(pb, lr=1e-4, betas=(0.9, 0.99), weight_decay=1e-5, foreach=False, fused=False)
    g = torch.Generator(device="cuda").manual_seed(5)
    for step in range(3):                              # step 0 builds the table, 1-2 take the steady-state path
        if step == 1:                                  # what lr_scheduler / bin_model.py:145 do
            ours.param_groups[0]["lr"] = ref.param_groups[0]["lr"] = 5e-5
        for a, b in zip(pa, pb):
            a.grad = torch.randn(a.shape, device="cuda", generator=g) * 1e-2
            b.grad = a.grad.clone()
        ours.step()
        ref.step()
    worst = max(float((a.detach() - b.detach()).abs().max()) for a, b in zip(pa, pb))
    assert worst <= 2e-7, worst
    # state dicts are interchangeable (base_model.save_training_state / resume_training)
    ref2 = torch.optim.Adam(pb, lr=1e-4, betas=(0.9, 0.99))
    ref2.load_state_dict(ours.state_dict())
    assert float(ref2.state[pb[0]]["step"]) == 3.0 and float(ref.state[pb[-1]]["step"]) == 3.0
    ours2 = Adam(pa, lr=1e-4, betas=(0.9, 0.99))
    ours2.load_state_dict(ref.state_dict())
    m_err = max(float((ours2.state[a]["exp_avg"] - ours.state[a]["exp_avg"]).abs().max()) for a in pa)
    assert m_err <= 1e-8, m_err


def test_adam_skips_params_without_grad_and_rejects_cpu():
    from bin_b200 import BinB200Error
    from bin_b200.optim import Adam
    a = torch.nn.Parameter(torch.ones(5000, device="cuda"))
    b = torch.nn.Parameter(torch.ones(7, device="cuda"))
    opt = Adam([a, b], lr=0.1)
    a.grad = torch.full_like(a, 2.0)
    opt.step()
    torch.cuda.synchronize()
    assert torch.all(b == 1) and len(opt.state[b]) == 0
    assert torch.allclose(a, torch.full_like(a, 0.9), atol=1e-6)     # first Adam step moves by lr * sign(g)
    c = torch.nn.Parameter(torch.ones(4))
    c.grad = torch.ones(4)
    with pytest.raises(BinB200Error):
        Adam([c]).step()


@pytest.mark.parametrize("ws", [7, 11])
def test_blur_average_matches_reference_fixture(golden_dir, ws):
    from bin_b200.dataprep import blur_average
    d = np.load(os.path.join(golden_dir, "blur_average.npz"))
    out = blur_average(torch.from_numpy(d["frames"]).cuda(), window_size=ws)
    assert out.dtype == torch.uint8 and np.array_equal(out.cpu().numpy(), d[f"ws{ws}_out"])


@pytest.mark.parametrize("shape,T,ws", [((352, 640, 3), 96, 11), ((5, 7, 3), 40, 7), ((33,), 200, 11), ((16, 16, 3), 24, 1)])
def test_blur_average_matches_oracle(shape, T, ws):
    """Script-size frames (640x352, :25-26 defaults scaled), an odd byte count (scalar path), a long clip, window 1."""
    from bin_b200.dataprep import blur_average, window_count
    rng = np.random.default_rng(T)
    frames = rng.integers(0, 256, size=(T,) + shape, dtype=np.uint8)
    out = blur_average(torch.from_numpy(frames).cuda(), window_size=ws)
    assert out.shape[0] == window_count(T)
    assert np.array_equal(out.cpu().numpy(), O.blur_average(frames, window_size=ws))


def test_blur_average_rejects_short_clips():
    from bin_b200 import BinB200Error
    from bin_b200.dataprep import blur_average
    with pytest.raises(BinB200Error):
        blur_average(torch.zeros((20, 4, 4, 3), dtype=torch.uint8, device="cuda"))      # floor(20/8)-2 = 0 windows
    with pytest.raises(BinB200Error):
        blur_average(torch.zeros((30, 4, 4, 3), dtype=torch.uint8, device="cuda"), nwin=3)   # window 2 needs frame 37
