"""Parity on the configurations the bench is quoted on: 256x256 against the fp32 CPU
oracle, 1280x720 and the 768x1344 test.py pads it to against the oracle executed ON THE SAME GPU in fp64 (cuDNN /
native double convolutions: no TF32, no fp16 -- the fp32 CPU oracle itself is 4e-7 from fp64), the reference callers'
sequence (DataParallel wrap -> test_set_input -> test_forward -> tensor2img crop) at 720p, a batch-2 256x256 backward
with the oracle storing activations in fp16 like the CUDA path (emulate_fp16_storage), and a 2-device DataParallel run.

Tolerances (north_star): fp16 mode <= 1e-3 max-abs and <= 0.01 dB PSNR; fp32 mode <= 1e-5."""
import pytest
import torch

from oracle import bin_oracle as O

pytestmark = pytest.mark.gpu
TOL_FP16, TOL_FP32_MODE, TOL_PSNR = 1e-3, 1e-5, 0.01


@pytest.fixture(scope="module")
def sd():
    return O.synth_state_dict(0)


@pytest.fixture(scope="module")
def net(sd):
    from bin_b200 import rdn
    m = rdn.bin_stage4_lstm()
    m.load_state_dict(sd, strict=True)
    return m.cuda().eval()


@pytest.fixture(scope="module")
def net32(sd):
    from bin_b200 import rdn
    m = rdn.bin_stage4_lstm()
    m.load_state_dict(sd, strict=True)
    return rdn.set_precision(m.cuda().eval(), "fp32")


def oracle_on_gpu(frames, sd, dtype=torch.float64):
    """The oracle's own window_forward with every tensor on cuda:0 in `dtype` (fp64: exact to ~1e-15 relative)."""
    tf32 = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        with torch.no_grad():
            sdd = {k: v.to("cuda", dtype) for k, v in sd.items()}
            out = O.window_forward([f.to("cuda", dtype) for f in frames], sdd)
            return [o.float() for o in out]
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf32


def _psnr_delta(got, ref, gt):
    p_ref = O.psnr_u8(O.tensor2img_u8(ref), O.tensor2img_u8(gt))
    p_got = O.psnr_u8(O.tensor2img_u8(got), O.tensor2img_u8(gt))
    return abs(p_ref - p_got)


def test_gpu_fp64_oracle_equals_cpu_oracle(sd):
    """The full-size tests trust the oracle run in fp64 on the GPU: pin it to the fp32 CPU oracle at a small size."""
    fr = O.synth_frames(6, 1, 64, 96, seed=1234, smooth=True)
    cpu = O.window_forward(fr, sd)
    gpu = oracle_on_gpu(fr, sd)
    assert max((a.cpu() - b).abs().max().item() for a, b in zip(gpu, cpu)) <= 2e-6


def test_window_256_vs_cpu_oracle(net, net32, sd):
    """BASELINE config-3 spatial size (256x256) against the fp32 CPU oracle, both precisions, PSNR delta."""
    torch.set_num_threads(min(16, torch.get_num_threads()))
    fr = O.synth_frames(6, 1, 256, 256, seed=1234, smooth=True)
    gt = O.synth_frames(14, 1, 256, 256, seed=4321, smooth=True)
    ref = O.window_forward(fr, sd)
    with torch.no_grad():
        outs = [o.cpu() for o in net(*[f.cuda() for f in fr])]
        outs32 = [o.cpu() for o in net32(*[f.cuda() for f in fr])]
    for k in range(14):
        assert (outs[k] - ref[k]).abs().max().item() <= TOL_FP16, k
        assert (outs32[k] - ref[k]).abs().max().item() <= TOL_FP32_MODE, k
        assert _psnr_delta(outs[k], ref[k], gt[k]) <= TOL_PSNR, k


@pytest.mark.parametrize("H,W", [(720, 1280), (768, 1344)])
def test_window_fullsize_vs_gpu_oracle(net, net32, sd, H, W):
    """The benchmarked configuration (test.py:348-372: 1280x720 and its 768x1344 padding): persistent CTAs with many
    tiles each, 5-6 batched calls per stage at 360x640 / 384x672 -- against the fp64 oracle on the same GPU."""
    fr = O.synth_frames(6, 1, H, W, seed=1234, smooth=True)
    gt = O.synth_frames(14, 1, H, W, seed=4321, smooth=True)
    ref = oracle_on_gpu(fr, sd)
    frc = [f.cuda() for f in fr]
    with torch.no_grad():
        outs = net(*frc)
        worst16 = max((o - r).abs().max().item() for o, r in zip(outs, ref))
        psnr = max(_psnr_delta(o.cpu(), r.cpu(), g) for o, r, g in zip(outs, ref, gt))
        del outs
        outs32 = net32(*frc)
        worst32 = max((o - r).abs().max().item() for o, r in zip(outs32, ref))
    print(f"[fullsize {W}x{H}] fp16 mode max-abs {worst16:.3e}  PSNR delta {psnr:.4f} dB  fp32 mode max-abs {worst32:.3e}")
    assert worst16 <= TOL_FP16, worst16
    assert psnr <= TOL_PSNR, psnr
    assert worst32 <= TOL_FP32_MODE, worst32


def test_reference_caller_sequence_720p(sd):
    """bin_model.__init__ (DataParallel wrap) -> load_network -> test_set_input -> test_forward -> tensor2img + crop
    exactly as test.py:334-402 drives them, on a 1280x720 window (padded to 768x1344 by the caller); the three images
    test.py writes must equal the oracle's within one uint8 level, the tensors within 1e-3."""
    import numpy as np
    from caller_harness import CallerModel, pad_like_test_py, run_test_py_window, tensor2img
    from bin_b200 import rdn
    model = CallerModel(rdn.bin_stage4_lstm(), "cuda:0", device_ids=[0])
    model.load_state_dict_like_load_network({"InterpNet." + k: v for k, v in sd.items()})
    frames = [f[0] for f in O.synth_frames(6, 1, 720, 1280, seed=77, smooth=True)]        # (3,H,W) like read_image
    imgs, Ft_p, (pl, pr, pt, pb) = run_test_py_window(model, frames)
    assert len(Ft_p) == 14 and Ft_p[13].shape == (1, 3, 768, 1344)
    padded, _ = pad_like_test_py(frames)
    ref = oracle_on_gpu(padded, sd)
    for img, k in zip(imgs, (13, 8, 12)):
        assert (Ft_p[k] - ref[k]).abs().max().item() <= TOL_FP16, k
        want = tensor2img(ref[k].squeeze(0))[pt:pt + 720, pl:pl + 1280, :]
        assert img.shape == (720, 1280, 3) and img.dtype == np.uint8
        d = np.abs(img.astype(np.int16) - want.astype(np.int16))
        assert d.max() <= 1 and (d != 0).mean() <= 0.02, (k, int(d.max()), float((d != 0).mean()))
    # second window through the same wrapper (demo.py / test.py loop): inputs untouched, result deterministic
    imgs2, _, _ = run_test_py_window(model, frames)
    assert all(np.array_equal(a, b) for a, b in zip(imgs, imgs2))


def test_adam_step_is_seen_by_the_next_forward(sd):
    """bin_b200.optim.Adam writes parameters through raw pointers; the packed-weight caches are keyed
    on (data_ptr, _version), so step() must bump the versions -- forward/backward/step/forward must track a
    torch.optim.Adam run on the same gradients, eager and graphed."""
    from bin_b200 import rdn
    from bin_b200.optim import Adam
    a = rdn.bin_stage4_lstm(); a.load_state_dict(sd, strict=True); a = a.cuda()
    b = rdn.bin_stage4_lstm(); b.load_state_dict(sd, strict=True); b = b.cuda()
    fr = [f.cuda() for f in O.synth_frames(6, 1, 32, 48, seed=3, smooth=True)]
    gt = [g.cuda() for g in O.synth_frames(14, 1, 32, 48, seed=4, smooth=True)]
    oa = Adam(a.parameters(), lr=1e-4, betas=(0.9, 0.99))
    ob = torch.optim.Adam(b.parameters(), lr=1e-4, betas=(0.9, 0.99))
    losses = {"a": [], "b": []}
    for it in range(3):
        for tag, net_, opt in (("a", a, oa), ("b", b, ob)):
            net_.train()
            opt.zero_grad(set_to_none=True)
            loss = sum((o - g).abs().sum() for o, g in zip(net_(*fr), gt)) / 14
            loss.backward()
            opt.step()
            losses[tag].append(loss.item())
    assert losses["a"][2] < losses["a"][0]                                # it learns: the weights the net runs on DO move
    for x, y in zip(losses["a"], losses["b"]):
        assert abs(x - y) <= 2e-3 * abs(y), (losses["a"], losses["b"])
    a.eval(); b.eval()
    with torch.no_grad():
        ya, yb = a(*fr), b(*fr)                                           # graphed inference after the steps
    assert max((p - q).abs().max().item() for p, q in zip(ya, yb)) <= 2e-3
    # Adam normalises every gradient to a +-lr step: an element whose (tiny) gradient differs in sign between the two
    # runs (atomics in the bias gradients are not bit-reproducible) moves 2 lr apart per step -- bound: 3 steps x 2 lr
    for p, q in zip(a.parameters(), b.parameters()):
        d = (p - q).abs()
        assert d.max().item() <= 6.5e-4 and d.mean().item() <= 5e-6


def _oracle_window_grads_gpu(fr, cots, sd, dtype=torch.float32):
    """Autograd through the oracle on the GPU with fp16-rounded storage (what the CUDA path keeps in HBM)."""
    tf32 = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        uniq, leaves = {}, {}
        for k, v in sd.items():                                           # aliases share one leaf
            key = v.data_ptr()
            if key not in uniq:
                uniq[key] = v.to("cuda", dtype).requires_grad_(True)
            leaves[k] = uniq[key]
        frg = [f.to("cuda", dtype).requires_grad_(True) for f in fr]
        with O.emulate_fp16_storage(grads=True):
            outs = O.window_forward(frg, leaves)
        loss = sum((o * c.to("cuda", dtype)).sum() for o, c in zip(outs, cots))
        names = list(dict.fromkeys(k for k in sd))
        first = {}
        for k in names:
            first.setdefault(leaves[k], k)
        plist = list(first.keys())
        grads = torch.autograd.grad(loss, frg + plist, allow_unused=True)
        return ([g.float() for g in grads[:6]], {first[p]: g for p, g in zip(plist, grads[6:]) if g is not None},
                [o.detach().float() for o in outs])
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf32


def test_window_backward_b2_256_vs_fp16_storage_oracle(sd):
    """BASELINE config 3 geometry (256x256 crops, batch 2 here to bound the oracle's memory): gradients of
    sum_k <out_k, cot_k> w.r.t. the 6 frames and all 540 parameter tensors against autograd through the oracle with
    fp16-rounded storage of activations AND gradients (what the CUDA path keeps in HBM).

    A network with fp16-stored activations is chaotic at the rounding boundaries: the SAME oracle evaluated with fp32
    and with fp64 accumulation already differs by ~4e-4 in its outputs (a value next to an fp16 rounding midpoint or a
    ReLU threshold lands on the other side) and by several per cent of the tensor maximum in individual weight
    gradients.  That oracle-vs-oracle discrepancy is the resolution of this comparison, so it is measured here and the
    CUDA path is required to be as close to the oracle as the oracle is to itself: per backbone, worst error
    <= 2 x the oracle's own worst discrepancy + 2 % of the tensor maximum, and correlation >= 0.997 for every tensor."""
    from bin_b200 import rdn
    B, H, W = 2, 256, 256
    fr = O.synth_frames(6, B, H, W, seed=9, smooth=True)
    cots = [c - 0.5 for c in O.synth_frames(14, B, H, W, seed=10)]
    gfr, gp, ref_outs = _oracle_window_grads_gpu(fr, cots, sd, torch.float64)
    gfr32, gp32, ref_outs32 = _oracle_window_grads_gpu(fr, cots, sd, torch.float32)
    self_fwd = max((a - b).abs().max().item() for a, b in zip(ref_outs, ref_outs32))
    net = rdn.bin_stage4_lstm(); net.load_state_dict(sd, strict=True); net = net.cuda().train()
    frg = [f.cuda().requires_grad_(True) for f in fr]
    outs = net(*frg)
    fwd = max((o.detach() - r).abs().max().item() for o, r in zip(outs, ref_outs))
    print(f"[bwd 2x256x256] forward max-abs: CUDA vs fp16-storage oracle {fwd:.3e}; that oracle fp32 vs fp64 accumulate {self_fwd:.3e}")
    assert fwd <= TOL_FP16
    sum((o * c.cuda()).sum() for o, c in zip(outs, cots)).backward()
    for k in range(6):
        err = (frg[k].grad - gfr[k]).abs().max().item() / gfr[k].abs().max().item()
        assert err <= 1e-2, ("frame", k, err)
    params = dict(net.named_parameters())
    rel = lambda a, b: (a.float() - b.float()).abs().max().item() / max(b.abs().max().item(), 1e-20)
    worst = {}
    for key, ref in gp.items():
        got = params[key].grad
        assert got is not None, key
        corr = torch.corrcoef(torch.stack([got.flatten().float(), ref.flatten().float()]))[0, 1].item() if got.numel() > 2 else 1.0
        grp = key.split(".")[1] if key.startswith("model.") else "clstm"
        worst.setdefault(grp, []).append((rel(got, ref), rel(gp32[key], ref), corr, key))
    for grp, rows in sorted(worst.items()):
        ours, floor = max(r[0] for r in rows), max(r[1] for r in rows)
        med = sorted(r[0] for r in rows)[len(rows) // 2]
        medf = sorted(r[1] for r in rows)[len(rows) // 2]
        print(f"[bwd 2x256x256] {grp}: CUDA vs oracle worst {ours:.4f} median {med:.4f} | oracle fp32 vs fp64 worst {floor:.4f} median {medf:.4f} "
              f"| min corr {min(r[2] for r in rows):.5f}")
        assert ours <= 2.0 * floor + 0.02, (grp, sorted(rows, reverse=True)[:4])
        assert min(r[2] for r in rows) >= 0.997, (grp, sorted(rows, key=lambda r: r[2])[:4])


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (runs on the multi-GPU box)")
def test_dataparallel_two_devices(sd):
    """bin_model.py:42: nn.DataParallel(netG) with more than one visible device scatters the batch and runs REPLICAS on
    worker threads / other devices.  Forward (no grad) must equal the single-device result bit for bit; a training call
    must route gradients back to the master parameters."""
    from bin_b200 import rdn
    net = rdn.bin_stage4_lstm(); net.load_state_dict(sd, strict=True); net = net.cuda(0)
    dp = torch.nn.DataParallel(net, device_ids=[0, 1])
    fr = [f.cuda(0) for f in O.synth_frames(6, 2, 48, 64, seed=5, smooth=True)]
    net.eval()
    with torch.no_grad():
        want = net(*fr)
        got = dp(*fr)
    assert len(got) == 14 and all(g.device.index == 0 for g in got)
    assert all(torch.equal(a, b) for a, b in zip(got, want))
    net.train()
    net.zero_grad(set_to_none=True)
    outs = dp(*fr)
    sum(o.sum() for o in outs).backward()
    g_dp = {k: p.grad.clone() for k, p in net.named_parameters()}
    net.zero_grad(set_to_none=True)
    sum(o.sum() for o in net(*fr)).backward()
    for k, p in net.named_parameters():
        ref = p.grad
        assert g_dp[k] is not None and (g_dp[k] - ref).abs().max().item() <= 2e-2 * max(ref.abs().max().item(), 1e-12), k


@pytest.mark.parametrize("kind", ["seed3", "trained_like"])
def test_window_other_weight_distributions(kind):
    """Every other test uses synth_state_dict(0) (U(+-1/sqrt(fan_in))).  Here: another seed, and a 'trained-like' set --
    heavier-tailed weights (normal, 1.15x the default scale, a few output channels boosted 4x, biases up to +-0.3) with
    hard-edged inputs that touch 0 and 1 -- a net whose four chained stages AMPLIFY (outputs reach ~150, hidden maps more;
    the default init contracts) to probe the fp16 storage range.  The
    bar scales with the output magnitude: max-abs <= 1e-3 * max|ref| for seed 3 (outputs ~1: the north_star bar itself) and
    5e-3 * max|ref| for the amplifying trained-like set (measured 2.5e-3: rounding noise grows with the gain of the four
    chained stages; the point of the case is that nothing overflows or degrades disproportionately)."""
    from bin_b200 import rdn
    if kind == "seed3":
        sd = O.synth_state_dict(3)
    else:
        sd = O.synth_state_dict(5)
        gen = torch.Generator().manual_seed(99)
        seen = {}
        for k in list(sd.keys()):
            t = sd[k]
            if t.data_ptr() in seen:                      # aliases keep sharing storage
                sd[k] = seen[t.data_ptr()]
                continue
            if k.endswith("weight") and t.dim() == 4 and "Gates" not in k:
                fan_in = t.shape[1] * t.shape[2] * t.shape[3]
                w = torch.randn(t.shape, generator=gen) * (1.15 / (3.0 * fan_in) ** 0.5)     # std = 1.15 x the uniform's
                boost = torch.randperm(t.shape[0], generator=gen)[: max(1, t.shape[0] // 24)]
                w[boost] *= 4.0
                new = w
            elif k.endswith("bias") and "Gates" not in k:
                new = (torch.rand(t.shape, generator=gen) - 0.5) * 0.6
            else:
                new = t
            seen[t.data_ptr()] = new
            sd[k] = new
    net = rdn.bin_stage4_lstm()
    net.load_state_dict(sd, strict=True)
    net = net.cuda().eval()
    fr = O.synth_frames(6, 1, 64, 96, seed=31, smooth=True)
    fr = [(f > 0.5).float() * 0.75 + f * 0.25 for f in fr]              # hard edges, values in [0, 1]
    ref = O.window_forward(fr, sd)
    with torch.no_grad():
        outs = [o.cpu() for o in net(*[f.cuda() for f in fr])]
    scale = max(1.0, max(r.abs().max().item() for r in ref))
    worst = max((o - r).abs().max().item() for o, r in zip(outs, ref))
    print(f"[weights {kind}] max|ref| {scale:.2f}  max-abs err {worst:.3e}")
    assert all(torch.isfinite(o).all() for o in outs)
    assert worst <= (TOL_FP16 if kind == "seed3" else 5 * TOL_FP16) * scale, (worst, scale)
