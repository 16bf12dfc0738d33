"""sliding-window reuse (13 backbone calls per window, bit-identical outputs) and the uint8 image
boundary kernels against the oracle's restatement of read_image / ReplicationPad2d / tensor2img."""
import numpy as np
import pytest
import torch

from oracle import bin_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def net():
    from bin_b200 import rdn
    m = rdn.bin_stage4_lstm()
    m.load_state_dict(O.synth_state_dict(0), strict=True)
    return m.cuda().eval()


def test_streaming_is_bit_identical_and_saves_calls(net):
    from bin_b200.streaming import StreamingBIN
    video = [f.cuda() for f in O.synth_frames(9, 1, 48, 80, seed=77, smooth=True)]
    st = StreamingBIN(net)
    got = [st.push(f) for f in video]
    assert all(g is None for g in got[:5]) and all(g is not None for g in got[5:])
    with torch.no_grad():
        for k in range(4):                                      # windows 0..3 = frames k..k+5
            ref = net(*video[k:k + 6])
            assert all(torch.equal(a, b) for a, b in zip(got[5 + k], ref)), k
    assert st.backbone_calls == 17 + 3 * 13                      # first window 17, then 13 per window


@pytest.mark.parametrize("h,w", [(37, 53), (128, 256), (720 // 8, 1280 // 8)])
def test_u8_frame_roundtrip_matches_reference_semantics(h, w):
    from bin_b200.streaming import tensor2img_u8, test_py_padding, upload_frame_u8
    g = np.random.default_rng(5)
    img = g.integers(0, 256, size=(h, w, 3), dtype=np.uint8)
    pad = test_py_padding(h, w)
    pl, pr, pt, pb = pad
    ref = torch.nn.ReplicationPad2d([pl, pr, pt, pb])(O.read_image_u8(img).unsqueeze(0))       # test.py:44-56, 366-371
    got = upload_frame_u8(torch.from_numpy(img), pad, "cuda")
    assert got.shape == ref.shape and torch.equal(got.cpu(), ref)
    # tensor2img incl. clamp and round-half-even, on values that exercise both
    t = (torch.from_numpy(g.uniform(-0.2, 1.2, size=(3, h + pt + pb, w + pl + pr)).astype(np.float32)))
    t.view(-1)[:6] = torch.tensor([0.5 / 255, 1.5 / 255, 2.5 / 255, 254.5 / 255, 1.0, 0.0])
    ref_img = O.tensor2img_bgr_u8(t)[pt:pt + h, pl:pl + w, :]                               # test.py:394-396 crop
    got_img = tensor2img_u8(t.cuda().unsqueeze(0), crop=(pt, pl, h, w)).cpu().numpy()
    assert got_img.shape == ref_img.shape and np.array_equal(got_img, ref_img)
