"""compute-sanitizer target for the CTA-pair kernels, one case per process:
  tail1  fused tail, 4 tiles (one tile pair per cluster)      tailN  fused tail, 220 tiles on 74 clusters (several per cluster, both streams)
  conv3  x-stacked conv with 6 resident chunks                convN  x-stacked conv, 3 chunks, 5 tiles per cluster
usage: BIN_B200_PAIR=1 compute-sanitizer --tool memcheck python tools/sanitize_pair.py <case>"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bin_b200 import ops  # noqa: E402

case = sys.argv[1]
gen = torch.Generator(device="cuda").manual_seed(0)
rnd = lambda *sh: torch.randn(*sh, device="cuda", generator=gen)
B, H, W = {"tail1": (1, 6, 45), "tailN": (2, 41, 290), "conv3": (1, 9, 37), "convN": (2, 100, 300)}[case]
x, g = rnd(B, 12, H, W, 8).half(), rnd(B, 16, H, W, 8).half()
if case.startswith("tail"):
    w3, wl = rnd(32, 192, 3, 3) / 1728 ** 0.5, rnd(96, 224, 1, 1) / 224 ** 0.5
    b3, bl = ops.pad_bias(rnd(32) * 0.1, 32), ops.pad_bias(rnd(96) * 0.1, 96)
    p3, pl = ops.pack_conv_weight(w3, 32, 192), ops.pack_conv_weight(wl, 96, 224)
    out = torch.zeros(B, 12, H, W, 8, device="cuda").half()
    ops.rdb_tail_fwd(x, g, p3, b3, pl, bl, out)
    torch.cuda.synchronize()
    print(case, "ran; finite:", bool(torch.isfinite(out.float()).all()), flush=True)
else:
    nch = 6 if case == "conv3" else 3
    cin = 32 * nch
    w = rnd(32, cin, 3, 3) / (cin * 9) ** 0.5
    wp, bp = ops.pack_conv_weight(w, 32, cin), ops.pad_bias(rnd(32) * 0.1, 32)
    gg = g.clone()
    ops.conv_fwd(x, wp, bp, 3, 32, in0_planes=12, in1=g, in1_planes=4 * (nch - 3), relu=True, out=gg, out_plane0=12)
    torch.cuda.synchronize()
    print(case, "ran; finite:", bool(torch.isfinite(gg.float()).all()), flush=True)
