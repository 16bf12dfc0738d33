"""A/B of the CTA-pair (cta_group::2) fused RDB tail against the single-CTA kernel at the bench shape (5 x 360 x 640),
each in its own process (the library reads BIN_B200_PAIR once).  usage: python tools/ab_pair.py [child pair]"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

if len(sys.argv) > 1 and sys.argv[1] == "child":
    import torch
    from bin_b200 import ops
    B, h, w = 5, 360, 640
    gen = torch.Generator(device="cuda").manual_seed(0)
    rnd = lambda *sh: torch.randn(*sh, device="cuda", generator=gen)
    x, g = rnd(B, 12, h, w, 8).half(), rnd(B, 16, h, w, 8).half()
    w3, wl = rnd(32, 192, 3, 3) / 1728 ** 0.5, rnd(96, 224, 1, 1) / 224 ** 0.5
    b3, bl = ops.pad_bias(rnd(32) * 0.1, 32), ops.pad_bias(rnd(96) * 0.1, 96)
    p3, pl = ops.pack_conv_weight(w3, 32, 192), ops.pack_conv_weight(wl, 96, 224)
    out = torch.zeros(B, 12, h, w, 8, device="cuda").half()
    run = lambda: ops.rdb_tail_fwd(x, g, p3, b3, pl, bl, out)
    for _ in range(5):
        run()
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            run()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) / 20)
    # reference result: layer by layer
    g_ref, out_ref = g.clone(), torch.zeros_like(out)
    ops.conv_fwd(x, p3, b3, 3, 32, in0_planes=12, in1=g_ref, in1_planes=12, relu=True, out=g_ref, out_plane0=12)
    ops.conv_fwd(x, pl, bl, 1, 96, in0_planes=12, in1=g_ref, in1_planes=16, out=out_ref, res=x)
    torch.cuda.synchronize()
    ms = sorted(ts)[len(ts) // 2]
    nbytes = B * h * w * 576
    print(json.dumps({"pair": os.environ.get("BIN_B200_PAIR", "0"), "tailq": os.environ.get("BIN_B200_TAILQ", "0"), "ms": round(ms, 4), "min_ms": round(min(ts), 4),
                      "GBps_algorithmic": round(nbytes / ms / 1e6, 1), "bit_identical_to_layerwise": bool(torch.equal(out, out_ref)),
                      "max_abs_diff": (out.float() - out_ref.float()).abs().max().item()}))
else:
    for cfg in ({"BIN_B200_PAIR": "0", "BIN_B200_TAILQ": "0"}, {"BIN_B200_PAIR": "0", "BIN_B200_TAILQ": "1"}, {"BIN_B200_PAIR": "1"},
                {"BIN_B200_PAIR": "0", "BIN_B200_TAILQ": "0"}, {"BIN_B200_PAIR": "0", "BIN_B200_TAILQ": "1"}):
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "child"], env=dict(os.environ, **cfg),
                           capture_output=True, text=True, timeout=300)
        print(r.stdout.strip() or r.stderr[-1500:], flush=True)
