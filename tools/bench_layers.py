"""Per-layer timing of the conv kernel at 720p shapes (CUDA events, warm-up, L2 flush between
iterations).  Prints one JSON line per layer: ms, TFLOP/s, fraction of the measured bf16 peak."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bin_b200 import ops  # noqa: E402

PEAK = 1664.6e12
try:
    PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops"] * 1e12
except Exception:
    pass


def bench(name, cin, cout, k, H, W, B=1, epilogue=0, relu=False, res=False, split=None, iters=10, variant=0):
    dev = "cuda"
    cin_pad = (cin + 31) // 32 * 32
    cout_pad = 16 if cout == 3 else cout
    w = torch.randn(cout, cin, k, k, device=dev) / (cin * k * k) ** 0.5
    b = torch.zeros(cout, device=dev)
    wp, bp = ops.pack_conv_weight(w, cout_pad, cin_pad, variant), ops.pad_bias(b, cout_pad)
    if split is None:
        in0 = torch.randn(B, cin_pad // 8, H, W, 8, device=dev).half()
        in1, p0, p1 = None, cin_pad // 8, 0
    else:
        in0 = torch.randn(B, split // 8, H, W, 8, device=dev).half()
        in1 = torch.randn(B, (cin - split) // 8, H, W, 8, device=dev).half()
        p0, p1 = split // 8, (cin - split) // 8
    kw = dict(in0_planes=p0, in1=in1, in1_planes=p1, relu=relu, epilogue=epilogue, variant=variant)
    if epilogue == 0:
        kw["out"] = ops.empty_p8(B, cout_pad // 8, H, W, dev)
        if res:
            kw["res"] = in0
    elif epilogue == 1:
        kw["out"] = ops.empty_p8(B, 8, 2 * H, 2 * W, dev)
    else:
        frames = [[torch.rand(B, 3, H, W, device=dev) for _ in range(2)]]
        kw["frames"] = ops.make_frames(frames, [torch.empty(B, 3, H, W, device=dev)])
        kw["_keep"] = frames
    keep = kw.pop("_keep", None)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    for _ in range(3):
        ops.conv_fwd(in0, wp, bp, k, cout_pad, **kw)
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.conv_fwd(in0, wp, bp, k, cout_pad, **kw)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    ms = ts[len(ts) // 2]
    flops = 2.0 * B * H * W * cin * cout * k * k
    print(json.dumps({"layer": name, "ms": round(ms, 4), "tflops": round(flops / ms / 1e9, 1),
                      "frac_peak": round(flops / (ms * 1e-3) / PEAK, 4), "min_ms": round(ts[0], 4)}), flush=True)
    return ms


if __name__ == "__main__":
    h, w = 360, 640
    if len(sys.argv) > 1:
        B = int(sys.argv[1])
        _bench = bench
        bench = lambda *a, **k: _bench(*a, **dict(k, B=B))
    tot = 0.0
    t = {}
    t["sfe1_24"] = bench("SFENet1 5x5 24->96", 24, 96, 5, h, w)
    t["sfe2"] = bench("SFENet2 3x3 96->96", 96, 96, 3, h, w)
    for c in range(4):
        bench(f"RDB conv{c} PLAIN 3x3 {96+32*c}->32", 96 + 32 * c, 32, 3, h, w, relu=True, split=96 if c else None, variant=1)
        t[f"rdb{c}"] = bench(f"RDB conv{c} 3x3 {96+32*c}->32", 96 + 32 * c, 32, 3, h, w, relu=True, split=96 if c else None)
    t["lff"] = bench("LFF 1x1 224->96", 224, 96, 1, h, w, res=True, split=96)
    t["gff0"] = bench("GFF.0 1x1 1152->96", 1152, 96, 1, h, w)
    t["gff1"] = bench("GFF.1 3x3 96->96 +res", 96, 96, 3, h, w, res=True)
    t["up0"] = bench("UPNet.0 3x3 96->256 pixshuf", 96, 256, 3, h, w, epilogue=1)
    t["up2"] = bench("UPNet.2 3x3 64->3 final", 64, 3, 3, 2 * h, 2 * w, epilogue=2)
    bb = t["sfe1_24"] + t["sfe2"] + 12 * (t["rdb0"] + t["rdb1"] + t["rdb2"] + t["rdb3"] + t["lff"]) + t["gff0"] + t["gff1"] + t["up0"] + t["up2"]
    macs = 702720 * 4 * h * w
    print(json.dumps({"backbone_2frame_est_ms": round(bb, 3), "tflops": round(2 * macs / bb / 1e9, 1),
                      "frac_peak": round(2 * macs / (bb * 1e-3) / PEAK, 4), "window_est_ms": round(bb * 17.3, 1)}))
