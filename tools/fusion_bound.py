"""The experiment behind DESIGN.md's decision NOT to fuse RDB convs 0..2 into one kernel.

A fused conv0..2 kernel keeps g0, g1 in shared memory, so per RDB it saves the HBM round trips of the growth maps -- but
it pays a halo: for an 8 x 30 output tile it must compute g0 on 12 x 34 and g1 on 10 x 32 pixels (chained 3x3 convs, UMMA
M tiles of 4 rows), i.e. 1.29x the MMAs, on 208 instead of 240 valid pixels per tile: 1.49x MMA work per output.
Its speed is therefore bounded above by (speed of the layer-by-layer convs when HBM is taken out of the picture) / 1.49.
This tool measures that bound: the three x-stacked convs on a tensor small enough to stay in the 126 MB L2 (warm, no
flush) against the same convs HBM-cold (L2 flushed between launches) and at the bench shape (5 x 360 x 640).
usage: python tools/fusion_bound.py"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bin_b200 import ops  # noqa: E402


def chain(B, h, w, flush, reps=20):
    dev = "cuda"
    x = torch.randn(B, 12, h, w, 8, device=dev).half()
    g = torch.randn(B, 16, h, w, 8, device=dev).half()
    ws = []
    for c in range(3):
        cin = 96 + 32 * c
        ws.append((ops.pack_conv_weight(torch.randn(32, cin, 3, 3, device=dev) / (cin * 9) ** 0.5, 32, cin),
                   ops.pad_bias(torch.zeros(32, device=dev), 32)))
    fl = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device=dev) if flush else None

    def run():
        for c in range(3):
            ops.conv_fwd(x, ws[c][0], ws[c][1], 3, 32, in0_planes=12, in1=g, in1_planes=4 * c, relu=True, out=g, out_plane0=4 * c)
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        if fl is not None:
            fl.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    ms = ts[len(ts) // 2]
    flops = sum(2.0 * B * h * w * (96 + 32 * c) * 32 * 9 for c in range(3))
    return {"B": B, "h": h, "w": w, "working_set_MB": round(B * h * w * 448 / 1e6, 1), "l2_flush": flush, "ms": round(ms, 4),
            "tflops": round(flops / ms / 1e9, 1)}


if __name__ == "__main__":
    res = {"pair": os.environ.get("BIN_B200_PAIR", "0"),
           "l2_resident": chain(1, 192, 640, False), "same_shape_hbm_cold": chain(1, 192, 640, True),
           "bench_shape": chain(5, 360, 640, False)}
    r = res["l2_resident"]["tflops"]
    res["fused_conv0_2_upper_bound_tflops"] = round(r / 1.49, 1)
    res["verdict"] = ("fusing convs 0..2 cannot beat the layer-by-layer kernels" if r / 1.49 <= res["bench_shape"]["tflops"]
                      else "a fused kernel could win by at most x%.2f" % (r / 1.49 / res["bench_shape"]["tflops"]))
    print(json.dumps(res))
