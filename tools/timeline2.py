import os
os.environ.setdefault("BIN_B200_LIB", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "bin_b200", "libbin_b200_tools.so"))  # tools build: timeline hooks + microbenchmarks
"""In-kernel role timeline (BIN_B200_DEBUG=8) for an arbitrary layer: usage timeline2.py {final|up0|lff|gff0|sfe1|conv3}"""
import ctypes as C, os, sys
os.environ["BIN_B200_DEBUG"] = "8"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from bin_b200 import ops, _lib
dev = "cuda"; B, h, w = 5, 360, 640
which = sys.argv[1] if len(sys.argv) > 1 else "final"
def W_(co, ci, k): return torch.randn(co, ci, k, k, device=dev) / (ci * k * k) ** 0.5
if which == "final":
    H, Wd = 2 * h, 2 * w
    wp, bp = ops.pack_conv_weight(W_(3, 64, 3), 16, 64), torch.zeros(16, device=dev)
    x = torch.randn(B, 8, H, Wd, 8, device=dev).half()
    frames = [[torch.rand(1, 3, H, Wd, device=dev) for _ in range(2)] for _ in range(B)]
    outs = [torch.empty(1, 3, H, Wd, device=dev) for _ in range(B)]
    fr = ops.make_frames(frames, outs)
    run = lambda: ops.conv_fwd(x, wp, bp, 3, 16, epilogue=2, frames=fr)
    units = 2
elif which == "up0":
    wp, bp = ops.pack_conv_weight(W_(256, 96, 3), 256, 96), torch.zeros(256, device=dev)
    x = torch.randn(B, 12, h, w, 8, device=dev).half(); out = ops.empty_p8(B, 8, 2 * h, 2 * w, dev)
    run = lambda: ops.conv_fwd(x, wp, bp, 3, 256, epilogue=1, out=out); units = 3
elif which == "lff":
    wp, bp = ops.pack_conv_weight(W_(96, 224, 1), 96, 224), torch.zeros(96, device=dev)
    x = torch.randn(B, 12, h, w, 8, device=dev).half(); g = torch.randn(B, 16, h, w, 8, device=dev).half(); out = ops.empty_p8(B, 12, h, w, dev)
    run = lambda: ops.conv_fwd(x, wp, bp, 1, 96, in0_planes=12, in1=g, in1_planes=16, out=out, res=x); units = 3
elif which == "gff0":
    wp, bp = ops.pack_conv_weight(W_(96, 1152, 1), 96, 1152), torch.zeros(96, device=dev)
    x = torch.randn(B, 144, h, w, 8, device=dev).half(); out = ops.empty_p8(B, 12, h, w, dev)
    run = lambda: ops.conv_fwd(x, wp, bp, 1, 96, out=out); units = 12
else:
    raise SystemExit("unknown layer")
for _ in range(3):
    run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); run(); e1.record(); torch.cuda.synchronize()
print(which, "kernel ms", e0.elapsed_time(e1))
buf = (C.c_longlong * (3 * 4096))()
_lib.check(_lib.lib().bin_tools_debug_timeline(buf, 3 * 4096))
a = list(buf)
def role(r, n, k): return [[a[r * 4096 + i * 4 + j] for j in range(k)] for i in range(n)]
mma = role(1, 40, 3); prod = role(0, 40, 2); epi = role(2, 40, 3)
t0 = min(v for v in (prod[0][0], mma[0][0], epi[0][0]) if v)
for i in range(20, 32): print("mmaA stage", i, [v - t0 for v in mma[i]], "wait=%d issue=%d" % (mma[i][1] - mma[i][0], mma[i][2] - mma[i][1]))
for i in range(20, 28): print("prodA", i, [v - t0 for v in prod[i]], "wait=%d" % (prod[i][1] - prod[i][0]))
for i in range(20, 30): print("epi tile", i, [v - t0 for v in epi[i]], "wait=%d work=%d" % (epi[i][1] - epi[i][0], epi[i][2] - epi[i][1]))
print("per tile cycles:", (epi[38][2] - epi[18][2]) / 20.0)
