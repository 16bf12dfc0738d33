import os
os.environ.setdefault("BIN_B200_LIB", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "bin_b200", "libbin_b200_tools.so"))  # tools build: timeline hooks + microbenchmarks
import ctypes as C, json, os, sys
os.environ["BIN_B200_DEBUG"] = os.environ.get("BIN_B200_DEBUG", "8")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from bin_b200 import ops, _lib
dev = "cuda"; B, h, w = 5, 360, 640
cin = int(sys.argv[1]) if len(sys.argv) > 1 else 96
wt = torch.randn(32, cin, 3, 3, device=dev) / (cin * 9) ** 0.5
wp, bp = ops.pack_conv_weight(wt, 32, cin), ops.pad_bias(torch.zeros(32, device=dev), 32)
x = torch.randn(B, 12, h, w, 8, device=dev).half(); g = torch.randn(B, 16, h, w, 8, device=dev).half()
kw = dict(in0_planes=12, in1=g, in1_planes=(cin - 96) // 8, relu=True, out=g, out_plane0=12)
for _ in range(3):
    ops.conv_fwd(x, wp, bp, 3, 32, **kw)
torch.cuda.synchronize()
buf = (C.c_longlong * (3 * 4096))()
_lib.check(_lib.lib().bin_tools_debug_timeline(buf, 3 * 4096))
a = list(buf)
def role(r, n, k): return [[a[r * 4096 + i * 4 + j] for j in range(k)] for i in range(n)]
nch = cin // 32
mma = role(1, 17 * nch, 3); prod = role(0, 17 * nch, 2); epi = role(2, 34, 4)
t0 = min(v for v in (prod[0][0], mma[0][0], epi[0][0]) if v)
print("MMA per-stage: [wait_start, wait_done, issued] rel cycles; first 14 and a steady-state slice")
for i in list(range(0, 10)) + list(range(30, 42)):
    print("mma", i, [v - t0 for v in mma[i]], " wait=%d issue=%d" % (mma[i][1] - mma[i][0], mma[i][2] - mma[i][1]))
print("producer: [wait_start, wait_done]")
for i in list(range(0, 8)) + list(range(30, 38)):
    print("prod", i, [v - t0 for v in prod[i]], " wait=%d" % (prod[i][1] - prod[i][0]))
print("epilogue per tile: [wait_start, acc_ready, done]")
for i in list(range(0, 6)) + list(range(20, 24)):
    print("epi", i, [v - t0 for v in epi[i][:3]], " wait=%d work=%d (tmem loads %d, math+stores+arrive %d)" %
          (epi[i][1] - epi[i][0], epi[i][2] - epi[i][1], epi[i][3] - epi[i][1], epi[i][2] - epi[i][3]))
print("total cycles block0:", max(v[2] for v in epi) - t0, " per tile:", (epi[30][2] - epi[10][2]) / 20.0)
