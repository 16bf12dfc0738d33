import os
os.environ.setdefault("BIN_B200_LIB", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "bin_b200", "libbin_b200_tools.so"))  # tools build: timeline hooks + microbenchmarks
"""In-kernel role timeline of the fused RDB tail (BIN_B200_DEBUG=8), 5 x 360 x 640.  usage: timeline_tail.py"""
import ctypes as C, os, sys
os.environ["BIN_B200_DEBUG"] = "8"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from bin_b200 import rdn, _lib
B, h, w = 5, 360, 640
net = rdn.bin_stage4_lstm().cuda().eval()
blob = net.model.model3_1.packed_blob()
x = torch.randn(B, 96, h, w, device="cuda")
y = torch.empty_like(x)
ws = torch.empty(B * 40 * h * w * 16 + 1024, dtype=torch.uint8, device="cuda")
run = lambda: _lib.check(_lib.lib().bin_rdb_fwd(blob.data_ptr(), 5, 7, x.data_ptr(), y.data_ptr(), B, h, w, ws.data_ptr(),
                                               ws.numel(), torch.cuda.current_stream().cuda_stream))
for _ in range(3):
    run()
torch.cuda.synchronize()
buf = (C.c_longlong * (3 * 4096))()
_lib.check(_lib.lib().bin_tools_debug_timeline(buf, 3 * 4096))
a = list(buf)
def role(r, n): return [[a[r * 4096 + i * 4 + j] for j in range(4)] for i in range(n)]
prod, mma, epa = role(0, 400), role(1, 240), role(2, 67)
t0 = min(v for v in (prod[0][0], mma[0][0], epa[0][0]) if v)
print("variant streams =", os.environ.get("BIN_B200_TAIL_STREAMS", "1"))
for i in range(36, 48):
    m = mma[i]
    if m[2] == 0:
        m[2] = m[1]                      # two-stream variant: no turn wait
    print("mma item", i, "(c=%d)" % (2 * (i % 3)), [v - t0 for v in m], "wait_data=%d wait_turn=%d issue=%d gap=%d" % (m[1] - m[0], m[2] - m[1], m[3] - m[2], mma[i + 1][0] - m[3]))
for i in range(120, 132):
    print("prod stage", i, prod[i][0] - t0, "wait_empty=%d" % (prod[i][1] - prod[i][0]))
for i in range(20, 28):
    e = epa[i]
    print("epiA tile", i, [v - t0 for v in e], "wait_conv=%d load+math=%d wait_h+store=%d" % (e[1] - e[0], e[2] - e[1], e[3] - e[2]),
          "| epiB wait_lff=%d at %d" % (prod[i][3] - prod[i][2], prod[i][2] - t0))
print("cycles per tile (epilogue A):", (epa[60][3] - epa[20][3]) / 40.0)
