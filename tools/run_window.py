"""Run N 720p windows (for ncu / timing). usage: run_window.py [n_windows] [H W]"""
import os
import sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bin_b200 import rdn  # noqa: E402
from oracle import bin_oracle as O  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
H, W = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (720, 1280)
net = rdn.bin_stage4_lstm()
net.load_state_dict(O.synth_state_dict(0), strict=True)
net = net.cuda().eval()
if "--fp32" in sys.argv:
    rdn.set_precision(net, "fp32")
BATCH = int(os.environ.get("RW_BATCH", "1"))          # windows per forward (batched along N)
fr = [f.cuda() for f in O.synth_frames(6, BATCH, H, W, seed=1234, smooth=True)]
with torch.no_grad():
    for i in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        outs = net(*fr)
        e1.record()
        torch.cuda.synchronize()
        print(f"window {i}: {e0.elapsed_time(e1):.3f} ms", flush=True)
if "--graph" in sys.argv:
    g = torch.cuda.CUDAGraph()
    with torch.no_grad():
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            net(*fr)
        torch.cuda.current_stream().wait_stream(s)
        with torch.cuda.graph(g):
            gouts = net(*fr)
    for i in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        print(f"graph window {i}: {e0.elapsed_time(e1):.3f} ms" + (f" ({e0.elapsed_time(e1) / BATCH:.3f} ms per window, {BATCH} per forward)" if BATCH > 1 else ""), flush=True)
    print("graph outputs equal eager:", all(torch.equal(a, b) for a, b in zip(gouts, outs)))
