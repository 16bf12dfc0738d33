"""BASELINE config 2a: stages 1-3 on 4 frames at 1280x720 (6 outputs)."""
import json, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bin_b200 import rdn
from oracle import bin_oracle as O
net = rdn.bin_stage4_lstm(); net.load_state_dict(O.synth_state_dict(0)); net = net.cuda().eval()
fr = [f.cuda() for f in O.synth_frames(4, 1, 720, 1280, seed=1234, smooth=True)]
with torch.no_grad():
    for _ in range(3): net.forward_pyramid3(*fr)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): net.forward_pyramid3(*fr)
    e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(json.dumps({"config": "2a: 4-frame 3-stage pyramid 1280x720 (3xS1+2xS2+1xS3), eager launches", "ms": round(ms, 2),
                  "windows_per_s": round(1e3 / ms, 2), "tflops": round(2 * 4252320 * 720 * 1280 / ms / 1e9, 1)}))
