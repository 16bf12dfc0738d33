import os
os.environ.setdefault("BIN_B200_LIB", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "bin_b200", "libbin_b200_tools.so"))  # tools build: options are re-read per call
"""A/B inside one process (eager launches): unfused RDB tail vs the fused kernel (two tile streams / hand-off between the MMA warps).
Alternates the configurations so that clock / power drift hits all of them equally.  usage: ab_tail.py [rounds]"""
import os
import statistics
import sys

os.environ["BIN_B200_GRAPH"] = "0"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch                       # noqa: E402
from bin_b200 import rdn           # noqa: E402
from oracle import bin_oracle as O  # noqa: E402

rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 4
net = rdn.bin_stage4_lstm()
net.load_state_dict(O.synth_state_dict(0), strict=True)
net = net.cuda().eval()
fr = [f.cuda() for f in O.synth_frames(6, 1, 720, 1280, seed=1234, smooth=True)]
CFG = {"unfused": {"BIN_B200_FUSE_LFF": "0"}, "fused/streams": {"BIN_B200_FUSE_LFF": "1", "BIN_B200_TAIL_STREAMS": "1"},
       "fused/handoff": {"BIN_B200_FUSE_LFF": "1", "BIN_B200_TAIL_STREAMS": "0"}}
times = {k: [] for k in CFG}
ref = None
with torch.no_grad():
    for r in range(rounds + 1):
        for name, env in CFG.items():
            os.environ.update(env)
            for i in range(3):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                outs = net(*fr)
                e1.record()
                torch.cuda.synchronize()
                if r > 0:
                    times[name].append(e0.elapsed_time(e1))
            if name == "unfused":
                ref = [o.clone() for o in outs]
            else:
                print(name, "max |diff| vs unfused over the 14 outputs: %.3e" % max(float((a - b).abs().max()) for a, b in zip(outs, ref)))
for k, v in times.items():
    print(f"{k:14s} median {statistics.median(v):.3f} ms  min {min(v):.3f}  max {max(v):.3f}  (n={len(v)})")
