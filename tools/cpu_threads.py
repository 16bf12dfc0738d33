import os, sys, time, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import bin_oracle as O
sd = O.synth_state_dict(0)
fr = O.synth_frames(6, 1, 128, 128)
print("cores", os.cpu_count())
for t in (8, 16, 32, 64):
    torch.set_num_threads(t)
    with torch.no_grad():
        O.window_forward([f[:, :, :32, :32].contiguous() for f in fr], sd)
        t0 = time.perf_counter(); O.window_forward(fr, sd); dt = time.perf_counter() - t0
    print(t, "threads:", round(dt, 2), "s per 128x128 window", flush=True)
