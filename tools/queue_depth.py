"""Why is bench.py's free-running timed loop slower than the same steps timed one by one?  Times 20 steps x 5 windows
(1280x720, graph replay) four ways: free-running, synchronize per step, event-limited run-ahead of one step, and
free-running while a 5 Hz nvidia-smi sampler runs (what bench.py does).  usage: python tools/queue_depth.py"""
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from bin_b200 import rdn
from oracle import bin_oracle as O

dev = torch.device("cuda", 0)
net = rdn.bin_stage4_lstm(); net.load_state_dict(O.synth_state_dict(0), strict=True); net = net.to(dev).eval()
wins = [[f.to(dev) for f in O.synth_frames(6, 1, 720, 1280, seed=1234 + i, smooth=True)] for i in range(5)]
STEPS = 20


def run(mode):
    evs = []
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(STEPS):
        for w in wins:
            outs = net(*w)
        if mode == "sync":
            torch.cuda.synchronize()
        elif mode == "ahead1":
            ev = torch.cuda.Event(); ev.record(); evs.append(ev)
            if len(evs) >= 2:
                evs[-2].synchronize()
        elif mode == "discard":
            del outs
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / STEPS


with torch.no_grad():
    for _ in range(5):
        for w in wins:
            net(*w)
    torch.cuda.synchronize()
    t0 = time.time()
    while time.time() - t0 < 3.0:
        for w in wins:
            net(*w)
        torch.cuda.synchronize()
    for rep in range(2):
        for mode in ("free", "sync", "ahead1", "discard"):
            print(f"{mode:8s} {run(mode):8.2f} ms/step", flush=True)
    smi = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader", "-lms", "200"],
                           stdout=subprocess.DEVNULL)
    time.sleep(1.0)
    for mode in ("free", "sync", "ahead1"):
        print(f"smi+{mode:8s} {run(mode):8.2f} ms/step", flush=True)
    smi.terminate()
