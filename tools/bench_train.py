"""BASELINE config 3: training step (L1 sum loss, bin_model.py:53-61,395-425 semantics without the optimizer),
fwd+bwd of the shipped 6-frame net on batch 8 x 256x256 crops.  Prints one JSON line."""
import json
import os
import sys
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bin_b200 import rdn  # noqa: E402
from oracle import bin_oracle as O  # noqa: E402

B, H, W = (int(v) for v in sys.argv[1:4]) if len(sys.argv) > 3 else (8, 256, 256)
steps, warm = int(os.environ.get("BT_STEPS", 5)), int(os.environ.get("BT_WARM", 2))
net = rdn.bin_stage4_lstm()
net.load_state_dict(O.synth_state_dict(0), strict=True)
net = net.cuda().train()
fr = [f.cuda() for f in O.synth_frames(6, B, H, W, seed=1234, smooth=True)]
gt = [f.cuda() for f in O.synth_frames(14, B, H, W, seed=4321, smooth=True)]
from bin_b200.loss import pixel_loss  # noqa: E402  (fused bin_model.get_loss: 14 GT terms + 3 cycle terms, L1 sum)
from bin_b200.optim import Adam  # noqa: E402        (one-launch multi-tensor Adam)
opt = Adam(net.parameters(), lr=1e-4, betas=(0.9, 0.99))     # yml :52-55


def step():
    opt.zero_grad(set_to_none=True)
    outs = net(*fr)
    loss, _ = pixel_loss(outs, gt, "l1")
    loss.backward()
    return loss


losses = []
for _ in range(warm):
    losses.append(step().item()); opt.step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps):
    l = step()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
# the whole of bin_model.optimize_parameters (:130-141): zero_grad, forward, get_loss, backward, optimizer step
e0.record()
for _ in range(steps):
    l = step()
    opt.step()
e1.record()
torch.cuda.synchronize()
ms_full = e0.elapsed_time(e1) / steps
flops = 3 * 2.0 * 14_234_976 * B * H * W
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops_sustained"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 1400.0
print(json.dumps({"config": f"train step fwd+bwd, 6-frame net, batch {B} x {H}x{W}, get_loss(l1, 17 terms) fused, Adam excluded from timing",
                  "ms_per_step": round(ms, 2), "ms_per_step_with_adam": round(ms_full, 2), "tflops_reference_as_executed(3xF_fwd)": round(flops / ms / 1e9, 1),
                  "frac_of_sustained_peak": round(flops / (ms * 1e-3) / 1e12 / peak, 4),
                  "loss_first_warmup_steps": losses, "loss_last": l.item(),
                  "max_mem_GB": round(torch.cuda.max_memory_allocated() / 2**30, 2)}))
