import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import bin_oracle as O
from bin_b200 import rdn
name, n = (sys.argv[1], int(sys.argv[2])) if len(sys.argv) > 2 else ("model2_1", 3)
B, H, W = 2, 44, 68
sd = O.synth_state_dict(0)
bsd = {k: v.clone().requires_grad_(True) for k, v in O.sub_sd(sd, "model." + name).items()}
fr = [f.requires_grad_(True) for f in O.synth_frames(n, B, H, W, seed=41)]
cot = O.synth_frames(1, B, H, W, seed=42)[0] - 0.5
import contextlib
ctx = O.emulate_fp16_storage() if os.environ.get('EMU') else contextlib.nullcontext()
with ctx:
    y = O.backbone(fr, bsd)
names = list(bsd.keys())
grads = torch.autograd.grad((y * cot).sum(), fr + [bsd[k] for k in names])
net = rdn.bin_stage4_lstm(); net.load_state_dict(sd); net = net.cuda()
model = getattr(net.model, name)
frames = [f.detach().cuda().requires_grad_(True) for f in fr]
yy = model(*frames)
(yy * cot.cuda()).sum().backward()
print("fwd err", (yy.detach().cpu() - y.detach()).abs().max().item())
for k in range(n):
    r = grads[k]; g = frames[k].grad.cpu()
    print("frame", k, "rel err %.4f" % ((g - r).abs().max() / r.abs().max()).item(), "max", r.abs().max().item())
got = dict(model.named_parameters())
rows = []
for key, r in zip(names, grads[n:]):
    g = got[key].grad.cpu()
    rows.append(((g - r).abs().max() / r.abs().max()).item())
    tag = "BAD" if rows[-1] > 0.02 else ""
    print("%-34s rel err %.4f  max|ref| %.4g  corr %.5f %s" % (key, rows[-1], r.abs().max().item(),
          torch.corrcoef(torch.stack([g.flatten(), r.flatten()]))[0, 1].item(), tag))
