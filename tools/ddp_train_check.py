"""2-rank DDP sanity check of the training path (the reference's DistributedDataParallel branch, bin_model.py:39-41):
each rank trains on its own crops, gradients are all-reduced by PyTorch DDP over NCCL, parameters stay in sync.
launch: python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/ddp_train_check.py"""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bin_b200 import rdn  # noqa: E402
from bin_b200.loss import pixel_loss  # noqa: E402
from oracle import bin_oracle as O  # noqa: E402  (synthetic weights / inputs only)

rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
net = rdn.bin_stage4_lstm()
net.load_state_dict(O.synth_state_dict(0), strict=True)
net = net.cuda().train()
ddp = torch.nn.parallel.DistributedDataParallel(net, device_ids=[local])
opt = torch.optim.Adam(ddp.parameters(), lr=1e-4, betas=(0.9, 0.99), fused=True)
B, H, W = 2, 128, 128
fr = [f.cuda() for f in O.synth_frames(6, B, H, W, seed=100 + rank, smooth=True)]       # different data per rank
gt = [f.cuda() for f in O.synth_frames(14, B, H, W, seed=200 + rank, smooth=True)]
losses = []
for step in range(3):
    opt.zero_grad(set_to_none=True)
    loss, _ = pixel_loss(ddp(*fr), gt, "cb")
    loss.backward()
    opt.step()
    losses.append(loss.item())
w = net.model.model2_1.RDBs[4].LFF.weight.detach()
g = net.model.model2_1.RDBs[4].LFF.weight.grad.detach()
ws = [torch.zeros_like(w) for _ in range(2)]
gs = [torch.zeros_like(g) for _ in range(2)]
dist.all_gather(ws, w)
dist.all_gather(gs, g)
if rank == 0:
    print(json.dumps({"losses_rank0": losses, "weights_in_sync": bool(torch.equal(ws[0], ws[1])),
                      "grads_in_sync": bool(torch.equal(gs[0], gs[1])), "grad_absmax": g.abs().max().item()}))
dist.destroy_process_group()
