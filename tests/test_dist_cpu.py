"""world_size-2 gloo test of the sharding / weight-broadcast plumbing."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from bin_b200 import dist as bd
    from bin_b200 import rdn
    torch.manual_seed(100 + rank)                    # ranks start with DIFFERENT weights
    net = rdn.bin_stage4_lstm()
    before = net.model.model3_1.GFF[0].weight.clone()
    nbytes = bd.broadcast_weights(net, src=0)
    after = net.model.model3_1.GFF[0].weight
    ref = [torch.zeros_like(after) for _ in range(world)]
    dist.all_gather(ref, after.detach())
    same = all(torch.equal(ref[0], r) for r in ref)
    mine = bd.shard_windows(11, rank, world)
    t = bd.max_over_ranks(float(rank + 1), "cpu")
    q.put((rank, nbytes, same, bool(torch.equal(before, after)), mine, t))
    dist.destroy_process_group()


def test_broadcast_and_sharding_world2():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    [p.start() for p in procs]
    res = sorted(q.get(timeout=180) for _ in range(world))
    [p.join(60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    (r0, n0, same0, unchanged0, w0, t0), (r1, n1, same1, unchanged1, w1, t1) = res
    assert n0 == n1 == 11_441_668 * 4
    assert same0 and same1
    assert unchanged0 and not unchanged1             # rank 0 keeps its weights, rank 1 receives them
    assert sorted(w0 + w1) == list(range(11)) and not set(w0) & set(w1)
    assert t0 == t1 == 2.0
