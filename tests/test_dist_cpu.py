"""N > 1 bookkeeping with gloo on CPU, world size 2: static stream sharding and whole-job
throughput aggregation (sum of frames / max device time over ranks), as bench.py uses them."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    import sys
    from conftest import PKG
    sys.path.insert(0, PKG)
    from consumers.streams import aggregate_throughput, shard_streams

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = shard_streams(7, world, rank)
    frames = 10.0 * len(mine)
    elapsed = 100.0 * (rank + 1)  # rank 1 is slower
    total, tmax, fps = aggregate_throughput(frames, elapsed)
    dist.barrier()
    q.put((rank, mine, total, tmax, fps))
    dist.destroy_process_group()


def test_stream_sharding_and_aggregation_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, s0, t0, m0, f0), (r1, s1, t1, m1, f1) = res
    assert s0 == [0, 2, 4, 6] and s1 == [1, 3, 5]          # disjoint, complete
    assert t0 == t1 == 70.0 and m0 == m1 == 200.0           # sum of frames, MAX of times
    assert abs(f0 - 350.0) < 1e-9 and f0 == f1


def test_single_process_passthrough():
    import sys
    from conftest import PKG
    sys.path.insert(0, PKG)
    from consumers.streams import aggregate_throughput, shard_streams

    assert shard_streams(4, 1, 0) == [0, 1, 2, 3]
    assert aggregate_throughput(30.0, 10.0) == (30.0, 10.0, 3000.0)


def _policy_worker(rank, world, port, q):
    import sys
    from conftest import PKG
    sys.path.insert(0, PKG)
    import torch.nn as nn
    from blockcopy.policy.policy import PolicyTrainRL

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(100 + rank)                       # ranks start from DIFFERENT weights and see different data
    net = nn.Sequential(nn.Conv2d(4, 8, 3, padding=1), nn.ReLU(), nn.Conv2d(8, 1, 3, padding=1))
    opt = torch.optim.RMSprop(net.parameters(), lr=1e-2)
    pol = PolicyTrainRL(block_size=8, block_target=0.3, optimizer=opt, complexity_weight=1.0, policy_net=net,
                        information_gain=None, shared_across_ranks=True)
    pol.sync_shared_policy()
    start = torch.cat([p.detach().reshape(-1) for p in net.parameters()]).clone()
    for step in range(3):
        x = torch.randn(1, 4, 6, 6)
        net(x).square().mean().backward()
        pol._allreduce_gradients()
        opt.step()
        opt.zero_grad(set_to_none=True)
    end = torch.cat([p.detach().reshape(-1) for p in net.parameters()]).clone()
    q.put((rank, start.tolist(), end.tolist()))  # plain lists: tensors in a Queue die with the worker
    dist.barrier()
    dist.destroy_process_group()


def test_shared_policy_gradient_allreduce_world2():
    """block_policy_shared: weights broadcast from rank 0, gradients averaged over ranks every training step
    -> the replicas stay bit-identical although every rank sees different frames."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_policy_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=180) for _ in procs), key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (_, s0, e0), (_, s1, e1) = res
    assert s0 == s1, "initial weights must come from rank 0"
    assert e0 == e1, "replicas diverged"
    assert s0 != e0, "no training happened"
