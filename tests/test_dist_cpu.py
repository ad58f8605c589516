"""World-size-2 gloo tests (CPU) of the multi-GPU host logic: unit sharding covers every unit
exactly once, the flat-arena gradient all-reduce equals the single-process mean, and the
per-sample gather restores global order."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from isegprobe_b200 import dist as idist


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
    try:
        # 1. sharding
        mine = idist.shard_indices(11)
        # 2. gradient arena: rank-dependent grads, mean must equal the analytic mean
        torch.manual_seed(0)
        lin = torch.nn.Linear(5, 3)
        arena = idist.FlatGradArena(list(lin.parameters()))
        x = torch.full((4, 5), float(rank + 1))
        lin(x).sum().backward()
        local = arena.flat.clone()
        arena.all_reduce_mean()
        # 3. gather of per-sample rows
        rows = torch.tensor([[float(i), float(i) * 10] for i in mine])
        full = idist.gather_sample_results(rows, 11)
        q.put((rank, mine, local.tolist(), arena.flat.tolist(), full.tolist(),
               lin.weight.grad.data_ptr() == arena.flat.data_ptr()))  # plain lists: no shared-memory handles
    finally:
        dist.destroy_process_group()


def test_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    (r0, m0, l0, a0, f0, v0), (r1, m1, l1, a1, f1, v1) = res
    l0, a0, f0, l1, a1, f1 = (torch.tensor(t) for t in (l0, a0, f0, l1, a1, f1))
    assert sorted(m0 + m1) == list(range(11)) and not set(m0) & set(m1)
    assert torch.allclose(a0, (l0 + l1) / 2) and torch.equal(a0, a1)
    want = torch.tensor([[float(i), float(i) * 10] for i in range(11)])
    assert torch.equal(f0, want) and torch.equal(f1, want)
    assert v0 and v1  # parameter .grad is a view into the arena


def test_single_process_defaults():
    assert idist.world() == 1 and idist.rank() == 0
    assert idist.shard_indices(5) == [0, 1, 2, 3, 4]
    t = torch.arange(6).reshape(6, 1)
    assert torch.equal(idist.shard_batch(t, 2, 1), torch.tensor([[1], [3], [5]]))
    assert torch.equal(idist.gather_sample_results(t.float(), 6), t.float())
