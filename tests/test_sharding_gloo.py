"""world_size-2 gloo test (CPU) of the multi-GPU host logic: episode sharding, best-cut gather, gradient mean."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, n_total, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from eco_dqn_b200 import sharding
    lo, hi = sharding.shard_range(n_total, rank, world)
    local = torch.arange(lo, hi, dtype=torch.int32) * 3 + 1          # stand-in for this rank's best cuts
    allc = sharding.gather_best(local, n_total)
    spins = torch.stack([torch.full((5,), int(i), dtype=torch.int8) for i in range(lo, hi)]) if hi > lo else \
        torch.zeros((0, 5), dtype=torch.int8)
    alls = sharding.gather_best(spins, n_total)
    grad = torch.full((7,), float(rank + 1))
    sharding.allreduce_mean_(grad)
    # gradient averaging over a module: one flat collective, scattered back in place
    torch.manual_seed(0)
    lin = torch.nn.Sequential(torch.nn.Linear(3, 4), torch.nn.Linear(4, 2, bias=False))
    lin(torch.full((5, 3), float(rank + 1))).sum().backward()
    local = [p.grad.clone() for p in lin.parameters()]
    flat = sharding.allreduce_mean_grads(lin.parameters())
    others = [torch.zeros_like(flat) for _ in range(world)]
    dist.all_gather(others, torch.cat([g.reshape(-1) for g in local]))
    assert torch.allclose(flat, sum(others) / world)
    assert torch.allclose(torch.cat([p.grad.reshape(-1) for p in lin.parameters()]), flat)
    # gradients that are consecutive views of one buffer (what eco_mpnn_grad leaves): reduced in place, no copies
    buf = torch.arange(20, dtype=torch.float32) * (rank + 1)
    lin[0].weight.grad = buf[0:12].view(4, 3)
    lin[0].bias.grad = buf[12:16]
    flat2 = sharding.allreduce_mean_grads(list(lin.parameters())[:2])
    assert flat2.data_ptr() == buf.data_ptr() and flat2.numel() == 16
    assert torch.allclose(buf[:16], torch.arange(16, dtype=torch.float32) * 1.5)
    assert torch.allclose(buf[16:], torch.arange(16, 20, dtype=torch.float32) * (rank + 1))
    assert torch.allclose(lin[0].bias.grad, torch.arange(12, 16, dtype=torch.float32) * 1.5)
    per_graph = sharding.best_per_graph(allc, torch.arange(n_total) % 3, 3)
    q.put((rank, lo, hi, allc.tolist(), alls[:, 0].tolist(), grad.tolist(), per_graph.tolist()))
    dist.destroy_process_group()


def test_shard_gather_allreduce_world2():
    from eco_dqn_b200 import sharding
    for n_total in (9, 8):
        ctx = mp.get_context("spawn")
        q = ctx.Queue()
        port = _free_port()
        procs = [ctx.Process(target=_worker, args=(r, 2, port, n_total, q)) for r in range(2)]
        for p in procs:
            p.start()
        out = sorted(q.get(timeout=120) for _ in range(2))
        for p in procs:
            p.join(timeout=60)
            assert p.exitcode == 0
        want = [3 * i + 1 for i in range(n_total)]
        covered = []
        for rank, lo, hi, allc, alls, grad, per_graph in out:
            assert (lo, hi) == sharding.shard_range(n_total, rank, 2)
            covered += list(range(lo, hi))
            assert allc == want and alls == list(range(n_total))
            assert grad == [1.5] * 7
            assert per_graph == [max(w for i, w in enumerate(want) if i % 3 == g) for g in range(3)]
        assert covered == list(range(n_total))


def test_shard_range_properties():
    from eco_dqn_b200.sharding import shard_range
    for n in (0, 1, 7, 4096, 32768):
        for w in (1, 2, 4, 8):
            blocks = [shard_range(n, r, w) for r in range(w)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(blocks, blocks[1:]))
            sizes = [b - a for a, b in blocks]
            assert max(sizes) - min(sizes) <= 1
