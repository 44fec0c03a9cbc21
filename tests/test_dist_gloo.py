"""Multi-rank host logic on CPU: world_size-2 gloo process group (the kernels themselves need a
GPU; what is covered here is everything around them that differs at N > 1)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from raytracetorch_b200 import dist as rdist, optim
    r, w, _ = rdist.init_from_env(backend="gloo")
    assert (r, w) == (rank, world)
    res = {}
    # shards tile the bundle, also when it is smaller than the world
    for n in (0, 1, 7, 1000, 1001):
        lo, hi = rdist.shard_bounds(n)
        t = torch.zeros(max(n, 1))
        t[lo:hi] += 1
        dist.all_reduce(t)
        res[f"tile{n}"] = bool((t[:n] == 1).all()) if n else (lo, hi) == (0, 0)
    # one flat all-reduce of [image || grads || moments]
    img = torch.full((3, 4, 5), float(rank + 1))
    grad = torch.arange(6.0).view(2, 3) * (rank + 1)
    mom = torch.tensor([1.0, 2.0]) * (rank + 1)
    red = rdist.FlatReducer()
    red.add(img), red.add(grad), red.add(None), red.add(mom)
    red.reduce()
    tot = sum(range(1, world + 1))
    res["flat"] = bool(torch.equal(img, torch.full((3, 4, 5), float(tot))) and
                       torch.equal(grad, torch.arange(6.0).view(2, 3) * tot) and torch.equal(mom, torch.tensor([1.0, 2.0]) * tot))
    # single contiguous tensor: reduced in place without the pack/unpack copy
    one = torch.ones(8) * (rank + 1)
    r1 = rdist.FlatReducer()
    r1.add(one)
    r1.reduce()
    res["single"] = bool(torch.equal(one, torch.ones(8) * tot))
    # autograd-aware sum used by the goals: sharded weighted centroid == unsharded, gradients too
    torch.manual_seed(0)
    x = torch.randn(10, requires_grad=False)
    wgt = torch.rand(10)
    p = torch.tensor(2.0, requires_grad=True)
    lo, hi = rdist.shard_bounds(10)
    xs, ws = (x * p)[lo:hi], wgt[lo:hi]
    mom = optim._dist_sum(torch.stack([ws.sum(), (xs * ws).sum()]))
    c = mom[1] / mom[0]
    c.backward()
    rdist.allreduce_scene_results([], [p])            # the one gradient collective of a step
    pr = torch.tensor(2.0, requires_grad=True)
    cr = ((x * pr) * wgt).sum() / wgt.sum()
    cr.backward()
    res["moments"] = bool(torch.allclose(c.detach(), cr.detach(), atol=1e-6) and torch.allclose(p.grad, pr.grad, atol=1e-6))
    out[rank] = res
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_world_size_two_gloo():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    assert len(out) == world
    for rank in range(world):
        for k, v in out[rank].items():
            assert v, f"rank {rank}: {k}"


def test_shard_bounds_are_balanced():
    from raytracetorch_b200.dist import shard_bounds
    for n in (0, 1, 5, 8, 10 ** 9 + 7):
        for w in (1, 2, 3, 8):
            spans = [shard_bounds(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
