"""Two-rank correctness check on real GPUs (launched by tests/test_multigpu.py through torch.distributed.run, one
rank per GPU, NCCL): a bundle sharded over the ranks must give the single-GPU result on the union bundle —
sensor image, parameter gradients of a final-ray loss, and the SpotSizeLoss value / gradients whose moments are
all-reduced inside the goal kernels' wrappers.  Test infrastructure; prints "MULTIGPU OK" on rank 0."""
import os
import sys
import types

import numpy as np
import torch
import torch.distributed as tdist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import raytracetorch_b200 as rtt  # noqa: E402
from raytracetorch_b200 import dist as rdist  # noqa: E402
import parity  # noqa: E402
import scenes  # noqa: E402


def bundle(n, dev, seed):
    g = torch.Generator(device=dev).manual_seed(seed)
    th = torch.rand(n, device=dev, generator=g) * (2 * np.pi)
    r = torch.sqrt(torch.rand(n, device=dev, generator=g)) * 5.0
    pos = torch.stack([r * torch.cos(th), r * torch.sin(th), torch.full_like(r, -10.0)], 1).contiguous()
    dirs = torch.zeros_like(pos)
    dirs[:, 2] = 1.0
    return pos, dirs, torch.ones(n, device=dev)


def build(ns, dev):
    els = scenes.c1_singlet(ns, physical=True, grads=True)
    els[1].set_image(256, 256, extent=(-2.0, 2.0, -2.0, 2.0))
    return rtt.scene.SequentialScene(els).to(dev), els


def grads_of(els):
    return [els[0].shape.surfaces[k].c.grad.detach().clone() for k in (0, 1)]


def main():
    rank, world, local = rdist.init_from_env()
    assert world >= 2, "launch with torch.distributed.run --nproc-per-node 2"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    ns = types.SimpleNamespace(elements=rtt.elements, geom=rtt.geom, phys=rtt.phys, rays=rtt.rays)
    n = 400_003                                           # not a multiple of the world size or of a tile
    pos, dirs, inten = bundle(n, dev, 7)                  # the same union bundle on every rank
    lo, hi = rdist.shard_bounds(n, rank, world)

    # ---- (1) image + final-ray loss gradients: sharded + all-reduced vs the union bundle on one GPU ----
    scene, els = build(ns, dev)
    tab = scene.table()
    out = rtt.ops.trace_sequential(tab, pos[lo:hi].contiguous(), dirs[lo:hi].contiguous(), inten[lo:hi].contiguous(),
                                   want_record=False)
    loss = torch.dot(out["intensity"], (out["pos"][:, :2] ** 2).sum(1)) / n
    loss.backward()
    img = out["images"][0].clone()
    red = rdist.FlatReducer()
    red.add(img)
    red.extend([els[0].shape.surfaces[k].c.grad for k in (0, 1)])
    red.reduce()                                          # ONE collective: image || gradients
    g_sharded = grads_of(els)

    scene1, els1 = build(ns, dev)
    out1 = rtt.ops.trace_sequential(scene1.table(), pos, dirs, inten, want_record=False)
    loss1 = torch.dot(out1["intensity"], (out1["pos"][:, :2] ** 2).sum(1)) / n
    loss1.backward()
    g_single = grads_of(els1)
    img1 = out1["images"][0]
    assert float(img1.sum()) > 0.5 * n
    l1 = float((img - img1).abs().sum() / img1.sum())
    assert l1 <= 1e-5, f"rank {rank}: sharded image differs from the union image, rel L1 {l1}"
    for k in (0, 1):
        e = parity.grad_rel(g_sharded[k].cpu().numpy(), g_single[k].cpu().numpy())
        assert e <= 1e-4, f"rank {rank}: d loss / d c{k + 1} sharded {float(g_sharded[k])} vs union {float(g_single[k])}"

    # ---- (2) SpotSizeLoss term with all-reduced moments vs the eager formula on the union records ----
    scene2, els2 = build(ns, dev)
    o2 = rtt.ops.trace_sequential(scene2.table(), pos[lo:hi].contiguous(), dirs[lo:hi].contiguous(),
                                  inten[lo:hi].contiguous(), want_record=True, want_rays=True)
    term = rtt.ops.spot_size(o2["records"][0])            # moments and sums are all-reduced inside
    term.backward()
    rdist.allreduce_scene_results([], [els2[0].shape.surfaces[k].c for k in (0, 1)])
    g2 = grads_of(els2)

    scene3, els3 = build(ns, dev)
    o3 = rtt.ops.trace_sequential(scene3.table(), pos, dirs, inten, want_record=True)
    rec = o3["records"][0]
    w = rec[:, 3]
    act = w > 0
    xy, w = rec[act, :2], w[act]
    W = w.sum()
    cx, cy = (xy[:, 0] * w).sum() / W, (xy[:, 1] * w).sum() / W
    ref = torch.sqrt(((xy[:, 0] - cx) ** 2 + (xy[:, 1] - cy) ** 2) * (w / W)).sum()
    ref.backward()
    g3 = grads_of(els3)
    assert abs(float(term.detach()) - float(ref.detach())) <= 2e-4 * abs(float(ref.detach())), \
        f"rank {rank}: sharded spot size {float(term)} vs union {float(ref)}"
    for k in (0, 1):
        e = parity.grad_rel(g2[k].cpu().numpy(), g3[k].cpu().numpy())
        assert e <= parity.TOL_GRAD, f"rank {rank}: spot-size gradient c{k + 1}: {float(g2[k])} vs {float(g3[k])}"

    # ---- (3) device ray sources: ranks that share a seed draw DIFFERENT rays (per-rank counter offset) ----
    torch.manual_seed(123)
    src = rtt.rays.CollimatedDisk(5.0, 0, device=dev, transform=rtt.geom.RayTransformBundle(
        translation=[0.0, 0.0, -10.0]).to(dev)).sample(1000)
    mine = src.pos[:, :2].contiguous()
    both = [torch.empty_like(mine) for _ in range(world)]
    tdist.all_gather(both, mine)
    assert not torch.equal(both[0], both[1]), "ranks with the same seed generated the same rays"

    tdist.barrier()
    if rank == 0:
        print("MULTIGPU OK", f"image relL1 {l1:.2e}", flush=True)
    tdist.destroy_process_group()


if __name__ == "__main__":
    main()
