"""Small end-to-end pass over every kernel family, sized for compute-sanitizer (memcheck / racecheck):
sequential forward (packed-pair streaming build, tile build, per-ray EXACT build; focused image = shared-memory image
cache with CAS / probing, spread image = cache switched off), compacting adjoint with and without pose gradients,
non-sequential forward + windowed adjoint, per-id reductions, goal reductions, render kernel, in-kernel sources.
Prints SANITIZE_CASE OK at the end.  Usage: compute-sanitizer --tool <memcheck|racecheck> python scripts/sanitize_case.py"""
import os
import sys
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import raytracetorch_b200 as rtt  # noqa: E402
import scenes  # noqa: E402

dev = torch.device("cuda", 0)
ns = types.SimpleNamespace(elements=rtt.elements, geom=rtt.geom, phys=rtt.phys, rays=rtt.rays, scene=rtt.scene,
                           render=rtt.render)
N = int(os.environ.get("SANITIZE_RAYS", "20011"))        # not a multiple of the 512-ray tile: ragged last tile


def bundle(radius, z, n=N, seed=1):
    g = torch.Generator(device=dev).manual_seed(seed)
    th = torch.rand(n, device=dev, generator=g) * 6.2831853
    r = torch.sqrt(torch.rand(n, device=dev, generator=g)) * radius
    pos = torch.stack([r * torch.cos(th), r * torch.sin(th), torch.full_like(r, z)], 1).contiguous()
    dirs = torch.zeros_like(pos)
    dirs[:, 2] = 1.0
    return pos, dirs, torch.ones(n, device=dev)


# ---- sequential forward, every build, C2 with wavelength table and a 3-channel image (spread -> cache goes off) ----
els = scenes.c2_cylindrical(ns, grads=True)
disp = rtt.Dispersion(scenes.C2_WAVELENGTHS, {els[0].ior_glass: [1.5 * s for s in scenes.C2_GLASS_SCALE],
                                              els[1].ior_glass: [1.6 * s for s in scenes.C2_GLASS_SCALE]})
els[3].set_image(256, 256, channels=3)
scene = rtt.scene.SequentialScene(els)
scene.set_dispersion(disp)
scene = scene.to(dev)
tab = scene.table()
pos, dirs, inten = bundle(8.0, -10.0)
wav = torch.tensor(scenes.C2_WAVELENGTHS, device=dev)[torch.arange(N, device=dev) % 3].contiguous()
ref = None
for tune in (16, 18, 3, 5, 9):
    rtt.ops.set_tuning(fwd=tune)
    out = rtt.ops.trace_sequential(tab, pos, dirs, inten, wav, want_record=True)
    if ref is None:
        ref = out["intensity"].clone()
    assert torch.equal(out["intensity"], ref), tune
rtt.ops.set_tuning(fwd=0)
out = rtt.ops.trace_sequential(tab, pos, dirs, inten, wav, want_record=True, mode=rtt.ops.MODE_EXACT)
# adjoint with pose gradients (C2 grads=True has a trainable pose) and compaction (no ray gradients requested)
out = rtt.ops.trace_sequential(tab, pos, dirs, inten, wav, want_record=True)
loss = torch.dot(out["intensity"], (out["pos"][:, :2] ** 2).sum(1)) + out["records"][0][:, :2].pow(2).sum()
loss.backward()

# ---- focused image (C1): the shared-memory image cache with CAS / probing; scalar-gradient adjoint build ----
els1 = scenes.c1_singlet(ns, physical=True, grads=True)
els1[1].set_image(128, 128, extent=(-2.0, 2.0, -2.0, 2.0))
sc1 = rtt.scene.SequentialScene(els1).to(dev)
p1, d1, i1 = bundle(5.0, -10.0)
p1.requires_grad_(True)
for tune in (16, 5):
    rtt.ops.set_tuning(fwd=tune)
    o1 = rtt.ops.trace_sequential(sc1.table(), p1, d1, i1, want_record=True)
rtt.ops.set_tuning(fwd=0)
(o1["intensity"] * (o1["pos"][:, :2] ** 2).sum(1)).sum().backward()          # ray gradients requested: no compaction
# goal reductions + per-id reductions on the records
src_bundle = rtt.rays.CollimatedDisk(5.0, 0, device=dev, transform=rtt.geom.RayTransformBundle(
    translation=[0.0, 0.0, -10.0]).to(dev))
goal = rtt.optim.SpotSizeLoss(els1[1], [src_bundle], N_rays=N)               # in-kernel generated bundle
goal(sc1).backward()
rec = o1["records"][0].detach()
ids = (torch.arange(N, device=dev) % 3).to(torch.int8)
size, wsum = rtt.ops.spot_size_per_id(rec.clone().requires_grad_(True), ids, [2, 0, 1])
size.sum().backward()

# ---- non-sequential forward + adjoint (C5), deep hit sequences through the windowed adjoint ----
els5 = scenes.c5_nonsequential(ns)
sc5 = rtt.scene.Scene()
for e in els5:
    sc5.add_element(e)
sc5 = sc5.to(dev)
els5[4].set_image(64, 64)
p5, d5, i5 = bundle(10.0, -5.0)
p5.requires_grad_(True)
o5 = rtt.ops.trace_nonsequential(sc5.table(), p5, d5, i5, 8, want_record=True, record_depth=2)
(o5["pos"] ** 2).sum().backward()
T = lambda z: rtt.geom.RayTransform(translation=[0.0, 0.0, z])
res = [rtt.elements.SphericalMirror(c1=-1 / 0.8, d=0.3, diameter=0.3, c1_grad=True, transform=T(0.2)),
       rtt.elements.SphericalMirror(c1=1 / 0.8, d=0.3, diameter=0.3, c1_grad=True, transform=T(-0.2))]
tabr = rtt.compile_elements(res).to(dev) if hasattr(rtt.compile_elements(res), "to") else None
scr = rtt.scene.Scene()
for e in res:
    scr.add_element(e)
scr = scr.to(dev)
g = torch.Generator(device=dev).manual_seed(3)
pr = torch.cat([(torch.rand(2000, 2, device=dev, generator=g) - 0.5) * 0.04, torch.zeros(2000, 1, device=dev)], 1)
dr = torch.nn.functional.normalize(torch.cat([(torch.rand(2000, 2, device=dev, generator=g) - 0.5) * 0.06,
                                              torch.ones(2000, 1, device=dev)], 1), dim=1)
orr = rtt.ops.trace_nonsequential(scr.table(), pr, dr, torch.ones(2000, device=dev), 48, want_record=False)
assert int(orr["n_hits"].max()) == 48
(orr["pos"] ** 2).sum().backward()

# ---- render kernel with in-kernel camera rays ----
scene_r, cam = scenes.render_setup(ns, device="cuda")
img = rtt.render.Renderer(scene_r.cuda()).render_3d(cam)
assert img.shape == (cam.height, cam.width, 3)
torch.cuda.synchronize()
print("SANITIZE_CASE OK")
