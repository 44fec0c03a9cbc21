"""Generate ``tests/golden/*.npz`` by running the UNMODIFIED reference in this container.

Test infrastructure only.  Usage (from the repo root, build container only — needs
``/root/reference``):

    python -m oracle.make_golden            # all cases
    python -m oracle.make_golden c1_singlet # some cases

For every case in ``tests/scenes.py::CASES`` the script
  1. builds the scene from the reference's own classes and samples a seeded bundle with
     the reference's own ``Bundle.sample``;
  2. runs the reference trace (``SequentialScene.simulate`` — scene/sequential.py:12 — or
     the ``Scene.step`` bounce loop — scene/base.py:129-235) in fp32 and again in fp64;
  3. for the gradient cases, back-propagates a scalar loss through the reference with
     torch autograd and stores d loss / d parameter and d loss / d input rays;
  4. compiles the *reference objects* with this repo's scene compiler and stores the
     surface table, so tests can check that this repo's mirror classes compile to the very
     same table without the reference being present.
The fixtures are small (N_RAYS rays per case) and committed.
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle.ref_loader import load_reference          # noqa: E402
from raytracetorch_b200.table import compile_elements  # noqa: E402
import scenes                                          # noqa: E402

N_RAYS = 3000
NBOUNCES = 8
OUT = os.path.join(ROOT, "tests", "golden")
SEED = 1234


def ref_namespace():
    R = load_reference()
    return types.SimpleNamespace(elements=R.elements, geom=R.geom, phys=R.phys, rays=R.rays, scene=R.scene,
                                 optim=R.optim, render=R.render)


def _run_seq(ns, elements, rays):
    scene = ns.scene.SequentialScene(elements)
    out = scene.simulate(rays)
    return out


def _run_nonseq(ns, elements, rays, nb=NBOUNCES):
    scene = ns.scene.Scene()
    for e in elements:
        scene.add_element(e)
    scene.rays = rays
    scene.Nbounces = nb
    scene._build_index_maps()
    # bounce loop of Scene.simulate (scene/base.py:139-142) with a per-bounce winner log
    seq = torch.full((rays.pos.shape[0], nb), -1, dtype=torch.long)
    for b in range(nb):
        if not (scene.rays.intensity > 0).any():
            break
        res = scene.ray_cast(scene.rays)
        if res is not None:
            hit_mask, we, ws = res
            active = hit_mask & (scene.rays.intensity > 0)
            # flat row index = position in (map_to_element, map_to_surface)
            starts = torch.tensor([0] + [len(e.shape) for e in elements]).cumsum(0)[:-1]
            seq[active, b] = (starts[we] + ws)[active]
        scene.step()
    return scene.rays, seq


def _sensor_dump(elements, dtype):
    out = {}
    slot = 0
    for e in elements:
        if type(e).__name__ == "Sensor":
            if e.hitLocs:
                locs, w, _ = e.getHitsTensors()
            else:
                locs, w = torch.zeros(0, 3, dtype=dtype), torch.zeros(0, dtype=dtype)
            out[f"sensor{slot}_loc"] = locs.detach().numpy()
            out[f"sensor{slot}_w"] = w.detach().numpy()
            slot += 1
    return out


def _cast_rays(ns, rays, dtype):
    c = lambda t: t.detach().clone().to(dtype)
    return ns.rays.Rays(pos=c(rays.pos), dir=c(rays.dir), intensity=c(rays.intensity),
                        id=rays.id, wavelength=c(rays.wavelength), batch_size=rays.batch_size)


def gen_forward_case(ns, name):
    builder, kw, mode, bspec = scenes.CASES[name]
    data = {}
    rays0 = scenes.make_bundle(ns, bspec, N_RAYS, SEED)
    data["in_pos"], data["in_dir"] = rays0.pos.numpy().copy(), rays0.dir.numpy().copy()
    data["in_intensity"] = rays0.intensity.numpy().copy()
    for tag, dtype in (("f32", torch.float32), ("f64", torch.float64)):
        elements = builder(ns, **kw)
        if dtype == torch.float64:
            for e in elements:
                e.double()
        rays = _cast_rays(ns, rays0, dtype)
        with torch.no_grad():
            if mode == "seq":
                out = _run_seq(ns, elements, rays)
            else:
                out, seq = _run_nonseq(ns, elements, rays)
                data[f"{tag}_seq"] = seq.numpy()
        data[f"{tag}_pos"], data[f"{tag}_dir"] = out.pos.numpy(), out.dir.numpy()
        data[f"{tag}_intensity"] = out.intensity.numpy()
        for k, v in _sensor_dump(elements, dtype).items():
            data[f"{tag}_{k}"] = v
        if dtype == torch.float32:
            tab = compile_elements(builder(ns, **kw))
            data["table_f"], data["table_i"] = tab.f.detach().numpy(), tab.i.numpy()
    data["mode"] = np.array(mode)
    data["nbounces"] = np.array(NBOUNCES)
    return data


GRAD_CASES = scenes.GRAD_CASES


def golden_loss(pos, dir_, intensity):
    """Scalar used by every gradient fixture: touches positions, directions and intensity."""
    return (intensity * (pos[:, 0] ** 2 + pos[:, 1] ** 2)).mean() \
        + 0.25 * (intensity * dir_[:, 2]).mean() + 0.1 * (dir_[:, 0] * pos[:, 1]).mean()


def gen_grad_case(ns, name):
    builder, kw, bspec = GRAD_CASES[name]
    data = {}
    rays0 = scenes.make_bundle(ns, bspec, N_RAYS, SEED)
    data["in_pos"], data["in_dir"] = rays0.pos.numpy().copy(), rays0.dir.numpy().copy()
    data["in_intensity"] = rays0.intensity.numpy().copy()
    for tag, dtype in (("f32", torch.float32), ("f64", torch.float64)):
        elements = builder(ns, **kw)
        if dtype == torch.float64:
            for e in elements:
                e.double()
        rays = _cast_rays(ns, rays0, dtype)
        rays.pos.requires_grad_(True)
        rays.dir.requires_grad_(True)
        rays.intensity.requires_grad_(True)
        leaf_pos, leaf_dir, leaf_int = rays.pos, rays.dir, rays.intensity
        out = _run_seq(ns, elements, rays)
        loss = golden_loss(out.pos, out.dir, out.intensity)
        loss.backward()
        data[f"{tag}_loss"] = np.array(loss.item())
        data[f"{tag}_g_pos"] = leaf_pos.grad.numpy()
        data[f"{tag}_g_dir"] = leaf_dir.grad.numpy()
        data[f"{tag}_g_intensity"] = leaf_int.grad.numpy()
        scene = ns.scene.SequentialScene(elements)
        for pname, p in scene.named_parameters():
            if p.requires_grad:
                g = p.grad if p.grad is not None else torch.zeros_like(p)
                data[f"{tag}_gp::{pname}"] = g.detach().numpy()
    return data


def gen_camera_case(ns):
    """Camera.generate_rays (render/camera.py:39-72), 64x36 pixels, BASELINE config-4 geometry."""
    cam = ns.render.Camera((0.0, 0.0, -200.0), (0.0, 0.0, 0.0), (0.0, 1.0, 0.0), 6.0, 64, 36)
    r = cam.generate_rays()
    return dict(pos=r.pos.numpy(), dir=r.dir.numpy(), intensity=r.intensity.numpy())


def gen_render_case(ns):
    """Renderer.render_3d (render/camera.py:191-257) of the reference, fp32, 96x64 pixels."""
    scene, cam = scenes.render_setup(ns)
    img = ns.render.Renderer(scene).render_3d(cam)
    return dict(image=img.numpy())


def gen_fresnel_case(ns):
    """Statistics of the reference's RefractFresnel singlet (phys/std.py:146-224): fraction of rays that leave the
    sequential trace travelling backwards.  Its draws are torch.rand_like, so only the statistic is comparable."""
    n = 200_000
    torch.manual_seed(99)
    els = scenes.c1_singlet(ns, physical=True, inked=False, fresnel=True)
    rays = scenes.make_bundle(ns, ("coll", 11.0, -10.0, [0.3, 0.1, 0.0]), n, 8)
    with torch.no_grad():
        out = _run_seq(ns, els, rays)
    return dict(n_rays=np.array(n), back_fraction=np.array(float((out.dir[:, 2] < 0).float().mean())))


GOAL_SEED, GOAL_RAYS, GOAL_BOUNCES = 77, 2000, 6


def goal_setup(ns):
    """Scene + bundles of the goal fixtures (shared with tests/test_goals.py)."""
    elements = scenes.c1_singlet(ns, physical=True, grads=True)
    scene = ns.scene.Scene()
    for e in elements:
        scene.add_element(e)
    scene.Nbounces = GOAL_BOUNCES
    mk = lambda rid, rot: ns.rays.CollimatedDisk(5.0, rid, transform=ns.geom.RayTransformBundle(
        translation=[0.0, 0.0, -10.0], rotation=rot))
    bundles = [mk(0, None), mk(1, [0.02, 0.0, 0.0]), mk(2, [0.0, -0.03, 0.0])]
    return scene, elements, bundles


class ieee_sqrt:
    """Run the (unmodified) reference with ``torch.sqrt`` rounded correctly.

    torch's CPU sqrt is MKL VML's (< 1 ulp, not correctly rounded, so its last bit is library
    dependent); every IEEE device, including sqrt.rn.f32 on the GPU, rounds correctly.  On the base
    ``Scene`` that last bit decides which rays re-hit the surface they are leaving, which
    moves a SpotSizeLoss by ~0.3 %: bit-level comparisons of non-sequential results therefore use the
    reference's arithmetic with this one function swapped (fp32 sqrt evaluated in double and rounded
    once, which is the correctly rounded fp32 result)."""

    def __enter__(self):
        self._orig = torch.sqrt
        orig = self._orig
        torch.sqrt = lambda x, *a, **k: orig(x.double(), *a, **k).float() if x.dtype == torch.float32 else orig(x, *a, **k)
        return self

    def __exit__(self, *exc):
        torch.sqrt = self._orig


def gen_goal_case(ns):
    """SpotSizeLoss / SpotTargetLoss (optim/goals.py) on the base Scene, loss and d loss / d c1, c2:
    the stock run (keys ``<goal>_*``) and the run with a correctly rounded sqrt (``ieee_<goal>_*``)."""
    data = _gen_goal_case(ns, "")
    with ieee_sqrt():
        data.update({k: v for k, v in _gen_goal_case(ns, "ieee_").items() if k.startswith("ieee_")})
    return data


def _gen_goal_case(ns, prefix):
    data = {}
    _scene, _els, bundles = goal_setup(ns)
    torch.manual_seed(GOAL_SEED)
    for k, b in enumerate(bundles):                   # the reference's own samples under this seed
        r = b.sample(GOAL_RAYS)
        data[f"bundle{k}_pos"], data[f"bundle{k}_dir"] = r.pos.numpy(), r.dir.numpy()
    for name in ("spot_size", "spot_size_target", "spot_target"):
        scene, elements, bundles = goal_setup(ns)
        sensor = elements[1]
        torch.manual_seed(GOAL_SEED)
        if name == "spot_size":
            loss = ns.optim.SpotSizeLoss(sensor, bundles, N_rays=GOAL_RAYS)(scene)
        elif name == "spot_size_target":
            loss = ns.optim.SpotSizeLoss(sensor, bundles, N_rays=GOAL_RAYS, target_xy=[0.1, -0.2])(scene)
        else:
            loss = ns.optim.SpotTargetLoss(sensor, torch.tensor([[0.0, 0.0], [0.0, 2.0], [3.0, 0.0]]))(
                scene, bundles, N_rays=GOAL_RAYS)
        loss.backward()
        data[f"{prefix}{name}_loss"] = np.array(loss.item())
        for k in (0, 1):
            data[f"{prefix}{name}_g_c{k}"] = elements[0].shape.surfaces[k].c.grad.numpy().copy()
    return data


def gen_spot_id_case(ns):
    """Sensor.getSpotSizeParallel_xy (elements/sensor.py:87-176) of the UNMODIFIED reference on a multi-bundle
    sequential trace: three bundles (ray ids 0, 1, 2; the goal fixture's samples) through the C1 singlet onto its
    sensor, then per-id spot sizes for query ids [2, 0, 1] — centroid and fixed targets, norm orders 2 and 3 — and the
    gradients of their sums w.r.t. the two curvatures.  (getSpotSizeID_xy raises in the reference at this snapshot —
    a 0-dim centroid is indexed with [None, :] — so only the parallel method pins this path.)"""
    data = {}
    query = [2, 0, 1]
    targets = torch.tensor([[0.05, -0.02], [0.0, 0.0], [-0.03, 0.04]])
    data["query"], data["targets"] = np.array(query), targets.numpy()
    for tag, tgt, p in (("centroid_p2", None, 2), ("target_p2", targets, 2), ("centroid_p3", None, 3)):
        _scene, elements, bundles = goal_setup(ns)
        sensor = elements[1]
        seq = ns.scene.SequentialScene(elements)
        torch.manual_seed(GOAL_SEED)
        for b in bundles:
            seq.simulate(b.sample(GOAL_RAYS))
        res, wsum = sensor.getSpotSizeParallel_xy(query, target_xy=tgt, norm_ord=p)
        res.sum().backward()
        data[f"{tag}_result"], data[f"{tag}_intensity_sum"] = res.detach().numpy().copy(), wsum.detach().numpy().copy()
        for k in (0, 1):
            data[f"{tag}_g_c{k}"] = elements[0].shape.surfaces[k].c.grad.numpy().copy()
    return data


def main(argv):
    os.makedirs(OUT, exist_ok=True)
    ns = ref_namespace()
    wanted = set(argv)
    for name in scenes.CASES:
        if wanted and name not in wanted:
            continue
        d = gen_forward_case(ns, name)
        np.savez_compressed(os.path.join(OUT, f"{name}.npz"), **d)
        alive = (d["f32_intensity"] > 0).mean()
        print(f"{name:28s} rows={d['table_f'].shape[0]:2d} alive={alive:.3f} "
              f"sensor_hits={d.get('f32_sensor0_w', np.zeros(0)).shape[0]}")
    if not wanted or "extras" in wanted:
        np.savez_compressed(os.path.join(OUT, "extra_camera_rays.npz"), **gen_camera_case(ns))
        fr = gen_fresnel_case(ns)
        np.savez_compressed(os.path.join(OUT, "extra_fresnel.npz"), **fr)
        print("fresnel: back fraction", float(fr["back_fraction"]))
        r3 = gen_render_case(ns)
        np.savez_compressed(os.path.join(OUT, "extra_render3d.npz"), **r3)
        print("render_3d: non-background pixels", int((np.abs(r3["image"] - 1.0).sum(-1) > 0).sum()))
        g = gen_goal_case(ns)
        np.savez_compressed(os.path.join(OUT, "extra_goals.npz"), **g)
        print("extras:", {k: float(v) for k, v in g.items() if k.endswith("_loss")})
    if not wanted or "extras" in wanted or "spot_id" in wanted:
        sid = gen_spot_id_case(ns)
        np.savez_compressed(os.path.join(OUT, "extra_spot_id.npz"), **sid)
        print("spot_id:", {k: v.tolist() for k, v in sid.items() if k.endswith("_result")})
    for name in GRAD_CASES:
        if wanted and name not in wanted:
            continue
        d = gen_grad_case(ns, name)
        np.savez_compressed(os.path.join(OUT, f"{name}.npz"), **d)
        keys = [k for k in d if k.startswith("f32_gp::")]
        print(f"{name:28s} loss={float(d['f32_loss']):.6f} params={len(keys)}")


if __name__ == "__main__":
    torch.set_num_threads(8)
    main(sys.argv[1:])
