"""Ray sources, camera and optimisation goals against fixtures produced by the reference
(oracle/make_golden.py: gen_camera_case, gen_goal_case)."""
import types

import numpy as np
import pytest
import torch

import parity


def _goal_setup(rtt_ns):
    import raytracetorch_b200 as rtt
    import scenes
    from oracle.make_golden import GOAL_BOUNCES
    ns = types.SimpleNamespace(elements=rtt.elements, geom=rtt.geom, phys=rtt.phys, rays=rtt.rays, scene=rtt.scene)
    elements = scenes.c1_singlet(ns, physical=True, grads=True)
    scene = rtt.scene.Scene()
    for e in elements:
        scene.add_element(e)
    scene.Nbounces = GOAL_BOUNCES
    mk = lambda rid, rot, dev: rtt.rays.CollimatedDisk(5.0, rid, device=dev, transform=rtt.geom.RayTransformBundle(
        translation=[0.0, 0.0, -10.0], rotation=rot))
    return scene, elements, mk


def test_camera_generate_rays_matches_reference():
    """render/camera.py:39-72 — same pixel order, origins and (normalised) directions."""
    import raytracetorch_b200 as rtt
    d = parity.load("extra_camera_rays")
    cam = rtt.render.Camera((0.0, 0.0, -200.0), (0.0, 0.0, 0.0), (0.0, 1.0, 0.0), 6.0, 64, 36)
    r = cam.generate_rays()
    np.testing.assert_array_equal(r.pos.numpy(), d["pos"])
    np.testing.assert_array_equal(r.dir.numpy(), d["dir"])
    np.testing.assert_array_equal(r.intensity.numpy(), d["intensity"])
    # sub-pixel sampling extension: sample 0 is the reference ray; shards tile the pixel set
    r3 = cam.generate_rays(samples=3, seed=5)
    np.testing.assert_array_equal(r3.dir.numpy()[:64 * 36], d["dir"])
    a = cam.generate_rays(pixel_range=(0, 1000)).dir
    b = cam.generate_rays(pixel_range=(1000, 64 * 36)).dir
    np.testing.assert_array_equal(torch.cat([a, b]).numpy(), d["dir"])


def test_bundles_draw_the_reference_samples_under_the_same_seed(rtt_ns):
    """rays/bundle.py:30-56: theta first, then r; local->global pose.  A seeded script sees the same rays."""
    from oracle.make_golden import GOAL_RAYS, GOAL_SEED
    d = parity.load("extra_goals")
    _scene, _els, mk = _goal_setup(rtt_ns)
    torch.manual_seed(GOAL_SEED)
    for k, rot in enumerate((None, [0.02, 0.0, 0.0], [0.0, -0.03, 0.0])):
        r = mk(k, rot, "cpu").sample(GOAL_RAYS)
        np.testing.assert_array_equal(r.pos.numpy(), d[f"bundle{k}_pos"])
        np.testing.assert_array_equal(r.dir.numpy(), d[f"bundle{k}_dir"])
        assert r.id.dtype == torch.int8 and int(r.id[0]) == k


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["spot_size", "spot_size_target", "spot_target"])
def test_goals_match_reference_loss_and_gradients(rtt_ns, name):
    """optim/goals.py SpotSizeLoss / SpotTargetLoss driving the fused kernels (forward + adjoint) vs
    the reference's eager run: same loss, same d loss / d c1, c2.

    Goals only run on the reference's base ``Scene``, whose fp32 paths are decided by the
    last bit of sqrt (rays that re-hit the lens face they are leaving end up far from the
    spot and dominate a sum of square roots).  The bit-level bar is therefore the reference executed with
    a correctly rounded sqrt (``ieee_*`` keys, oracle/make_golden.py::ieee_sqrt) — the arithmetic the GPU
    implements; against the stock MKL-sqrt run the loss moves by ~0.3 %, the reference's own noise."""
    import raytracetorch_b200 as rtt
    from oracle.make_golden import GOAL_RAYS, GOAL_SEED
    d = parity.load("extra_goals")
    scene, elements, mk = _goal_setup(rtt_ns)
    scene = scene.cuda()
    sensor = elements[1]
    # sample on the CPU generator (like the reference fixture), then the goal moves the rays to the scene's device
    bundles = [mk(0, None, "cpu"), mk(1, [0.02, 0.0, 0.0], "cpu"), mk(2, [0.0, -0.03, 0.0], "cpu")]
    torch.manual_seed(GOAL_SEED)
    if name == "spot_size":
        loss = rtt.optim.SpotSizeLoss(sensor, bundles, N_rays=GOAL_RAYS)(scene)
    elif name == "spot_size_target":
        loss = rtt.optim.SpotSizeLoss(sensor, bundles, N_rays=GOAL_RAYS, target_xy=[0.1, -0.2])(scene)
    else:
        loss = rtt.optim.SpotTargetLoss(sensor, torch.tensor([[0.0, 0.0], [0.0, 2.0], [3.0, 0.0]]))(
            scene, bundles, N_rays=GOAL_RAYS)
    loss.backward()
    ref = float(d[f"ieee_{name}_loss"])
    assert abs(float(loss.detach()) - ref) <= 1e-4 * abs(ref)
    for k in (0, 1):
        g = elements[0].shape.surfaces[k].c.grad.cpu().numpy()
        assert parity.grad_rel(g, d[f"ieee_{name}_g_c{k}"]) < parity.TOL_GRAD, (k, g, d[f"ieee_{name}_g_c{k}"])
    stock = float(d[f"{name}_loss"])
    assert abs(float(loss.detach()) - stock) <= 2e-2 * abs(stock)


@pytest.mark.gpu
def test_device_bundles_are_generated_inside_the_trace_kernels(rtt_ns):
    """Bundle.sample on a CUDA device returns SourceRays; SequentialScene.simulate generates them in registers
    and gives bit-identical results to tracing the materialised bundle (rtt_sample_bundle)."""
    import raytracetorch_b200 as rtt
    import scenes
    from raytracetorch_b200.rays import SourceRays
    dev = torch.device("cuda", 0)
    for make in (lambda: rtt.rays.CollimatedDisk(5.0, 3, device=dev, transform=rtt.geom.RayTransformBundle(
                     translation=[0.0, 0.0, -10.0], rotation=[0.01, -0.02, 0.0]).to(dev)),
                 lambda: rtt.rays.PointSource(0.05, 1, device=dev, transform=rtt.geom.RayTransformBundle(
                     translation=[0.0, 0.0, -60.0]).to(dev))):
        bundle = make()
        r = bundle.sample(20_000)
        assert isinstance(r, SourceRays) and r.generated and len(r) == 20_000
        twin = SourceRays(r.source, r.pose, r.state, r.n, r.ray_id)        # same {key, counter}: the same rays
        pos0 = twin.pos.clone()                                           # materialises
        assert not twin.generated and twin.id.dtype == torch.int8 and int(twin.id[0]) == bundle.ray_id
        np.testing.assert_allclose(torch.linalg.norm(twin.dir, dim=1).cpu().numpy(), 1.0, atol=1e-6)
        outs = []
        for rays in (r, twin):
            els = scenes.c1_singlet(rtt_ns, physical=True)
            scene = rtt.scene.SequentialScene(els).to(dev)
            out = scene.simulate(rays)
            locs, w, ids = els[1].getHitsTensors()
            outs.append([t.cpu().numpy() for t in (out.pos, out.dir, out.intensity, locs, w, ids)])
        for a, b in zip(*outs):
            np.testing.assert_array_equal(a, b)
        assert outs[0][3].shape[0] > 1000 and not np.array_equal(outs[0][0], pos0.cpu().numpy())
    # consecutive samples advance the counter: different rays
    a, b = bundle.sample(100), bundle.sample(100)
    assert not torch.equal(a.dir, b.dir)           # (a point source: every position is the origin)


@pytest.mark.gpu
@pytest.mark.parametrize("target", [None, [0.1, -0.2]])
def test_fused_goal_equals_eager_goal_on_the_same_rays(rtt_ns, target, monkeypatch):
    """SpotSizeLoss through the fused path (in-kernel bundle, no final-ray outputs, rtt_spot_* reductions) vs the
    same goal evaluated with eager torch ops on the materialised records: same loss, same parameter gradients."""
    import raytracetorch_b200 as rtt
    import scenes
    from raytracetorch_b200 import optim, rays as R
    dev = torch.device("cuda", 0)
    res = []
    for fused in (True, False):
        R._SRC_STATE.clear()
        torch.manual_seed(11)
        els = scenes.c1_singlet(rtt_ns, physical=True, grads=True)
        scene = rtt.scene.SequentialScene(els).to(dev)
        bundles = [rtt.rays.CollimatedDisk(5.0, k, device=dev, transform=rtt.geom.RayTransformBundle(
            translation=[0.0, 0.0, -10.0], rotation=rot).to(dev)) for k, rot in enumerate((None, [0.02, 0.0, 0.0]))]
        if not fused:
            monkeypatch.setattr(optim, "_sensor_records", lambda *a: None)
        loss = rtt.optim.SpotSizeLoss(els[1], bundles, N_rays=50_000, target_xy=target)(scene)
        loss.backward()
        res.append((float(loss), [float(els[0].shape.surfaces[k].c.grad) for k in (0, 1)]))
    (l1, g1), (l2, g2) = res
    assert abs(l1 - l2) <= 2e-5 * abs(l2)
    for a, b in zip(g1, g2):
        assert abs(a - b) <= 1e-3 * abs(b), (g1, g2)


@pytest.mark.gpu
def test_goal_on_memory_resident_rays_skips_the_final_ray_outputs(rtt_ns):
    """A goal over rays that live in device memory (any Bundle whose sample() returns plain Rays): the trace writes no
    final rays (empty outputs, the Rays object is left as it was), loss and gradients equal the explicit evaluation."""
    import raytracetorch_b200 as rtt
    import scenes
    dev = torch.device("cuda", 0)
    n = 40_000
    g = torch.Generator().manual_seed(4)
    th = torch.rand(n, generator=g) * (2 * np.pi)
    r = torch.sqrt(torch.rand(n, generator=g)) * 5.0
    pos = torch.stack([r * torch.cos(th), r * torch.sin(th), torch.full_like(r, -10.0)], 1).contiguous().to(dev)
    dirs = torch.zeros(n, 3, device=dev)
    dirs[:, 2] = 1.0
    inten = torch.ones(n, device=dev)

    class Resident(rtt.rays.Bundle):
        def sample(self, N):
            return rtt.rays.Rays._wrap(pos=pos, dir=dirs, intensity=inten, wavelength=torch.zeros(n, device=dev),
                                       id=torch.zeros(n, dtype=torch.int8, device=dev))

    res = []
    for through_goal in (True, False):
        els = scenes.c1_singlet(rtt_ns, physical=True, grads=True)
        scene = rtt.scene.SequentialScene(els).to(dev)
        if through_goal:
            loss = rtt.optim.SpotSizeLoss(els[1], [Resident(0, device=dev)], N_rays=n)(scene)
            assert scene.last_trace["pos"].numel() == 0 and scene.rays.pos.data_ptr() == pos.data_ptr()
            assert scene.final_rays is True                  # restored after the evaluation
        else:
            out = rtt.ops.trace_sequential(scene.table(), pos, dirs, inten, None, want_record=True)
            assert out["pos"].shape == (n, 3)
            loss = rtt.ops.spot_size(out["records"].reshape(-1, 4), None)
        loss.backward()
        res.append((float(loss), [float(els[0].shape.surfaces[k].c.grad) for k in (0, 1)]))
    (l1, g1), (l2, g2) = res
    assert abs(l1 - l2) <= 1e-6 * abs(l2)
    for a, b in zip(g1, g2):
        assert abs(a - b) <= 1e-4 * abs(b), (g1, g2)


@pytest.mark.gpu
def test_spot_target_fused_equals_eager(rtt_ns, monkeypatch):
    import raytracetorch_b200 as rtt
    import scenes
    from raytracetorch_b200 import optim, rays as R
    dev = torch.device("cuda", 0)
    res = []
    for fused in (True, False):
        R._SRC_STATE.clear()
        torch.manual_seed(5)
        els = scenes.c1_singlet(rtt_ns, physical=True, grads=True)
        scene = rtt.scene.SequentialScene(els).to(dev)
        bundles = [rtt.rays.CollimatedDisk(5.0, k, device=dev, transform=rtt.geom.RayTransformBundle(
            translation=[0.0, 0.0, -10.0], rotation=rot).to(dev)) for k, rot in enumerate((None, [0.0, -0.03, 0.0]))]
        if not fused:
            monkeypatch.setattr(optim, "_sensor_records", lambda *a: None)
        goal = rtt.optim.SpotTargetLoss(els[1], torch.tensor([[0.0, 0.0], [3.0, 0.0]])).to(dev)
        loss = goal(scene, bundles, N_rays=30_000)
        loss.backward()
        res.append((float(loss), [float(els[0].shape.surfaces[k].c.grad) for k in (0, 1)]))
    (l1, g1), (l2, g2) = res
    assert abs(l1 - l2) <= 2e-5 * abs(l2) + 1e-9
    for a, b in zip(g1, g2):
        assert abs(a - b) <= 1e-3 * abs(b) + 1e-9, (g1, g2)


@pytest.mark.gpu
def test_renderer_render_3d_matches_reference(rtt_ns):
    """Renderer.render_3d (render/camera.py:191-257) on the C5 elements against the reference's own image: same
    winner surface, normal and shading per pixel; pixels whose nearest hit is decided at a silhouette edge may differ."""
    import raytracetorch_b200 as rtt
    import scenes
    ns = types.SimpleNamespace(elements=rtt.elements, geom=rtt.geom, phys=rtt.phys, rays=rtt.rays, scene=rtt.scene,
                               render=rtt.render)
    scene, cam = scenes.render_setup(ns, device="cuda")
    scene = scene.cuda()
    img = rtt.render.Renderer(scene).render_3d(cam).numpy()
    ref = parity.load("extra_render3d")["image"]
    assert img.shape == ref.shape == (64, 96, 3)
    diff = np.abs(img - ref).max(axis=2)
    assert (diff > 1e-4).mean() <= 0.005, f"{(diff > 1e-4).sum()} pixels differ"
    assert (np.abs(ref - 1.0).sum(-1) > 0).sum() > 800      # the fixture really shows the elements
    assert len(np.unique(np.round(ref.reshape(-1, 3), 3), axis=0)) > 20


@pytest.mark.gpu
def test_graphed_optimisation_step_follows_the_eager_loop(rtt_ns):
    """optim.GraphedStep: zero_grad -> SpotSizeLoss -> backward -> Adam captured in one CUDA graph.  Every replay
    draws fresh rays (device-side counter) and moves the parameters like the eager loop under the same seed."""
    import raytracetorch_b200 as rtt
    import scenes
    dev = torch.device("cuda", 0)

    def setup():
        els = scenes.c1_singlet(rtt_ns, physical=True, grads=True)
        scene = rtt.scene.SequentialScene(els).to(dev)
        bundle = rtt.rays.CollimatedDisk(5.0, 0, device=dev, transform=rtt.geom.RayTransformBundle(
            translation=[0.0, 0.0, -10.0]).to(dev))
        goal = rtt.optim.SpotSizeLoss(els[1], [bundle], N_rays=200_000)
        params = [p for p in scene.parameters() if p.requires_grad]
        opt = torch.optim.Adam(params, lr=1e-5, capturable=True)
        return scene, goal, opt, params

    warm, steps = 2, 5
    torch.manual_seed(7)
    scene, goal, opt, params = setup()
    losses_e = []
    for _ in range(warm + steps):
        opt.zero_grad(set_to_none=True)
        loss = goal(scene)
        loss.backward()
        opt.step()
        losses_e.append(float(loss))
    eager = [p.detach().clone() for p in params]
    torch.manual_seed(7)
    scene, goal, opt, params = setup()
    g = rtt.optim.GraphedStep(scene, goal, opt, warmup=warm)
    assert g.launches_per_step >= 5                         # trace + goal reductions + their adjoints
    losses_g = [float(g()) for _ in range(steps)]
    for a, b in zip(eager, params):
        assert torch.allclose(a, b.detach(), rtol=1e-5, atol=1e-9), (a, b)
    # the loss is steep here (it falls 3x in five steps): 1e-5 parameter differences show up as ~1e-3 in the loss
    np.testing.assert_allclose(losses_g, losses_e[warm:], rtol=3e-3)
    assert len(set(losses_g)) == steps                      # fresh rays on every replay


@pytest.mark.gpu
def test_bundle_transform_gradients_reach_the_source_pose(rtt_ns):
    """rays/bundle.py:30-37 + geom/transform.py:245-276: a Bundle whose pose is being optimised.  The
    reference's tests/test_ideal.py:142-168 differentiates an image position w.r.t. the source origin; here the
    device path must hand d loss / d trans, d rot_vec of the bundle to autograd (it used to drop them silently):
    the same Philox samples are drawn in the local frame by the kernel and posed with differentiable torch ops.
    Checked against oracle autograd on the same local samples, and against the analytic dZi/dZo of an ideal lens."""
    import raytracetorch_b200 as rtt
    import scenes
    from oracle import trace_oracle as O
    from raytracetorch_b200.rays import SourceRays
    dev = torch.device("cuda", 0)
    n = 30_000
    tr = rtt.geom.RayTransformBundle(translation=[0.3, -0.2, -10.0], rotation=[0.01, -0.02, 0.0],
                                     trans_grad=True, rot_grad=True).to(dev)
    bundle = rtt.rays.CollimatedDisk(4.0, 2, device=dev, transform=tr)
    rtt.rays._SRC_STATE.clear()
    torch.manual_seed(11)
    state0 = rtt.rays.source_state(dev).clone()
    rays = bundle.sample(n)
    assert not isinstance(rays, SourceRays) and rays.pos.requires_grad and rays.dir.requires_grad
    # with frozen pose parameters the in-kernel source is used, and it draws the same rays
    tr_frozen = rtt.geom.RayTransformBundle(translation=[0.3, -0.2, -10.0], rotation=[0.01, -0.02, 0.0]).to(dev)
    rtt.rays._SRC_STATE.clear()                      # re-key: the same seed again does not reset the device counter
    torch.manual_seed(11)
    twin = rtt.rays.CollimatedDisk(4.0, 2, device=dev, transform=tr_frozen).sample(n)
    assert isinstance(twin, SourceRays) and torch.equal(twin.state, state0)
    np.testing.assert_allclose(twin.pos.cpu().numpy(), rays.pos.detach().cpu().numpy(), atol=1e-5)
    np.testing.assert_allclose(twin.dir.cpu().numpy(), rays.dir.detach().cpu().numpy(), atol=2e-6)

    els = scenes.c1_singlet(rtt_ns, physical=True, grads=True)
    scene = rtt.scene.SequentialScene(els).to(dev)
    scene.simulate(rays)
    locs, wt, _ = els[1].getHitsTensors()
    loss = (wt * ((locs[:, 0] - 0.05) ** 2 + locs[:, 1] ** 2)).sum() / wt.sum()
    loss.backward()
    assert tr.trans.grad is not None and tr.rot_vec.grad is not None

    # oracle autograd on the same local samples (identity-pose twin), pose applied with the same torch ops
    local = SourceRays(twin.source, torch.tensor([1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0.0], device=dev), state0, n, 2)
    lp, ld = local.pos.cpu(), local.dir.cpu()
    tr_c = rtt.geom.RayTransformBundle(translation=[0.3, -0.2, -10.0], rotation=[0.01, -0.02, 0.0],
                                       trans_grad=True, rot_grad=True)
    p, d = tr_c.transform_(lp, ld)
    d = torch.nn.functional.normalize(d, p=2, dim=1)
    els_c = scenes.c1_singlet(rtt_ns, physical=True, grads=True)
    tab = rtt.scene.SequentialScene(els_c).table()
    o = O.trace_sequential(tab.f, tab.i_host, p, d, torch.ones(n))
    _m, hl, ww = o["sensor"][0]
    loss_o = (ww * ((hl[:, 0] - 0.05) ** 2 + hl[:, 1] ** 2)).sum() / ww.sum()
    loss_o.backward()
    assert abs(float(loss.detach()) - float(loss_o.detach())) <= 1e-4 * abs(float(loss_o.detach()))
    assert parity.grad_rel(tr.trans.grad.cpu().numpy(), tr_c.trans.grad.numpy()) < parity.TOL_GRAD
    assert parity.grad_rel(tr.rot_vec.grad.cpu().numpy(), tr_c.rot_vec.grad.numpy()) < parity.TOL_GRAD
    for k in (0, 1):
        assert parity.grad_rel(els[0].shape.surfaces[k].c.grad.cpu().numpy(),
                               els_c[0].shape.surfaces[k].c.grad.numpy()) < parity.TOL_GRAD


@pytest.mark.gpu
def test_source_origin_gradient_of_an_ideal_lens_is_the_axial_magnification(rtt_ns):
    """tests/test_ideal.py:142-186 of the reference on the device path: a point source at z_o = -3f in front of
    an ideal thin lens (f = 100) images at z_i = 1.5f; d z_i / d z_o = (z_i / z_o)^2 = 0.25.  The source origin is
    the bundle's `trans` Parameter."""
    import raytracetorch_b200 as rtt
    dev = torch.device("cuda", 0)
    tr = rtt.geom.RayTransformBundle(translation=[0.0, 0.0, -300.0], trans_grad=True).to(dev)
    bundle = rtt.rays.PointSource(0.02, 0, device=dev, transform=tr)
    torch.manual_seed(3)
    rays = bundle.sample(4096)
    lens = rtt.elements.IdealThinLens(focal=100.0).to(dev)
    out_pos, out_dir, _ = lens(rays, surf_idx=0)
    # axial crossing of each ray: the point on the ray closest to the z axis
    txy = -(out_pos[:, 0] * out_dir[:, 0] + out_pos[:, 1] * out_dir[:, 1]) / \
        (out_dir[:, 0] ** 2 + out_dir[:, 1] ** 2).clamp_min(1e-12)
    zi = out_pos[:, 2] + txy * out_dir[:, 2]
    off_axis = (out_dir[:, 0] ** 2 + out_dir[:, 1] ** 2) > 1e-8
    zi_mean = zi[off_axis].mean()
    zi_mean.backward()
    assert abs(float(zi_mean.detach()) - 150.0) < 0.5
    g = tr.trans.grad.cpu().numpy()
    assert abs(g[2] - 0.25) < 2e-3, g


def test_render_shade_entry_matches_the_reference_image(run_exact, rtt_ns):
    """rtt_render_shade (nearest hit + normal + shading in one call) against the reference's own render_3d image
    (fixture extra_render3d), on both back-ends: same winner surface, normal and colour per pixel; pixels whose
    nearest hit is decided at a silhouette edge may differ."""
    import raytracetorch_b200 as rtt
    import scenes
    from raytracetorch_b200 import render as RD
    ns = types.SimpleNamespace(elements=rtt.elements, geom=rtt.geom, phys=rtt.phys, rays=rtt.rays, scene=rtt.scene,
                               render=rtt.render)
    scene, cam = scenes.render_setup(ns, device="cpu")
    rays = cam.generate_rays()
    renderable = [el for el in scene.elements if not RD._is_aperture(el)]
    tab = rtt.compile_elements(renderable)
    base = torch.stack([RD._base_color(el.surface_functions[j]) for el in renderable for j in range(len(el.shape))])
    rd = RD.Renderer(scene)
    rgb, row = run_exact.render_shade(tab.f.detach().numpy(), tab.i.numpy(), rays.pos.numpy(), rays.dir.numpy(),
                                      base.numpy(), rd.light_dir.tolist(), rd.bg_color.tolist())
    ref = parity.load("extra_render3d")["image"]
    img = rgb.reshape(ref.shape)
    diff = np.abs(img - ref).max(axis=2)
    assert (diff > 1e-4).mean() <= 0.005, f"{(diff > 1e-4).sum()} pixels differ"
    hit = row.reshape(ref.shape[:2]) != 255
    assert hit.sum() > 800 and np.array_equal(hit, np.abs(ref - 1.0).sum(-1) > 0) or (hit != (np.abs(ref - 1.0).sum(-1) > 0)).mean() < 0.005
