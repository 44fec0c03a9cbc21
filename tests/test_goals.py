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
