"""In-kernel ray sources (rays/bundle.py:30-171, render/camera.py:39-72) and the fused goal reductions
(optim/goals.py:42-187) on both back-ends: host build of the per-ray source (CPU suite) and the CUDA library."""
import math

import numpy as np
import pytest
import torch

import parity
from raytracetorch_b200 import _cabi, codes as C

EYE_POSE = np.array([1, 0, 0, 0, 1, 0, 0, 0, 1, 0.5, -0.25, -10.0], np.float32)


def _tilted_pose():
    from raytracetorch_b200.geom import rotation_from_vector
    R = rotation_from_vector(torch.tensor([0.1, -0.2, 0.3])).numpy()
    return np.concatenate([R.reshape(-1), [1.0, 2.0, -3.0]]).astype(np.float32)


def test_disk_source_is_area_uniform_and_posed(run_exact):
    n = 200_000
    pose = _tilted_pose()
    R, T = pose[:9].reshape(3, 3), pose[9:]
    src = dict(kind=_cabi.SRC_DISK, a=[1.0, 25.0, 0.0, 2 * math.pi], pose=pose, seed=1234, first=0)
    r = run_exact.sample(src, n)
    local = (r["pos"] - T) @ R                    # invert pos @ R^T + T
    rad2 = local[:, 0] ** 2 + local[:, 1] ** 2
    assert np.abs(local[:, 2]).max() < 1e-5
    assert rad2.min() >= 1.0 - 1e-4 and rad2.max() <= 25.0 + 1e-4
    assert abs(rad2.mean() - 13.0) < 0.1          # r^2 ~ U(1, 25): area-uniform (rays/bundle.py:50)
    assert np.abs(local[:, :2].mean(0)).max() < 0.03
    want_dir = np.array([0, 0, 1], np.float32) @ R.T
    np.testing.assert_allclose(r["dir"], np.broadcast_to(want_dir, (n, 3)), atol=1e-6)   # + renormalisation
    assert np.all(r["intensity"] == 1.0) and np.all(r["wavelength"] == 0.0)
    # counter-based: a shard starting at ray 1000 reproduces rays [1000, 1100) of the full bundle
    part = run_exact.sample(dict(src, first=1000), 100)
    np.testing.assert_array_equal(part["pos"], r["pos"][1000:1100])
    # device-resident {key, counter} state replaces seed/first (CUDA-graph replay)
    st = run_exact.sample(dict(src, seed=0, first=0, state=[1234, 1000]), 100)
    np.testing.assert_array_equal(st["pos"], r["pos"][1000:1100])
    other = run_exact.sample(dict(src, seed=99), 100)
    assert np.abs(other["pos"] - r["pos"][:100]).max() > 0.1


def test_point_line_fan_sources(run_exact):
    n = 100_000
    NA = 0.3
    F_max = (1 - math.cos(math.asin(NA))) / math.pi          # rays/bundle.py:75-80, 153 (the reference's CDF)
    r = run_exact.sample(dict(kind=_cabi.SRC_POINT, a=[0.0, F_max, 0.0, 2 * math.pi], pose=EYE_POSE, seed=5), n)
    np.testing.assert_allclose(np.linalg.norm(r["dir"], axis=1), 1.0, atol=1e-6)
    cosphi = r["dir"][:, 2]
    assert cosphi.min() >= 1 - 2 * F_max - 1e-6 and cosphi.max() <= 1.0
    assert abs(cosphi.mean() - (1 - F_max)) < 2e-3            # cos(phi) = 1 - 2F, F uniform
    np.testing.assert_array_equal(r["pos"], np.broadcast_to(EYE_POSE[9:], (n, 3)))
    ln = run_exact.sample(dict(kind=_cabi.SRC_LINE, a=[3.0], pose=EYE_POSE, seed=5), n)
    x = ln["pos"][:, 0] - EYE_POSE[9]
    assert x.min() >= -3.0 and x.max() <= 3.0 and abs(x.mean()) < 0.03 and abs(x.var() - 3.0) < 0.05
    fan = run_exact.sample(dict(kind=_cabi.SRC_FAN, a=[0.25], pose=EYE_POSE, seed=5), n)
    th = np.arctan2(fan["dir"][:, 1], fan["dir"][:, 2])
    assert np.all(fan["dir"][:, 0] == 0) and th.min() >= -0.25 - 1e-6 and th.max() <= 0.25 + 1e-6


def test_camera_source_matches_generate_rays(run_exact):
    """Sample 0 of every pixel is the reference's pixel-centre ray (fixture from render/camera.py:39-72)."""
    import raytracetorch_b200 as rtt
    d = parity.load("extra_camera_rays")
    cam = rtt.render.Camera((0.0, 0.0, -200.0), (0.0, 0.0, 0.0), (0.0, 1.0, 0.0), 6.0, 64, 36)
    src = cam.source_spec()
    r = run_exact.sample(dict(src, pose=cam.source_pose().numpy(), seed=3), 3 * 64 * 36)
    np.testing.assert_allclose(r["dir"][:64 * 36], d["dir"], atol=3e-7)
    np.testing.assert_array_equal(r["pos"][:64 * 36], d["pos"])
    # jittered samples stay within half a pixel of the centre ray
    px = 2 * math.tan(math.radians(3.0)) * (64 / 36) / 63
    dev = np.abs(r["dir"][64 * 36:2 * 64 * 36] - d["dir"]).max()
    assert 0 < dev <= 0.51 * px * 1.01


@pytest.mark.parametrize("variant", ["exact", "fast"])
@pytest.mark.parametrize("name", ["c1_singlet_physical", "c2_cylindrical"])
def test_trace_from_source_equals_trace_of_its_rays(runner_of, name, variant):
    """The trace kernels generate the rays of a source in registers: same rays, bit for bit, as the bundle materialised
    by rtt_sample_bundle, hence the same trace — forward and adjoint.  EXACT: every output bit for bit.  FAST: the builds
    for generated rays and for rays in memory are separate instantiations (the ray-source code is compiled out of the
    latter), and the compiler contracts their multiply-adds independently: masks and intensities equal, points and
    directions to rounding (2e-6 relative)."""
    run_fast = runner_of(variant)
    d = parity.load(name)
    n = 5000
    pose = np.array([1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, -10.0], np.float32)
    src = dict(kind=_cabi.SRC_DISK, a=[0.0, 30.0, 0.0, 2 * math.pi], pose=pose, seed=77, first=12345)
    rays = run_fast.sample(src, n)
    a = run_fast.trace_seq(d["table_f"], d["table_i"], rays["pos"], rays["dir"], rays["intensity"], sensor_specs=[None])
    b = run_fast.trace_seq_src(d["table_f"], d["table_i"], src, n, sensor_specs=[None])
    for k in ("intensity", "hitmask"):
        np.testing.assert_array_equal(a[k], b[k])
    if variant == "exact":
        for k in ("pos", "dir"):
            np.testing.assert_array_equal(a[k], b[k])
        np.testing.assert_array_equal(a["sensors"][0][0], b["sensors"][0][0])
    else:
        scale = float(np.abs(a["pos"]).max())
        assert parity.vec_rel(b["pos"], a["pos"], floor=scale).max() <= 2e-6
        assert parity.vec_rel(b["dir"], a["dir"]).max() <= 2e-6
        np.testing.assert_allclose(b["sensors"][0][0], a["sensors"][0][0], rtol=0, atol=2e-6 * scale)
    assert (a["hitmask"] != 0).mean() > 0.5
    # records only: no final-ray outputs
    c = run_fast.trace_seq_src(d["table_f"], d["table_i"], src, n, sensor_specs=[None], want_rays=False)
    assert c["pos"] is None
    np.testing.assert_array_equal(c["sensors"][0][0], b["sensors"][0][0])      # same build, with and without ray outputs
    g_rec = np.random.default_rng(0).standard_normal((n, 4)).astype(np.float32)
    ti = d["table_i"].copy()
    ti[:, C.I_FLAGS] = C.FLAG_GRAD_CK | C.FLAG_GRAD_POSE_S | C.FLAG_GRAD_IOR        # ask for parameter gradients
    d = dict(d, table_i=ti)
    ga = run_fast.trace_seq_bwd(d["table_f"], d["table_i"], rays["pos"], rays["dir"], rays["intensity"], a["hitmask"],
                                None, None, None, g_records=[g_rec])
    gb = run_fast.trace_seq_src_bwd(d["table_f"], d["table_i"], src, n, a["hitmask"], [g_rec])
    scale = np.abs(ga["g_table"]).max()
    assert scale > 0
    np.testing.assert_allclose(gb["g_table"], ga["g_table"], rtol=1e-4, atol=1e-5 * scale)   # atomics: order only


def _records(n=20000, seed=0):
    g = torch.Generator().manual_seed(seed)
    rec = torch.randn((n, 4), generator=g)
    rec[:, 0] = rec[:, 0] * 0.3 + 0.1
    rec[:, 1] = rec[:, 1] * 0.2 - 0.05
    rec[:, 3] = torch.rand(n, generator=g) + 0.1
    dead = torch.rand(n, generator=g) < 0.3
    rec[dead] = 0.0                                        # rays that never reached the sensor
    rec[::97, 3] = -0.5                                    # negative weights are "inactive" for SpotSize only
    return rec


@pytest.mark.parametrize("active_only", [False, True])
def test_spot_moments_and_adjoint(run_exact, active_only):
    rec = _records().double().requires_grad_(True)
    w = rec[:, 3]
    on = (w > 0) if active_only else torch.ones_like(w, dtype=torch.bool)
    mom = torch.stack([(w * on).sum(), (rec[:, 0] * w * on).sum(), (rec[:, 1] * w * on).sum()])
    g3 = torch.tensor([0.3, -1.2, 0.7], dtype=torch.float64)
    (mom * g3).sum().backward()
    got = run_exact.spot_moments(rec.detach().float().numpy(), active_only)
    np.testing.assert_allclose(got[:3], mom.detach().numpy(), rtol=2e-6)
    assert got[3] == float((w > 0).sum())
    gg = run_exact.spot_moments_bwd(rec.detach().float().numpy(), active_only, g3.float().numpy())
    np.testing.assert_allclose(gg, rec.grad.numpy(), rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("target", [None, (0.05, -0.1)])
def test_spot_size_and_adjoint_match_torch_autograd(run_exact, target):
    """optim/goals.py:165-183 restated in torch (float64) with autograd as the checker."""
    rec0 = _records(seed=3)
    rec = rec0.double().requires_grad_(True)
    act = rec[:, 3] > 0
    xy, w = rec[act][:, :2], rec[act][:, 3]
    W = w.sum().clamp(min=1e-12)
    if target is None:
        cx, cy = (xy[:, 0] * w).sum() / W, (xy[:, 1] * w).sum() / W
    else:
        cx, cy = torch.tensor(target[0], dtype=torch.float64), torch.tensor(target[1], dtype=torch.float64)
    loss = torch.sqrt(((xy[:, 0] - cx) ** 2 + (xy[:, 1] - cy) ** 2) * (w / W)).sum()
    (2.5 * loss).backward()
    mom = run_exact.spot_moments(rec0.numpy(), True)
    out3, g = run_exact.spot_size(rec0.numpy(), mom, target, g_loss=2.5, repeat=3)
    assert abs(out3[0] - float(loss)) <= 3e-6 * float(loss)
    scale = rec.grad.abs().max().item()
    np.testing.assert_allclose(g, rec.grad.numpy(), rtol=2e-4, atol=2e-5 * scale)
    assert np.all(g[~act.numpy()] == 0)


def test_spot_reductions_on_empty_and_all_dead_records(run_exact):
    dead = np.zeros((1000, 4), np.float32)
    mom = run_exact.spot_moments(dead, True)
    np.testing.assert_array_equal(mom, np.zeros(4, np.float32))
    out3, g = run_exact.spot_size(dead, mom)
    assert out3[0] == 0 and np.all(g == 0) and np.all(np.isfinite(g))


# ---------------------------------------------------------------------------------------------
# per-id sensor moments: Sensor.getSpotSizeParallel_xy (elements/sensor.py:87-176)
# ---------------------------------------------------------------------------------------------
def _per_id_reference(rec, ids, query, targets, p):
    """The reference's computation restated with torch index ops (differentiable): result [K] in query order."""
    K = len(query)
    lut = torch.full((256,), -1, dtype=torch.int64)
    lut[torch.as_tensor(query) + 128] = torch.arange(K)
    grp = lut[ids.long() + 128]
    keep = grp >= 0
    xy, w, grp = rec[keep, :2], rec[keep, 3], grp[keep]
    W = torch.zeros(K, dtype=rec.dtype).index_add(0, grp, w)
    safe = torch.where(W == 0, torch.ones_like(W), W)
    if targets is None:
        c = torch.zeros(K, 2, dtype=rec.dtype).index_add(0, grp, xy * w[:, None]) / safe[:, None]
    else:
        c = torch.as_tensor(targets, dtype=rec.dtype)
    mom = (w[:, None] * (xy - c[grp]).abs() ** p).sum(1)
    return torch.zeros(K, dtype=rec.dtype).index_add(0, grp, mom) / (2 * safe), W


@pytest.mark.parametrize("p", [2.0, 3.0, 1.0])
@pytest.mark.parametrize("targets", [None, [[0.1, -0.2], [0.0, 0.3], [1.0, 1.0], [0.0, 0.0]]])
def test_spot_id_kernels_and_adjoint_match_torch_autograd(run_exact, targets, p):
    """rtt_spot_id_moments / _size / _size_bwd against the same sums taken with torch index_add in float64 and its
    autograd: ids in runs (bundles) and shuffled, an id that is not queried, a queried id without hits, dead records."""
    g = torch.Generator().manual_seed(4)
    m = 30_000
    ids = torch.cat([torch.full((m // 3,), k, dtype=torch.int8) for k in (0, 1, 5)])
    ids = torch.cat([ids, torch.randint(-3, 7, (5000,), generator=g, dtype=torch.int64).to(torch.int8)])
    M = ids.numel()
    rec = torch.randn(M, 4, generator=g, dtype=torch.float64) * 0.3
    rec[:, 3] = torch.rand(M, generator=g, dtype=torch.float64)
    rec[torch.rand(M, generator=g) < 0.3, 3] = 0.0                      # rays that never reached the sensor
    rec = rec.float().double()                                           # the values the fp32 kernels see
    query = [5, 0, 1, 100]                                               # 100: queried, never present
    rec64 = rec.clone().requires_grad_(True)
    want, W = _per_id_reference(rec64, ids, query, targets, p)
    g_out = torch.tensor([1.0, -0.5, 2.0, 0.7], dtype=torch.float64)
    (want * g_out).sum().backward()
    mom, s4, got, g_rec = run_exact.spot_id(rec.float().numpy(), ids.numpy(), query, targets, p, g_out.numpy(), repeat=2)
    np.testing.assert_allclose(mom[:, 0], W.detach().numpy(), rtol=2e-6)
    assert mom[3, 0] == 0 and got[3] == 0
    np.testing.assert_allclose(got, want.detach().numpy(), rtol=2e-5, atol=1e-9)
    ref_g = rec64.grad.numpy().copy()
    ref_g[:, 2] = 0.0
    assert parity.grad_rel(g_rec, ref_g) < 2e-5
    assert np.all(g_rec[~np.isin(ids.numpy(), query)] == 0)


def test_sensor_per_id_spot_sizes_match_the_reference_fixture(rtt_ns):
    """Sensor.getSpotSizeParallel_xy of this package (torch path: hit lists recorded on the CPU) against the
    UNMODIFIED reference's results on the multi-bundle fixture (oracle/make_golden.py::gen_spot_id_case): spot sizes
    in query order, intensity sums in sorted-id order, gradients to the lens curvatures through the oracle trace."""
    import raytracetorch_b200 as rtt
    import scenes
    from oracle import trace_oracle as O
    from oracle.make_golden import GOAL_RAYS
    d = parity.load("extra_spot_id")
    goals = parity.load("extra_goals")
    query = d["query"].tolist()
    for tag, tgt, p in (("centroid_p2", None, 2), ("target_p2", torch.from_numpy(d["targets"]), 2),
                        ("centroid_p3", None, 3)):
        els = scenes.c1_singlet(rtt_ns, physical=True, grads=True)
        sensor = els[1]
        tab = rtt.compile_elements(els)
        for k in range(3):
            pos, dr = torch.from_numpy(goals[f"bundle{k}_pos"]), torch.from_numpy(goals[f"bundle{k}_dir"])
            o = O.trace_sequential(tab.f, tab.i_host, pos, dr, torch.ones(GOAL_RAYS))
            mask, hl, w = o["sensor"][0]
            sensor.record(hl, w, torch.full((int(mask.sum()),), k, dtype=torch.int8))
        res, wsum = sensor.getSpotSizeParallel_xy(query, target_xy=tgt, norm_ord=p)
        res.sum().backward()
        np.testing.assert_allclose(res.detach().numpy(), d[f"{tag}_result"], rtol=2e-4)
        np.testing.assert_allclose(wsum.detach().numpy(), d[f"{tag}_intensity_sum"], rtol=1e-6)
        for k in (0, 1):
            assert parity.grad_rel(els[0].shape.surfaces[k].c.grad.numpy(), d[f"{tag}_g_c{k}"]) < parity.TOL_GRAD, (tag, k)


@pytest.mark.gpu
def test_sensor_per_id_spot_sizes_fused_path_matches_the_reference_fixture(rtt_ns):
    """The same fixture through the product path: three bundles traced by the fused CUDA kernel, per-id spot sizes by
    rtt_spot_id_* on the dense records (no hit lists), gradients through the hand-written adjoints."""
    import raytracetorch_b200 as rtt
    import scenes
    d = parity.load("extra_spot_id")
    goals = parity.load("extra_goals")
    query = d["query"].tolist()
    dev = torch.device("cuda", 0)
    for tag, tgt, p in (("centroid_p2", None, 2), ("target_p2", torch.from_numpy(d["targets"]), 2),
                        ("centroid_p3", None, 3)):
        els = scenes.c1_singlet(rtt_ns, physical=True, grads=True)
        sensor = els[1]
        scene = rtt.scene.SequentialScene(els).to(dev)
        for k in range(3):
            pos, dr = torch.from_numpy(goals[f"bundle{k}_pos"]).to(dev), torch.from_numpy(goals[f"bundle{k}_dir"]).to(dev)
            n = pos.shape[0]
            rays = rtt.rays.Rays._wrap(pos=pos, dir=dr, intensity=torch.ones(n, device=dev),
                                       id=torch.full((n,), k, dtype=torch.int8, device=dev),
                                       wavelength=torch.zeros(n, device=dev))
            scene.simulate(rays)
        assert sensor._dense_records() is not None                   # the fused path is the one that runs
        res, wsum = sensor.getSpotSizeParallel_xy(query, target_xy=tgt, norm_ord=p)
        res.sum().backward()
        np.testing.assert_allclose(res.detach().cpu().numpy(), d[f"{tag}_result"], rtol=2e-4)
        np.testing.assert_allclose(wsum.detach().cpu().numpy(), d[f"{tag}_intensity_sum"], rtol=1e-6)
        for k in (0, 1):
            g = els[0].shape.surfaces[k].c.grad.cpu().numpy()
            assert parity.grad_rel(g, d[f"{tag}_g_c{k}"]) < parity.TOL_GRAD, (tag, k, g, d[f"{tag}_g_c{k}"])
        locs, w, ids = sensor.getHitsTensors()                        # the lists are still available afterwards
        assert locs.shape[0] == int(d[f"{tag}_intensity_sum"].sum())
