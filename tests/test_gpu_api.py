"""GPU tests of the drop-in Python API (Element / Scene / SequentialScene / Rays) — the calls a
RayTraceTorch user makes — against the reference's golden fixtures and the oracle, plus
size-independent properties at BASELINE sizes.  Everything here goes
scene object -> torch.library op -> ctypes -> librtt_b200.so (CUDA)."""
import numpy as np
import pytest
import torch

import parity
import scenes
from oracle import trace_oracle as O

pytestmark = pytest.mark.gpu


def _cuda_rays(rtt, d, grads=False):
    p, dd, inten = (t.cuda() for t in parity.inputs_t(d))
    rays = rtt.rays.Rays._wrap(pos=p, dir=dd, intensity=inten,
                               id=torch.zeros(p.shape[0], dtype=torch.int8, device="cuda"),
                               wavelength=torch.zeros(p.shape[0], device="cuda"))
    if grads:
        for t in (rays.pos, rays.dir, rays.intensity):
            t.requires_grad_(True)
    return rays


def _build(rtt_ns, name):
    if name in scenes.CASES:
        builder, kw, mode, _ = scenes.CASES[name]
    else:
        builder, kw, _ = scenes.GRAD_CASES[name]
        mode = "seq"
    return builder(rtt_ns, **kw), mode


@pytest.mark.parametrize("name", parity.forward_names("seq"))
def test_sequential_scene_simulate_matches_reference(rtt_ns, name):
    import raytracetorch_b200 as rtt
    d = parity.load(name)
    els, _ = _build(rtt_ns, name)
    scene = rtt.scene.SequentialScene(els).cuda()
    rays = _cuda_rays(rtt, d)
    out = scene.simulate(rays)
    assert out is rays                                            # mutated in place, like the reference
    parity.assert_close_noise_aware(out.pos.cpu().numpy(), out.dir.cpu().numpy(), out.intensity.cpu().numpy(), d, name)
    # sensor hit lists: same hits, same order, weight = intensity before the sensor
    sensors = [e for e in els if type(e).__name__ == "Sensor"]
    if sensors and "f32_sensor0_loc" in d.files:
        locs, w, ids = sensors[0].getHitsTensors()
        assert locs.shape[0] == d["f32_sensor0_loc"].shape[0]
        np.testing.assert_array_equal(w.cpu().numpy(), d["f32_sensor0_w"])
        live = d["f32_sensor0_w"] > 0
        # sensor-local coordinates are small numbers (a focused spot sits at ~0): "relative" means
        # relative to the scene scale, i.e. to the global position of the same hit (|p| ~ 1e2)
        scale = float(np.abs(d["f32_pos"]).max())
        e = parity.vec_rel(locs.cpu().numpy()[live], d["f32_sensor0_loc"][live], floor=scale)
        noise = parity.vec_rel(d["f32_sensor0_loc"], d["f64_sensor0_loc"].astype(np.float32), floor=scale)[live] \
            if d["f64_sensor0_loc"].shape == d["f32_sensor0_loc"].shape else np.zeros_like(e)
        assert np.all(e <= np.maximum(parity.TOL_POINT, np.maximum(16 * noise, 2 * noise.max(initial=0.0))))


@pytest.mark.parametrize("name", parity.golden_names(grads=True))
def test_backward_through_autograd_function_matches_reference(rtt_ns, name):
    """loss.backward() on the fused op = the reference's autograd: every trainable Parameter and the
    input rays (tests/test_optimize_singlet.py:66-116 is this loop)."""
    import raytracetorch_b200 as rtt
    d = parity.load(name)
    els, _ = _build(rtt_ns, name)
    scene = rtt.scene.SequentialScene(els).cuda()
    rays = _cuda_rays(rtt, d, grads=True)
    leaf = (rays.pos, rays.dir, rays.intensity)
    out = scene.simulate(rays)
    loss = parity.golden_loss(out.pos, out.dir, out.intensity)
    loss.backward()
    assert abs(float(loss.detach()) - float(d["f32_loss"])) <= 1e-5 * abs(float(d["f32_loss"]))
    for t, k in zip(leaf, ("g_pos", "g_dir", "g_intensity")):
        assert parity.grad_rel(t.grad.cpu().numpy(), d["f32_" + k]) < parity.TOL_GRAD, k
    params = dict(scene.named_parameters())
    for k in [k[len("f32_gp::"):] for k in d.files if k.startswith("f32_gp::")]:
        ref = d["f64_gp::" + k]
        g = params[k].grad
        g = np.zeros_like(ref) if g is None else g.cpu().numpy()
        if np.linalg.norm(ref) == 0:
            assert np.linalg.norm(g) == 0, k
        else:
            assert parity.grad_rel(g, ref) < parity.TOL_GRAD, (k, g, ref)


@pytest.mark.parametrize("name", parity.forward_names("nonseq"))
def test_nonsequential_scene_matches_reference_on_stable_rays(rtt_ns, name):
    import raytracetorch_b200 as rtt
    d = parity.load(name)
    els, _ = _build(rtt_ns, name)
    scene = rtt.scene.Scene()
    for e in els:
        scene.add_element(e)
    scene = scene.cuda()
    scene.Nbounces = int(d["nbounces"])
    scene.rays = _cuda_rays(rtt, d)
    scene.simulate()
    seq = scene.last_trace["hit_seq"].cpu().numpy().astype(np.int64)
    seq[seq == 255] = -1
    stable = parity.stable_nonseq_rows(d) & parity.self_hit_free(d, torch.from_numpy(d["table_f"]), d["table_i"].tolist())
    np.testing.assert_array_equal(seq[stable], d["f32_seq"][stable])
    r = scene.rays
    parity.assert_close_noise_aware(r.pos.cpu().numpy(), r.dir.cpu().numpy(), r.intensity.cpu().numpy(), d, name,
                                    rows=stable)
    # ray_cast: winners of the first bounce as (element, surface) ids
    scene.rays = _cuda_rays(rtt, d)
    hit_mask, we, ws = scene.ray_cast(scene.rays)
    first = d["f32_seq"][:, 0]
    np.testing.assert_array_equal(hit_mask.cpu().numpy(), first >= 0)
    flat = torch.tensor(d["table_i"][:, 8:10])
    sel = first >= 0
    np.testing.assert_array_equal(we.cpu().numpy()[sel], flat[first[sel], 0].numpy())
    np.testing.assert_array_equal(ws.cpu().numpy()[sel], flat[first[sel], 1].numpy())


def test_element_forward_and_intersect_test_like_the_reference_scripts(rtt_ns):
    """Direct lens(rays, surf_idx) calls and the projection loss of tests/test_optimize_singlet.py:80-106."""
    import raytracetorch_b200 as rtt
    d = parity.load("grad_c3_singlet")
    els = scenes.c1_singlet(rtt_ns, physical=True, grads=True)
    lens = els[0].cuda()
    rays = _cuda_rays(rtt, d)
    p1, d1, _ = lens(rays, 0)
    rays.pos, rays.dir = p1, d1
    p2, d2, _ = lens(rays, 1)
    t = (100.0 - p2[:, 2]) / (d2[:, 2] + 1e-6)
    x, y = p2[:, 0] + t * d2[:, 0], p2[:, 1] + t * d2[:, 1]
    loss = (x ** 2 + y ** 2).mean()
    loss.backward()
    # oracle: same two element steps + same loss through torch autograd
    els_o = scenes.c1_singlet(rtt_ns, physical=True, grads=True)
    tab = rtt.compile_elements(els_o)
    rows = O.make_rows(tab.f, tab.i_host)
    p, dd, _ = parity.inputs_t(d)
    h1, n1, *_ = O.element_step(rows, 0, p, dd)
    h2, n2, *_ = O.element_step(rows, 1, h1, n1)
    to = (100.0 - h2[:, 2]) / (n2[:, 2] + 1e-6)
    lo = ((h2[:, 0] + to * n2[:, 0]) ** 2 + (h2[:, 1] + to * n2[:, 1]) ** 2).mean()
    lo.backward()
    assert abs(float(loss.detach()) - float(lo.detach())) <= 1e-5 * abs(float(lo.detach()))
    for k in (0, 1):
        g = lens.shape.surfaces[k].c.grad.cpu().numpy()
        go = els_o[0].shape.surfaces[k].c.grad.numpy()
        assert parity.grad_rel(g, go) < parity.TOL_GRAD
    tm = lens.intersectTest(_cuda_rays(rtt, d))
    assert tm.shape == (d["in_pos"].shape[0], 3)
    with torch.no_grad():
        ref = torch.stack([O.intersect_row(rows, r, p, dd) for r in range(3)], 1).numpy()
    np.testing.assert_array_equal(np.isfinite(tm.cpu().numpy()), np.isfinite(ref))


def test_sensor_image_and_wavelength_channels(rtt_ns):
    """C2: 3 wavelengths -> 3-channel image; equals the oracle's histogram of the reference hit list."""
    import raytracetorch_b200 as rtt
    n = 200_000
    els = scenes.c2_cylindrical(rtt_ns)
    disp = rtt.Dispersion(scenes.C2_WAVELENGTHS, {
        els[0].ior_glass: [1.5 * s for s in scenes.C2_GLASS_SCALE],
        els[1].ior_glass: [1.6 * s for s in scenes.C2_GLASS_SCALE]})
    els[3].set_image(256, 256, channels=3)
    scene = rtt.scene.SequentialScene(els)
    scene.set_dispersion(disp)
    scene = scene.cuda()
    rays = scenes.make_bundle(rtt_ns, ("coll", 8.0, -10.0, None), n, 7)
    wav = torch.tensor(scenes.C2_WAVELENGTHS)[torch.arange(n) % 3]
    rays.wavelength = wav
    cpu = (rays.pos.clone(), rays.dir.clone(), rays.intensity.clone())
    rays = rays.to("cuda")
    scene.simulate(rays)
    img = els[3].image.cpu().numpy()
    assert img.shape == (3, 256, 256)
    tab = rtt.compile_elements([e.cpu() for e in els], dispersion=disp)
    o = O.trace_sequential(tab.f, tab.i_host, *cpu, wavelength=wav, lut=tab.lut, lut_w=tab.lut_wavelengths)
    mask, hl, w = o["sensor"][0]
    ref = O.sensor_image(hl, w, els[3].image_spec, channel=(torch.arange(n) % 3)[mask]).numpy()
    assert ref.sum() > 0.3 * n
    assert parity.rel_l1(img, ref) <= parity.TOL_IMAGE_L1
    assert abs(img.sum() - ref.sum()) <= 1e-6 * ref.sum()


@pytest.mark.parametrize("n", [10 ** 6, 10 ** 8])
def test_full_size_properties_sequential(rtt_ns, n):
    """BASELINE sizes (C1: 1e6, C2: 1e8 rays): properties that need no CPU run of the same size —
    chunking invariance (rays are independent), determinism, image = checksum of the hit list,
    and a random sub-sample traced by the oracle."""
    import raytracetorch_b200 as rtt
    els = scenes.c2_cylindrical(rtt_ns)
    els[3].set_image(1024, 1024)
    scene = rtt.scene.SequentialScene(els).cuda()
    scene.record_hits = False
    g = torch.Generator(device="cuda").manual_seed(11)
    th = torch.rand(n, device="cuda", generator=g) * (2 * np.pi)
    r = torch.sqrt(torch.rand(n, device="cuda", generator=g)) * 8.0
    pos = torch.stack([r * torch.cos(th), r * torch.sin(th), torch.full_like(r, -10.0)], 1)
    del th, r
    dirs = torch.zeros_like(pos)
    dirs[:, 2] = 1.0
    inten = torch.ones(n, device="cuda")
    tab = scene.table()
    out = rtt.ops.trace_sequential(tab, pos, dirs, inten, want_record=False)
    img = out["images"][0]
    # (1) determinism of per-ray outputs
    out2 = rtt.ops.trace_sequential(tab, pos, dirs, inten, want_record=False)
    assert torch.equal(out["pos"], out2["pos"]) and torch.equal(out["hitmask"], out2["hitmask"])
    # (2) chunking invariance: tracing two halves == tracing the whole bundle
    h = n // 2
    a = rtt.ops.trace_sequential(tab, pos[:h], dirs[:h], inten[:h], want_record=False)
    b = rtt.ops.trace_sequential(tab, pos[h:], dirs[h:], inten[h:], want_record=False)
    # FAST mode: a ray's arithmetic may round differently depending on its slot in a launch tile (the
    # compiler contracts the unrolled copies independently), so halves == whole to rounding, with
    # boundary-tie flips allowed on <= 1e-6 of the rays; EXACT mode is bit-identical (checked below)
    cat_pos, cat_int = torch.cat([a["pos"], b["pos"]]), torch.cat([a["intensity"], b["intensity"]])
    same = cat_int == out["intensity"]
    assert float((~same).float().mean()) <= 1e-6
    assert float((cat_pos - out["pos"])[same].abs().max()) <= 1e-6 * 50.0
    if n <= 10 ** 6:
        ex = [rtt.ops.trace_sequential(tab, pos[s], dirs[s], inten[s], want_record=False, mode=rtt.ops.MODE_EXACT)
              for s in (slice(None), slice(0, h), slice(h, None))]
        assert torch.equal(torch.cat([ex[1]["pos"], ex[2]["pos"]]), ex[0]["pos"])
        assert torch.equal(torch.cat([ex[1]["intensity"], ex[2]["intensity"]]), ex[0]["intensity"])
        del ex
    del cat_pos, cat_int, same
    img_halves = a["images"][0] + b["images"][0]
    assert float((img_halves - img).abs().sum() / img.sum()) < parity.TOL_IMAGE_L1
    del a, b, out2
    # (3) checksum: image total == total weight of rays that hit the sensor inside the extent
    srow = tab.sensor_rows[0]
    hit = ((out["hitmask"] >> srow) & 1).bool()
    inside = hit & (out["pos"][:, 0].abs() < 15.0) & (out["pos"][:, 1].abs() < 15.0)
    total = float(out["intensity"][inside].double().sum())      # weight = intensity at the sensor (Transmit)
    assert abs(float(img.double().sum()) - total) <= 1e-4 * total
    # (4) oracle on a random sub-sample
    idx = torch.randint(0, n, (20000,), device="cuda", generator=g)
    tabc = rtt.compile_elements([e.cpu() for e in scenes.c2_cylindrical(rtt_ns)])
    o = O.trace_sequential(tabc.f, tabc.i_host, pos[idx].cpu(), dirs[idx].cpu(), inten[idx].cpu())
    np.testing.assert_array_equal(out["intensity"][idx].cpu().numpy(), o["intensity"].numpy())
    live = o["intensity"].numpy() > 0
    assert parity.vec_rel(out["pos"][idx].cpu().numpy()[live], o["pos"].numpy()[live]).max() <= parity.TOL_POINT
    S = tabc.n_rows
    np.testing.assert_array_equal(parity.mask_bits(out["hitmask"][idx].cpu().numpy().view(np.uint64), S),
                                  o["hit"].numpy())


def test_full_size_properties_adjoint(rtt_ns):
    """The adjoint at BASELINE config 3's size (1e7 rays per step): linearity in the upstream gradient, additivity over
    ray chunks (rays are independent, the parameter gradient is a sum over rays), run-to-run agreement to accumulation
    order, and oracle autograd on a random sub-sample (per-ray input gradients, parameter gradients of that sample)."""
    import raytracetorch_b200 as rtt
    n = 10_000_000
    els = scenes.c1_singlet(rtt_ns, physical=True, grads=True)
    scene = rtt.scene.SequentialScene(els).cuda()
    g = torch.Generator(device="cuda").manual_seed(21)
    th = torch.rand(n, device="cuda", generator=g) * (2 * np.pi)
    r = torch.sqrt(torch.rand(n, device="cuda", generator=g)) * 5.0
    pos = torch.stack([r * torch.cos(th), r * torch.sin(th), torch.full_like(r, -10.0)], 1).contiguous()
    del th, r
    dirs = torch.zeros_like(pos)
    dirs[:, 2] = 1.0
    inten = torch.ones(n, device="cuda")
    tab = scene.table()
    mode = rtt.ops.get_default_mode()
    hint = rtt.ops.adjoint_hint(tab)
    tf = tab.f.detach()
    fwd = torch.ops.rtt_b200.trace_seq_fwd(pos, dirs, inten, None, tf, tab.i, None, None, [], False, mode)
    opos, odir, oint, hm = fwd[:4]
    g1 = torch.zeros_like(opos)
    g1[:, :2] = 2.0 * oint[:, None] * opos[:, :2]
    g2 = torch.randn(opos.shape, device="cuda", generator=g) * oint[:, None]
    gi = (opos[:, :2] ** 2).sum(1)

    def bwd(sl, gp, gint, need_rays=False):
        return torch.ops.rtt_b200.trace_seq_bwd(pos[sl], dirs[sl], inten[sl], None, hm[sl], gp[sl].contiguous(), None,
                                                None if gint is None else gint[sl], None, tf, tab.i, None, None,
                                                need_rays, True, mode | hint)

    full = slice(None)
    a, b, ab = bwd(full, g1, gi)[3], bwd(full, g2, None)[3], bwd(full, g1 + g2, gi)[3]
    assert float(a.abs().sum()) > 0 and float(b.abs().sum()) > 0
    assert parity.grad_rel((a + b).cpu().numpy(), ab.cpu().numpy()) < 1e-4                 # linearity
    again = bwd(full, g1, gi)[3]
    assert parity.grad_rel(again.cpu().numpy(), a.cpu().numpy()) < 1e-5                     # accumulation order only
    h = n // 2 + 12345
    halves = bwd(slice(0, h), g1, gi)[3] + bwd(slice(h, n), g1, gi)[3]
    assert parity.grad_rel(halves.cpu().numpy(), a.cpu().numpy()) < 1e-4                    # additivity over chunks
    # oracle autograd on a random sub-sample
    idx = torch.randint(0, n, (20000,), device="cuda", generator=g)
    elc = scenes.c1_singlet(rtt_ns, physical=True, grads=True)
    tabc = rtt.compile_elements(elc)
    p, dd, ii = (t[idx].cpu().clone().requires_grad_(True) for t in (pos, dirs, inten))
    o = O.trace_sequential(tabc.f, tabc.i_host, p, dd, ii)
    (o["intensity"] * (o["pos"][:, :2] ** 2).sum(1)).sum().backward()
    sub = torch.ops.rtt_b200.trace_seq_bwd(pos[idx], dirs[idx], inten[idx], None, hm[idx], g1[idx].contiguous(), None,
                                           gi[idx], None, tf, tab.i, None, None, True, True, mode | hint)
    assert parity.grad_rel(sub[0].cpu().numpy(), p.grad.numpy()) < parity.TOL_GRAD
    assert parity.grad_rel(sub[1].cpu().numpy(), dd.grad.numpy()) < parity.TOL_GRAD
    assert parity.grad_rel(sub[2].cpu().numpy(), ii.grad.numpy()) < parity.TOL_GRAD
    tab.f.backward(sub[3])
    for k in (0, 1):
        ref = elc[0].shape.surfaces[k].c.grad.numpy()
        got = els[0].shape.surfaces[k].c.grad.cpu().numpy()
        assert parity.grad_rel(got, ref) < parity.TOL_GRAD, (k, got, ref)


def test_full_size_nonsequential_properties(rtt_ns):
    """C5 at 2e7 rays: determinism, chunking invariance, first-bounce winners vs the oracle."""
    import raytracetorch_b200 as rtt
    n = 20_000_000
    els = scenes.c5_nonsequential(rtt_ns)
    scene = rtt.scene.Scene()
    for e in els:
        scene.add_element(e)
    scene = scene.cuda()
    tab = scene.table()
    g = torch.Generator(device="cuda").manual_seed(5)
    th = torch.rand(n, device="cuda", generator=g) * (2 * np.pi)
    r = torch.sqrt(torch.rand(n, device="cuda", generator=g)) * 10.0
    pos = torch.stack([r * torch.cos(th), r * torch.sin(th), torch.full_like(r, -5.0)], 1)
    dirs = torch.zeros_like(pos)
    dirs[:, 2] = 1.0
    inten = torch.ones(n, device="cuda")
    out = rtt.ops.trace_nonsequential(tab, pos, dirs, inten, 8, want_record=False)
    out2 = rtt.ops.trace_nonsequential(tab, pos, dirs, inten, 8, want_record=False)
    assert torch.equal(out["hit_seq"], out2["hit_seq"]) and torch.equal(out["pos"], out2["pos"])
    h = n // 3
    a = rtt.ops.trace_nonsequential(tab, pos[:h], dirs[:h], inten[:h], 8, want_record=False)
    assert torch.equal(a["hit_seq"], out["hit_seq"][:h]) and torch.equal(a["pos"], out["pos"][:h])
    idx = torch.randint(0, n, (5000,), device="cuda", generator=g)
    tabc = rtt.compile_elements([e.cpu() for e in scenes.c5_nonsequential(rtt_ns)])
    O.IEEE_SQRT = True
    try:
        o = O.trace_nonsequential(tabc.f, tabc.i_host, pos[idx].cpu(), dirs[idx].cpu(), inten[idx].cpu(), 1)
    finally:
        O.IEEE_SQRT = False
    first = out["hit_seq"][idx, 0].cpu().numpy().astype(np.int64)
    first[first == 255] = -1
    np.testing.assert_array_equal(first, o["seq"].numpy()[:, 0])
    nh = out["n_hits"].cpu().numpy()
    assert nh.max() <= 8 and (out["hit_seq"].cpu().numpy() != 255).sum(1).tolist()[:1000] == nh.tolist()[:1000]


def test_host_resident_bundle_is_streamed_and_equals_the_device_trace(rtt_ns):
    """SequentialScene.simulate on Rays that live in (pinned) host memory: H2D chunks pipelined with per-chunk
    launches (ops.trace_sequential_host).  Same kernel, chunk starts on tile boundaries => per-ray outputs are
    bit-identical to tracing the device-resident bundle in one launch; images agree to accumulation order."""
    import raytracetorch_b200 as rtt
    n = 300_001
    g = torch.Generator().manual_seed(5)
    th = torch.rand(n, generator=g) * (2 * np.pi)
    r = torch.sqrt(torch.rand(n, generator=g)) * 8.0
    pos = torch.stack([r * torch.cos(th), r * torch.sin(th), torch.full_like(r, -10.0)], 1).contiguous().pin_memory()
    dirs = torch.zeros(n, 3)
    dirs[:, 2] = 1.0
    dirs, inten = dirs.pin_memory(), torch.ones(n).pin_memory()
    lam = torch.tensor(scenes.C2_WAVELENGTHS)[torch.arange(n) % 3].contiguous().pin_memory()
    ids = (torch.arange(n) % 5).to(torch.int8).pin_memory()

    def make():
        els = scenes.c2_cylindrical(rtt_ns)
        disp = rtt.Dispersion(scenes.C2_WAVELENGTHS, {els[0].ior_glass: [1.5 * s for s in scenes.C2_GLASS_SCALE],
                                                      els[1].ior_glass: [1.6 * s for s in scenes.C2_GLASS_SCALE]})
        els[3].set_image(256, 256, channels=3)
        scene = rtt.scene.SequentialScene(els).cuda()
        scene.set_dispersion(disp)
        return scene, els

    # functional level, small chunks (multiple of the 512-ray launch tile)
    scene, els = make()
    tab = scene.table()
    a = rtt.ops.trace_sequential_host(tab, pos, dirs, inten, lam, want_record=True, chunk_rays=65536, ids=ids)
    b = rtt.ops.trace_sequential(tab, pos.cuda(), dirs.cuda(), inten.cuda(), lam.cuda(), want_record=True)
    for k in ("pos", "dir", "intensity", "hitmask", "records"):
        assert torch.equal(a[k], b[k]), k
    assert parity.rel_l1(a["images"][0].cpu().numpy(), b["images"][0].cpu().numpy()) <= 1e-6
    assert torch.equal(a["in_id"].cpu(), ids) and torch.equal(a["in_wavelength"].cpu(), lam)
    # object level: host Rays in, device Rays out, sensor hit lists as from the device path
    outs = []
    for host in (True, False):
        scene, els = make()
        mv = (lambda t: t) if host else (lambda t: t.cuda())
        rays = rtt.rays.Rays._wrap(pos=mv(pos), dir=mv(dirs), intensity=mv(inten), id=mv(ids), wavelength=mv(lam))
        out = scene.simulate(rays)
        assert out is rays and out.pos.is_cuda and out.id.is_cuda and out.wavelength.is_cuda
        locs, w, hid = els[3].getHitsTensors()
        outs.append([t.cpu() for t in (out.pos, out.intensity, locs, w, hid)])
    for x, y in zip(*outs):
        assert torch.equal(x, y)


def test_adjoint_builds_and_compaction_agree(rtt_ns):
    """The sequential adjoint has four routes to the same parameter gradients: with / without the block-level
    compaction to rays that carry an upstream gradient (no ray gradients requested / requested), and the builds with /
    without pose-gradient code (RTT_MODE_SCALAR_GRADS).  On a bundle where most rays are dead or masked out they must
    agree to accumulation order, and so must the ray gradients of the two builds."""
    import raytracetorch_b200 as rtt
    n = 400_003
    g = torch.Generator(device="cuda").manual_seed(11)
    th = torch.rand(n, device="cuda", generator=g) * (2 * np.pi)
    r = torch.sqrt(torch.rand(n, device="cuda", generator=g)) * 8.0
    pos = torch.stack([r * torch.cos(th), r * torch.sin(th), torch.full_like(r, -10.0)], 1).contiguous()
    dirs = torch.zeros(n, 3, device="cuda")
    dirs[:, 2] = 1.0
    inten = torch.ones(n, device="cuda")
    lam = torch.tensor(scenes.C2_WAVELENGTHS, device="cuda")[torch.arange(n, device="cuda") % 3].contiguous()
    els = scenes.c2_cylindrical(rtt_ns)
    for el in els:
        for s_ in getattr(el.shape, "surfaces", []):
            if hasattr(s_, "c") and isinstance(s_.c, torch.nn.Parameter):
                s_.c.requires_grad_(True)
    scene = rtt.scene.SequentialScene(els).cuda()
    scene.set_dispersion(rtt.Dispersion(scenes.C2_WAVELENGTHS, {
        els[0].ior_glass: [1.5 * s_ for s_ in scenes.C2_GLASS_SCALE],
        els[1].ior_glass: [1.6 * s_ for s_ in scenes.C2_GLASS_SCALE]}))
    tab = scene.table()
    assert rtt.ops.adjoint_hint(tab) == rtt.ops.MODE_SCALAR_GRADS
    mode = rtt.ops.get_default_mode()
    tf = tab.f.detach()
    fwd = torch.ops.rtt_b200.trace_seq_fwd(pos, dirs, inten, lam, tf, tab.i, tab.lut, tab.lut_wavelengths, [], False, mode)
    opos, _odir, oint, hitmask = fwd[:4]
    keep = (torch.rand(n, device="cuda", generator=g) < 0.6).float()          # 40 % of the survivors masked out too
    g_pos = torch.zeros_like(opos)
    g_pos[:, :2] = 2.0 * (oint * keep)[:, None] * opos[:, :2]
    g_int = (opos[:, :2] ** 2).sum(1)
    live = float((g_pos != 0).any(1).float().mean())
    assert 0.1 < live < 0.4

    def run(need_rays, hint):
        return torch.ops.rtt_b200.trace_seq_bwd(pos, dirs, inten, lam, hitmask, g_pos, None, g_int, None, tf, tab.i,
                                                tab.lut, tab.lut_wavelengths, need_rays, True, mode | hint)

    ref = run(True, 0)
    assert float(ref[3].abs().sum()) > 0
    for need_rays in (False, True):
        for hint in (0, rtt.ops.MODE_SCALAR_GRADS):
            out = run(need_rays, hint)
            assert parity.grad_rel(out[3].cpu().numpy(), ref[3].cpu().numpy()) < 2e-5, (need_rays, hint)
            assert parity.grad_rel(out[4].cpu().numpy(), ref[4].cpu().numpy()) < 2e-5
            if need_rays:                                   # FAST arithmetic: the two builds may contract differently
                for k in (0, 1, 2):
                    assert parity.grad_rel(out[k].cpu().numpy(), ref[k].cpu().numpy()) < 1e-5, (k, hint)


def test_host_resident_bundle_nonsequential_equals_the_device_trace(rtt_ns):
    """Scene.simulate on Rays in (pinned) host memory: H2D chunks pipelined with per-chunk bounce-loop launches
    (ops.trace_nonsequential_host).  Rays are independent and the entry runs EXACT arithmetic, so every per-ray output
    is bit-identical to tracing the device-resident bundle in one launch; images agree to accumulation order."""
    import raytracetorch_b200 as rtt
    n = 200_003
    g = torch.Generator().manual_seed(9)
    th = torch.rand(n, generator=g) * (2 * np.pi)
    r = torch.sqrt(torch.rand(n, generator=g)) * 10.0
    pos = torch.stack([r * torch.cos(th), r * torch.sin(th), torch.full_like(r, -5.0)], 1).contiguous().pin_memory()
    dirs = torch.zeros(n, 3)
    dirs[:, 2] = 1.0
    dirs, inten = dirs.pin_memory(), torch.ones(n).pin_memory()
    lam = torch.full((n,), 0.55).pin_memory()
    ids = (torch.arange(n) % 7).to(torch.int8).pin_memory()

    def make():
        els = scenes.c5_nonsequential(rtt_ns)
        els[4].set_image(128, 128)
        scene = rtt.scene.Scene()
        for e in els:
            scene.add_element(e)
        scene.Nbounces = 6
        return scene.cuda(), els

    scene, els = make()
    tab = scene.table()
    a = rtt.ops.trace_nonsequential_host(tab, pos, dirs, inten, 6, lam, want_record=True, record_depth=2,
                                         chunk_rays=50_000, ids=ids)
    b = rtt.ops.trace_nonsequential(tab, pos.cuda(), dirs.cuda(), inten.cuda(), 6, lam.cuda(), want_record=True,
                                    record_depth=2)
    for k in ("pos", "dir", "intensity", "hit_seq", "n_hits", "records", "sensor_counts"):
        assert a[k].shape == b[k].shape and torch.equal(a[k], b[k]), k
    assert parity.rel_l1(a["images"][0].cpu().numpy(), b["images"][0].cpu().numpy()) <= 1e-6
    assert torch.equal(a["in_id"].cpu(), ids) and torch.equal(a["in_wavelength"].cpu(), lam)
    outs = []
    for host in (True, False):
        scene, els = make()
        mv = (lambda t: t) if host else (lambda t: t.cuda())
        scene.rays = rtt.rays.Rays._wrap(pos=mv(pos), dir=mv(dirs), intensity=mv(inten), id=mv(ids), wavelength=mv(lam))
        scene.simulate()
        out = scene.rays
        assert out.pos.is_cuda and out.id.is_cuda and out.wavelength.is_cuda
        locs, w, hid = els[4].getHitsTensors()
        outs.append([t.cpu() for t in (out.pos, out.intensity, locs, w, hid)])
    for x, y in zip(*outs):
        assert torch.equal(x, y)


def test_paths_proxy_records_one_snapshot_per_bounce(rtt_ns):
    """rays/ray.py:100-225 ``Paths``: the GUI's per-bounce position history.  Each ``Scene.step()`` that hit
    something appends one CPU snapshot; ``simulate()`` on a Paths runs bounce by bounce and ends where the fused
    multi-bounce launch ends (same per-bounce arithmetic)."""
    import raytracetorch_b200 as rtt
    from raytracetorch_b200.rays import Paths

    def make():
        scene = rtt.scene.Scene()
        for e in scenes.c5_nonsequential(rtt_ns):
            scene.add_element(e)
        scene.Nbounces = 6
        return scene.cuda()

    rays0 = scenes.make_bundle(rtt_ns, ("coll", 10.0, -5.0, None), 20_000, 5).to("cuda")
    # step loop with the proxy vs the same loop on plain rays
    a, b = make(), make()
    a.rays, b.rays = Paths(rays0.clone()), rays0.clone()
    expect = [rays0.pos.cpu()]
    for _ in range(4):
        a.step()
        b.step()
        expect.append(b.rays.pos.cpu())
    hist = a.rays.get_history()
    assert len(hist) == 5 and all(h.device.type == "cpu" and h.shape == (20_000, 3) for h in hist)
    for h, e in zip(hist, expect):
        assert torch.equal(h, e)
    assert not torch.equal(hist[1], hist[0]) and not torch.equal(hist[2], hist[1])
    plain = a.rays.unwrap()
    assert isinstance(plain, rtt.rays.Rays) and torch.equal(plain.pos.cpu(), hist[-1])
    # simulate(): bounce-by-bounce with the proxy == one fused launch without it
    c, d = make(), make()
    c.rays, d.rays = Paths(rays0.clone()), rays0.clone()
    c.simulate()
    d.simulate()
    assert torch.equal(c.rays.pos, d.rays.pos) and torch.equal(c.rays.intensity, d.rays.intensity)
    assert 2 <= len(c.rays.get_history()) <= 7


def test_fresnel_lens_through_the_scene_api(rtt_ns):
    """SingletLens(fresnel=True) (elements/lens.py:35-38) in a SequentialScene: every simulate() draws a fresh seed
    from torch's generator (reproducible under torch.manual_seed, pinned by scene.rng_seed), forward + backward run
    through the same branches and match the oracle under that seed."""
    import raytracetorch_b200 as rtt
    rays0 = scenes.make_bundle(rtt_ns, ("coll", 9.0, -10.0, [0.1, 0.05, 0.0]), 30_000, 4).to("cuda")

    def run(seed_fn):
        els = scenes.c1_singlet(rtt_ns, physical=True, inked=False, fresnel=True, grads=True)
        scene = rtt.scene.SequentialScene(els).cuda()
        seed_fn(scene)
        out = scene.simulate(rays0.clone())
        return scene, els, out

    def pinned(s):
        s.rng_seed = 5

    _, _, a = run(pinned)
    _, _, b = run(pinned)
    assert torch.equal(a.dir, b.dir)
    _, _, c = run(lambda s: setattr(s, "rng_seed", 6))
    assert not torch.equal(a.dir, c.dir)
    torch.manual_seed(123)
    _, _, d1 = run(lambda s: None)
    _, _, d2 = run(lambda s: None)                        # next draw of the generator: different branches
    torch.manual_seed(123)
    _, _, d3 = run(lambda s: None)
    assert torch.equal(d1.dir, d3.dir) and not torch.equal(d1.dir, d2.dir)
    back = float((a.dir[:, 2] < 0).float().mean())
    assert 0.03 < back < 0.2
    # gradients through the branches taken, against the oracle under the same seed
    scene, els, out = run(pinned)
    loss = parity.golden_loss(out.pos, out.dir, out.intensity)
    loss.backward()
    els_c = scenes.c1_singlet(rtt_ns, physical=True, inked=False, fresnel=True, grads=True)
    tab = rtt.compile_elements(els_c).with_seed(5)
    rc = rays0.to("cpu")
    o = O.trace_sequential(tab.f, tab.i_host, rc.pos, rc.dir, rc.intensity)
    parity.golden_loss(o["pos"], o["dir"], o["intensity"]).backward()
    for k in (0, 1):
        g, r = float(els[0].shape.surfaces[k].c.grad), float(els_c[0].shape.surfaces[k].c.grad)
        assert abs(g - r) <= parity.TOL_GRAD * abs(r), (k, g, r)


# ---------------------------------------------------------------------------------------------
# the BENCHMARKED configurations at full size (bench.py builds exactly these)
# ---------------------------------------------------------------------------------------------
def _bench_module():
    import importlib
    import os
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if root not in sys.path:
        sys.path.insert(0, root)
    return importlib.import_module("bench")


def test_full_size_c2_as_benchmarked_with_wavelength_table_and_three_channel_image(rtt_ns):
    """BASELINE configs[1] exactly as bench.py times it: 1e8 rays, 14 rows, per-wavelength index table (3 lines),
    3x1024x1024 sensor image, FAST arithmetic.  Per-channel image checksum against the hit masks, determinism, and
    the oracle (with the same table) on a random sub-sample: alive masks / hit masks exact, points 1e-5, the
    sub-sample's own 3-channel image within 1e-4 relative L1 with identical bin occupancy away from ties."""
    import raytracetorch_b200 as rtt
    bench = _bench_module()
    dev = torch.device("cuda", 0)
    w = bench.build_workload("c2", dev)
    n = w["rays"]
    assert n == 10 ** 8
    scene = rtt.scene.SequentialScene(w["elements"])
    scene.set_dispersion(w["dispersion"])
    scene = scene.to(dev)
    scene.record_hits = False
    tab = scene.table()
    assert tab.n_rows == 14 and tab.lut is not None and tab.lut.shape[0] == 3
    cfg = rtt.ops.sensor_cfg_of(tab)
    pos, dirs, inten, wav = bench.synth_bundle(w, n, dev, 1000)
    out = rtt.ops.trace_sequential(tab, pos, dirs, inten, wav, want_record=False, sensor_cfg=cfg)
    img = out["images"][0]
    assert img.shape == (3, 1024, 1024)
    # (1) per-channel checksum: channel c holds the weight of the rays of wavelength c that hit the sensor in range
    srow = tab.sensor_rows[0]
    hit = ((out["hitmask"] >> srow) & 1).bool()
    inside = hit & (out["pos"][:, 0].abs() < 15.0) & (out["pos"][:, 1].abs() < 15.0)
    lam_idx = torch.arange(n, device=dev) % 3
    for c in range(3):
        total = float(out["intensity"][inside & (lam_idx == c)].double().sum())
        assert total > 0.1 * n / 3
        assert abs(float(img[c].double().sum()) - total) <= 1e-4 * total, c
    # (2) determinism of the per-ray outputs, image equal up to accumulation order
    out2 = rtt.ops.trace_sequential(tab, pos, dirs, inten, wav, want_record=False, sensor_cfg=cfg)
    assert torch.equal(out["pos"], out2["pos"]) and torch.equal(out["hitmask"], out2["hitmask"])
    assert float((out2["images"][0] - img).abs().sum() / img.sum()) <= 1e-5
    del out2, inside, hit
    # (3) oracle with the same wavelength table on a random sub-sample
    g = torch.Generator(device=dev).manual_seed(3)
    idx = torch.randint(0, n, (30000,), device=dev, generator=g)
    els_c = _bench_module().build_workload("c2", "cpu")
    tabc = rtt.compile_elements([e.cpu() for e in els_c["elements"]], dispersion=els_c["dispersion"])
    sp, sd, si, sw = (t[idx].cpu() for t in (pos, dirs, inten, wav))
    o = O.trace_sequential(tabc.f, tabc.i_host, sp, sd, si, wavelength=sw, lut=tabc.lut, lut_w=tabc.lut_wavelengths)
    np.testing.assert_array_equal(out["intensity"][idx].cpu().numpy(), o["intensity"].numpy())
    np.testing.assert_array_equal(parity.mask_bits(out["hitmask"][idx].cpu().numpy().view(np.uint64), 14),
                                  o["hit"].numpy())
    live = o["intensity"].numpy() > 0
    assert live.mean() > 0.3
    assert parity.vec_rel(out["pos"][idx].cpu().numpy()[live], o["pos"].numpy()[live]).max() <= parity.TOL_POINT
    assert parity.vec_rel(out["dir"][idx].cpu().numpy()[live], o["dir"].numpy()[live]).max() <= parity.TOL_POINT
    # the sub-sample's own image through the kernel vs the oracle's histogram
    sub = rtt.ops.trace_sequential(tab, pos[idx].contiguous(), dirs[idx].contiguous(), inten[idx].contiguous(),
                                   wav[idx].contiguous(), want_record=False, sensor_cfg=cfg)
    mask, hl, ww = o["sensor"][0]
    ch = (idx.cpu() % 3)[mask]
    ref = O.sensor_image(hl, ww, w["sensor"].image_spec, channel=ch).numpy()
    got = sub["images"][0].cpu().numpy()
    assert ref.sum() > 0.3 * idx.numel()
    assert parity.rel_l1(got, ref) <= parity.TOL_IMAGE_L1
    assert ((got > 0) != (ref > 0)).sum() <= 2            # bin indices exact away from measure-zero ties


def test_full_size_c4_camera_render_as_benchmarked(rtt_ns):
    """BASELINE configs[3] as bench.py --workload c4cam runs it on one GPU: 1.24e8 pinhole-camera rays generated in
    the kernel (15 samples per pixel of a 3840x2160 camera), 17-row lens, 4K sensor image.  The image must be
    additive over sample ranges (the multi-GPU sharding), reproducible, and equal to the ORACLE's histogram on
    sub-ranges of the same Philox counters (rays materialised by rtt_sample_bundle, traced on the CPU)."""
    import raytracetorch_b200 as rtt
    bench = _bench_module()
    dev = torch.device("cuda", 0)
    w = bench.build_workload("c4cam", dev)
    n = w["rays"]
    assert n == 3840 * 2160 * 15
    scene = rtt.scene.SequentialScene(w["elements"]).to(dev)
    tab = scene.table()
    assert tab.n_rows == 17
    cfg = rtt.ops.sensor_cfg_of(tab)
    cam = rtt.render.Camera((0.0, 0.0, -200.0), (0.0, 0.0, 0.0), (0.0, 1.0, 0.0), 6.0, 3840, 2160, device=dev)

    def render(first, count):
        src = cam.generate_source_rays(samples=15, seed=1234, first=first, count=count)
        return rtt.ops.trace_sequential(tab, want_record=False, sensor_cfg=cfg, source=src, want_rays=False)

    full = render(0, n)
    img = full["images"][0]
    assert img.shape == (1, 2160, 3840)
    tot = float(img.double().sum())
    srow = tab.sensor_rows[0]
    hits = float(((full["hitmask"] >> srow) & 1).double().sum())
    assert 0.05 * n < tot <= hits                           # unit weights: image total = rays binned on the sensor
    again = render(0, n)["images"][0]
    assert float((again - img).abs().sum()) <= 1e-5 * tot
    h = (n // 3 // 256) * 256 + 77                          # an uneven split, like a shard boundary
    parts = render(0, h)["images"][0] + render(h, n - h)["images"][0]
    assert float((parts - img).abs().sum()) <= 1e-5 * tot
    del again, parts
    # oracle on sub-ranges of the same counters: sample 0 (pixel centres) and two jittered samples
    npix = 3840 * 2160
    elc = _bench_module().build_workload("c4cam", "cpu")["elements"]
    tabc = rtt.compile_elements([e.cpu() for e in elc])
    checked = 0
    for first in (npix // 2 - 10000, 7 * npix + npix // 2 + 1234, 14 * npix + npix // 2 - 20000):
        m = 20000
        src = cam.generate_source_rays(samples=15, seed=1234, first=first, count=m)
        twin = rtt.rays.SourceRays(src.source, src.pose, src.state, src.n, 0)
        sp, sd, si = twin.pos.cpu(), twin.dir.cpu(), twin.intensity.cpu()        # rtt_sample_bundle, same counters
        o = O.trace_sequential(tabc.f, tabc.i_host, sp, sd, si)
        sub = rtt.ops.trace_sequential(tab, want_record=False, sensor_cfg=cfg, source=src, want_rays=True)
        np.testing.assert_array_equal(sub["intensity"].cpu().numpy(), o["intensity"].numpy())
        np.testing.assert_array_equal(parity.mask_bits(sub["hitmask"].cpu().numpy().view(np.uint64), 17),
                                      o["hit"].numpy())
        np.testing.assert_array_equal(full["hitmask"][first:first + m].cpu().numpy(), sub["hitmask"].cpu().numpy())
        mask, hl, ww = o["sensor"][0]
        ref = O.sensor_image(hl, ww, w["sensor"].image_spec).numpy()
        got = sub["images"][0].cpu().numpy()
        if ref.sum() > 0:
            # unit weights: |got - ref|.sum() = 2 x (rays binned differently).  FAST arithmetic places a hit within
            # ~1e-6 of the oracle's; a 4K bin is 6.25e-3 wide, so a fraction ~3e-4 of the rays sits within rounding of
            # a bin edge (the camera's regular grid imaged onto the regular bin grid makes such near-ties systematic).
            # Bin indices must be exact AWAY from those ties: at most 1e-3 of the rays may move, and only to the
            # neighbouring bin — the images agree to 1e-4 once 2x2 bins are pooled with every alignment.
            moved = np.abs(got - ref)
            assert moved.sum() <= 2 * max(3, 1e-3 * ref.sum()), (moved.sum(), ref.sum())
            best = 1.0
            for oy in (0, 1):
                for ox in (0, 1):
                    g2 = got[0, oy:2160 - 2 + oy, ox:3840 - 2 + ox].reshape(1079, 2, 1919, 2).sum((1, 3))
                    r2 = ref[0, oy:2160 - 2 + oy, ox:3840 - 2 + ox].reshape(1079, 2, 1919, 2).sum((1, 3))
                    best = min(best, float(np.abs(g2 - r2).sum() / ref.sum()))
            assert best <= 2e-3, best
            checked += int(ref.sum())
    assert checked > 3000


def test_full_size_nonsequential_whole_hit_sequences(rtt_ns):
    """C5 at 2e7 rays, 8 bounces: WHOLE hit sequences, bounce counts, final positions and intensities of a random
    sub-sample against the oracle, on the rays whose path does not depend on the t > 1e-6 threshold
    (tests/parity.py::self_hit_free); the fraction of such rays is pinned."""
    import raytracetorch_b200 as rtt
    n, nb = 20_000_000, 8
    els = scenes.c5_nonsequential(rtt_ns)
    scene = rtt.scene.Scene()
    for e in els:
        scene.add_element(e)
    scene = scene.cuda()
    tab = scene.table()
    g = torch.Generator(device="cuda").manual_seed(5)
    th = torch.rand(n, device="cuda", generator=g) * (2 * np.pi)
    r = torch.sqrt(torch.rand(n, device="cuda", generator=g)) * 10.0
    pos = torch.stack([r * torch.cos(th), r * torch.sin(th), torch.full_like(r, -5.0)], 1)
    dirs = torch.zeros_like(pos)
    dirs[:, 2] = 1.0
    inten = torch.ones(n, device="cuda")
    out = rtt.ops.trace_nonsequential(tab, pos, dirs, inten, nb, want_record=False)
    idx = torch.randint(0, n, (6000,), device="cuda", generator=g)
    tabc = rtt.compile_elements([e.cpu() for e in scenes.c5_nonsequential(rtt_ns)])
    sp, sd, si = pos[idx].cpu(), dirs[idx].cpu(), inten[idx].cpu()
    O.IEEE_SQRT = True
    try:
        o = O.trace_nonsequential(tabc.f, tabc.i_host, sp, sd, si, nb)
    finally:
        O.IEEE_SQRT = False
    clean = parity.self_hit_free(dict(in_pos=sp.numpy(), in_dir=sd.numpy(), in_intensity=si.numpy(), nbounces=nb),
                                 tabc.f, tabc.i_host, nbounces=nb)
    assert abs(clean.mean() - parity.CLEAN_FRACTION["c5_nonsequential"]) < 0.03, clean.mean()
    seq = out["hit_seq"][idx].cpu().numpy().astype(np.int64)
    seq[seq == 255] = -1
    np.testing.assert_array_equal(seq[clean], o["seq"].numpy()[clean])
    np.testing.assert_array_equal(out["n_hits"][idx].cpu().numpy()[clean], o["nb"].numpy()[clean])
    np.testing.assert_array_equal(out["intensity"][idx].cpu().numpy()[clean], o["intensity"].numpy()[clean])
    assert parity.vec_rel(out["pos"][idx].cpu().numpy()[clean], o["pos"].numpy()[clean]).max() <= parity.TOL_POINT
    assert (seq == o["seq"].numpy()).all(axis=1).mean() > 0.99       # and almost everywhere on the noisy rays too
