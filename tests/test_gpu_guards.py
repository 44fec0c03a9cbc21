"""Out-of-bounds guard tests (compute-sanitizer is closed on this GPU pool): every output buffer of a launch is carved
out of one arena between guard zones filled with a sentinel; after the launch the guard zones must be untouched and the
results must equal an un-guarded launch.  Sizes straddle the 512-ray tile of the streaming kernel (bulk-async copies of
whole tiles, cooperative copy of the ragged last one); an unaligned carve makes the library fall back to plain stores."""
import ctypes as ct

import numpy as np
import pytest
import torch

import parity

pytestmark = pytest.mark.gpu
GUARD = 4096
SENTINEL = 0xA5


class Arena:
    def __init__(self, nbytes_list, misalign=0):
        self.spans = []
        off = GUARD + misalign
        for nb in nbytes_list:
            self.spans.append((off, nb))
            off += (nb + 15) // 16 * 16 + GUARD + misalign
        self.buf = torch.full((off + GUARD,), SENTINEL, dtype=torch.uint8, device="cuda")

    def view(self, k, dtype):
        off, nb = self.spans[k]
        return self.buf[off:off + nb].view(dtype)

    def guards_intact(self):
        keep = torch.ones(self.buf.numel(), dtype=torch.bool, device="cuda")
        for off, nb in self.spans:
            keep[off:off + nb] = False
        return bool((self.buf[keep] == SENTINEL).all())


@pytest.mark.parametrize("misalign", [0, 4])
@pytest.mark.parametrize("tune", [0, 12, 7, 8, 16, 18, 3, 5, 9])         # 0 / 12: the default (one 1024-thread block per SM)
@pytest.mark.parametrize("n", [1, 511, 512, 513, 1024, 1536, 2047, 2048, 2049, 20011])
def test_sequential_forward_stays_inside_its_output_buffers(n, tune, misalign):
    from raytracetorch_b200 import _cabi
    from gpusim import GpuSim, _dev, _p
    d = parity.load("c2_cylindrical")
    sim = GpuSim(tune << 16)
    reps = (n + d["in_pos"].shape[0] - 1) // d["in_pos"].shape[0]
    pos = np.tile(d["in_pos"], (reps, 1))[:n]
    dr = np.tile(d["in_dir"], (reps, 1))[:n]
    inten = np.tile(d["in_intensity"], reps)[:n]
    want = sim.trace_seq(d["table_f"], d["table_i"], pos, dr, inten)
    arena = Arena([12 * n, 12 * n, 4 * n, 8 * n], misalign=misalign)
    op, od, oi, mask = (arena.view(k, t) for k, t in enumerate((torch.float32, torch.float32, torch.float32, torch.int64)))
    req, hold = sim._table(d["table_f"], d["table_i"], None, None)
    p_, d_, i_ = _dev(pos), _dev(dr), _dev(inten)
    sim.lib.call("rtt_trace_seq_fwd", _p(p_), _p(d_), _p(i_), 0, None, op.data_ptr(), od.data_ptr(), oi.data_ptr(),
                 mask.data_ptr(), ct.byref(req), None, 0, n, sim.mode, sim._stream())
    torch.cuda.synchronize()
    assert arena.guards_intact(), "a kernel wrote outside its output buffers"
    np.testing.assert_array_equal(oi.cpu().numpy(), want["intensity"])
    np.testing.assert_array_equal(op.cpu().numpy().reshape(n, 3), want["pos"])
    np.testing.assert_array_equal(mask.cpu().numpy().view(np.uint64), want["hitmask"])


@pytest.mark.parametrize("n", [1, 4095, 4097, 70001])
def test_sequential_adjoint_stays_inside_its_output_buffers(n, rtt_ns):
    import raytracetorch_b200 as rtt
    import scenes
    from raytracetorch_b200 import codes as C
    from gpusim import GpuSim, _dev, _p
    tab = rtt.compile_elements(scenes.c2_cylindrical(rtt_ns, grads=True))      # curvature and pose gradients requested
    d = dict(parity.load("c2_cylindrical"))
    d["table_f"], d["table_i"] = tab.f.detach().numpy(), tab.i.numpy()
    sim = GpuSim(0)
    reps = (n + d["in_pos"].shape[0] - 1) // d["in_pos"].shape[0]
    pos = np.tile(d["in_pos"], (reps, 1))[:n]
    dr = np.tile(d["in_dir"], (reps, 1))[:n]
    inten = np.tile(d["in_intensity"], reps)[:n]
    fwd = sim.trace_seq(d["table_f"], d["table_i"], pos, dr, inten)
    gp, gd, gi = parity.golden_loss_grads(fwd["pos"], fwd["dir"], fwd["intensity"])
    want = sim.trace_seq_bwd(d["table_f"], d["table_i"], pos, dr, inten, fwd["hitmask"], gp, gd, gi)
    S = d["table_f"].shape[0]
    arena = Arena([12 * n, 12 * n, 4 * n, 4 * S * C.ROW_G])
    o = [arena.view(k, torch.float32) for k in range(4)]
    for t in o:
        t.zero_()
    req, hold = sim._table(d["table_f"], d["table_i"], None, None)
    args = [_dev(x) for x in (pos, dr, inten)]
    mask = _dev(np.asarray(fwd["hitmask"]).view(np.int64), torch.int64)
    g = [_dev(x) for x in (gp, gd, gi)]
    sim.lib.call("rtt_trace_seq_bwd", _p(args[0]), _p(args[1]), _p(args[2]), 0, None, _p(mask), _p(g[0]), _p(g[1]),
                 _p(g[2]), None, o[0].data_ptr(), o[1].data_ptr(), o[2].data_ptr(), o[3].data_ptr(), 0, ct.byref(req), 0, n,
                 sim.mode, sim._stream())
    torch.cuda.synchronize()
    assert arena.guards_intact(), "the adjoint wrote outside its output buffers"
    assert parity.grad_rel(o[0].cpu().numpy().reshape(n, 3), want["g_pos"]) < 1e-6
    assert parity.grad_rel(o[3].cpu().numpy().reshape(S, C.ROW_G), want["g_table"]) < 1e-4


@pytest.mark.parametrize("n", [1, 4095, 4097, 70001, (1 << 25) + 12345])
def test_lean_adjoint_stays_inside_its_gradient_table(n, rtt_ns):
    """The lean adjoint path (csrc/rtt_lean.cuh; scalar gradients, no ray-gradient outputs) in both launch shapes — four
    256-thread blocks per SM below 2^25 rays, one 1024-thread block per SM from there on: the gradient table sits between
    guard zones, and equals the general adjoint's (tune bit 8) to summation order."""
    import raytracetorch_b200 as rtt
    import scenes
    from raytracetorch_b200 import codes as C
    from gpusim import GpuSim, _dev, _p
    els = scenes.c2_cylindrical(rtt_ns)
    for el in els:
        for s_ in getattr(el.shape, "surfaces", []):
            if hasattr(s_, "c") and isinstance(s_.c, torch.nn.Parameter):
                s_.c.requires_grad_(True)
    tab = rtt.compile_elements(els)
    tf, ti = tab.f.detach().numpy(), tab.i.numpy()
    assert rtt.ops.adjoint_hint(tab) == rtt.ops.MODE_SCALAR_GRADS
    S = tf.shape[0]
    g = torch.Generator(device="cuda").manual_seed(11)
    th = torch.rand(n, device="cuda", generator=g) * (2 * np.pi)
    r = torch.sqrt(torch.rand(n, device="cuda", generator=g)) * 8.0
    pos = torch.stack([r * torch.cos(th), r * torch.sin(th), torch.full_like(r, -10.0)], 1).contiguous()
    del th, r
    dirs = torch.zeros_like(pos)
    dirs[:, 2] = 1.0
    inten = torch.ones(n, device="cuda")
    sim = GpuSim(0)
    req, hold = sim._table(tf, ti, None, None)
    opos, odir, oint = torch.empty_like(pos), torch.empty_like(dirs), torch.empty_like(inten)
    mask = torch.zeros(n, dtype=torch.int64, device="cuda")
    sim.lib.call("rtt_trace_seq_fwd", _p(pos), _p(dirs), _p(inten), 0, None, _p(opos), _p(odir), _p(oint), _p(mask),
                 ct.byref(req), None, 0, n, sim.mode, sim._stream())
    g_pos = torch.zeros_like(opos)
    g_pos[:, :2] = 2.0 * oint[:, None] * opos[:, :2]
    del odir
    out = {}
    for name, extra in (("lean", 0), ("general", 8 << 16)):
        arena = Arena([4 * S * C.ROW_G])
        gt = arena.view(0, torch.float32)
        gt.zero_()
        sim.lib.call("rtt_trace_seq_bwd", _p(pos), _p(dirs), _p(inten), 0, None, _p(mask), _p(g_pos), 0, 0, None,
                     0, 0, 0, gt.data_ptr(), 0, ct.byref(req), 0, n, sim.mode | rtt.ops.MODE_SCALAR_GRADS | extra,
                     sim._stream())
        torch.cuda.synchronize()
        assert arena.guards_intact(), f"the {name} adjoint wrote outside its gradient table"
        out[name] = gt.cpu().numpy().reshape(S, C.ROW_G).copy()
    if n > 1000:
        assert np.abs(out["general"][:, C.F_C]).sum() > 0
        assert parity.grad_rel(out["lean"][:, C.F_C:C.N_DIFF], out["general"][:, C.F_C:C.N_DIFF]) < 1e-4
    assert not out["lean"][:, :C.F_C].any()


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["c5_nonsequential", "x2_nonsequential"])
def test_nonsequential_forward_builds_agree_bit_for_bit(name):
    """The non-sequential forward in EXACT arithmetic: the default (one 1024-thread block per SM, a barrier per bounce
    trip), the free-running 256-thread kernel of round 1 (tune 7) and the other barrier placements (1, 3..6) run the same
    per-ray arithmetic — every output equal bit for bit, sensor images to accumulation order."""
    from gpusim import GpuSim
    d = parity.load(name)
    nb = int(d["nbounces"])
    reps = 40                                                # enough rays for several trips of every block
    pos, dr, inten = (np.tile(d[k], (reps,) + (1,) * (d[k].ndim - 1)) for k in ("in_pos", "in_dir", "in_intensity"))
    spec = [(64, 64, -5.0, 5.0, -5.0, 5.0, 1)]
    ref = GpuSim(1 | (7 << 16)).trace_nonseq(d["table_f"], d["table_i"], pos, dr, inten, nb, sensor_specs=spec)
    assert (ref["nb"] > 0).mean() > 0.3
    for tune in (0, 1, 3, 4, 5, 6):
        out = GpuSim(1 | (tune << 16)).trace_nonseq(d["table_f"], d["table_i"], pos, dr, inten, nb, sensor_specs=spec)
        for k in ("pos", "dir", "intensity", "seq", "nb"):
            np.testing.assert_array_equal(out[k], ref[k], err_msg=f"tune {tune}: {k}")
        np.testing.assert_array_equal(out["sensors"][0][0], ref["sensors"][0][0])
        img, img_ref = out["sensors"][0][1], ref["sensors"][0][1]
        assert np.abs(img - img_ref).sum() <= 1e-5 * max(np.abs(img_ref).sum(), 1.0)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["c2_cylindrical", "c4_camera_lens", "c1_singlet_physical"])
def test_sequential_forward_builds_agree(name):
    """FAST builds of the sequential forward against each other: the default (tile kernel, one persistent 1024-thread
    block per SM), the same kernel in 256- and 512-thread blocks, with a barrier per tile, and the packed-pair kernels —
    identical hit masks and intensities, points and directions to rounding (north_star: 1e-5; measured <= 2e-6)."""
    from gpusim import GpuSim
    d = parity.load(name)
    reps = 7
    pos, dr, inten = (np.tile(d[k], (reps,) + (1,) * (d[k].ndim - 1)) for k in ("in_pos", "in_dir", "in_intensity"))
    ref = GpuSim(3 << 16).trace_seq(d["table_f"], d["table_i"], pos, dr, inten)
    scale = float(np.abs(ref["pos"][np.isfinite(ref["pos"]).all(1)]).max())
    for tune in (0, 12, 7, 13, 6, 8, 5, 16, 18):
        out = GpuSim(tune << 16).trace_seq(d["table_f"], d["table_i"], pos, dr, inten)
        np.testing.assert_array_equal(out["hitmask"], ref["hitmask"], err_msg=f"tune {tune}")
        np.testing.assert_array_equal(out["intensity"], ref["intensity"], err_msg=f"tune {tune}")
        assert parity.vec_rel(out["pos"], ref["pos"], floor=scale).max() <= 2e-6, tune
        assert parity.vec_rel(out["dir"], ref["dir"]).max() <= 2e-6, tune
