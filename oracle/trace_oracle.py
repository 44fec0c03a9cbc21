"""CPU ORACLE — test infrastructure, not product code.

A table-driven restatement, in eager torch on the CPU, of the reference's batched
ray-propagation path.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import this file; the
product package ``raytracetorch_b200`` never does (``tests/test_no_oracle_in_product.py``).

It consumes the same flat surface table the CUDA kernels consume
(``raytracetorch_b200/codes.py``) and follows the reference's arithmetic step by step, with
the same torch primitives where rounding matters (``@`` for poses, ``** 2``,
``F.normalize``, ``torch.norm``), so that in fp32 it reproduces the reference to rounding
and in fp64 it serves as the high-precision yardstick.  It is differentiable by torch
autograd (masked gather / ``index_put`` like ``rays/ray.py:29-40``), which is what the
adjoint kernel's gradients are checked against.

PARITY PINNING: ``oracle/make_golden.py`` runs the UNMODIFIED reference (through
``oracle/ref_loader.py``) on seeded inputs and stores its outputs under ``tests/golden/``;
``tests/test_oracle_golden.py`` checks this oracle (and the scene compiler) against those
fixtures on every CPU test run.  The reference's own analytic known answers
(``tests/test_primitive.py:121-128,166-242,244-307``) are asserted in
``tests/test_known_answers.py``.

Reference lines followed (all under /root/reference):
  pose                      geom/transform.py:90-93 (global->local), :116-117 (inverse)
  renormalisation           rays/ray.py:22-25 via with_coords (geom/shape.py:38,75)
  root selection            geom/primitives.py:28-36 (unbounded), geom/bounded.py:20-36 (bounded)
  Plane                     geom/primitives.py:124-143
  Sphere                    geom/primitives.py:155-187
  Cylinder                  geom/primitives.py:201-241
  Quadric / QuadricZY       geom/primitives.py:266-343, 356-395
  bounds                    geom/bounded.py:60-64, 77-82, 98-106, 123-127, 151-159, 171-174
  shape validity            geom/shape.py:47-55, 122-132; geom/spherics.py:27-46; geom/cylindrics.py:23-55
  surface/shape forward     geom/primitives.py:59-96, geom/shape.py:61-87
  physics                   phys/std.py:97-108, 123-145, 227-254; phys/filter.py:24-33
  element step              elements/parent.py:44-58; sensor record elements/sensor.py:22-39
  sequential loop           scene/sequential.py:12-36
  non-sequential loop       scene/base.py:129-235
  sensor image              gui/workbench.py:615-624 (np.histogram2d with fixed range)
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch
import torch.nn.functional as F

# layout constants are data, shared with the product package on purpose
from raytracetorch_b200 import codes as C

INF = float("inf")
EPS_T = 1e-6      # geom/primitives.py:6
EPS_S = 1e-6      # Surface.epsilon, geom/primitives.py:21
EPS_HALF = 1e-6   # HalfSphere/HalfCyl bound slack (geom/bounded.py:127,174 use intersectEpsilon too)


# torch's CPU sqrt goes through MKL VML (high-accuracy mode: < 1 ulp, NOT correctly rounded), so
# its last bit is library dependent.  IEEE_SQRT=True replaces it by a correctly rounded fp32
# sqrt (computed in double, rounded once) — the arithmetic every IEEE device, including the
# CUDA kernels' sqrt.rn.f32, implements.  Default False = exactly what the reference executes on
# this host; the parity tests use True when they compare bit for bit against EXACT-mode kernels.
IEEE_SQRT = False


def _sqrt(x):
    if IEEE_SQRT and x.dtype == torch.float32:
        return torch.sqrt(x.double()).float()
    return torch.sqrt(x)


class Row:
    """Typed view of one table row."""

    def __init__(self, f: torch.Tensor, i: List[int]):
        self.f, self.i = f, i
        self.Re, self.Te = f[C.F_RE:C.F_RE + 9].view(3, 3), f[C.F_TE:C.F_TE + 3]
        self.Rs, self.Ts = f[C.F_RS:C.F_RS + 9].view(3, 3), f[C.F_TS:C.F_TS + 3]
        self.c, self.k, self.radius = f[C.F_C], f[C.F_K], f[C.F_RADIUS]
        self.ior_in, self.ior_out = f[C.F_IOR_IN], f[C.F_IOR_OUT]
        self.sb, self.hb = f[C.F_SB:C.F_SB + 4].detach(), f[C.F_HB:C.F_HB + 8].detach()
        self.surf, self.bound, self.invert = i[C.I_SURF], i[C.I_BOUND], bool(i[C.I_INVERT])
        self.shape, self.phys, self.sensor = i[C.I_SHAPE], i[C.I_PHYS], i[C.I_SENSOR]
        self.poly_first, self.poly_count = i[C.I_POLY_FIRST], i[C.I_POLY_COUNT]


def _pose(p, d, R, T):
    return (p - T[None, :]) @ R, d @ R            # geom/transform.py:90-93


def _roots(row: Row, o, d):
    """Candidate distances in the surface frame; list of [N] tensors."""
    if row.surf == C.SURF_PLANE:                  # geom/primitives.py:124-136
        dz = d[:, 2]
        safe = torch.where(torch.abs(dz) < EPS_S, torch.full_like(dz, 1e-8), dz)
        return [-o[:, 2] / safe]
    if row.surf == C.SURF_SPHERE:                 # geom/primitives.py:155-184
        b = 2.0 * torch.sum(o * d, dim=1)
        cc = torch.sum(o * o, dim=1) - row.radius ** 2
        disc = b ** 2 - 4 * cc
        ok = disc >= 0
        sq = _sqrt(torch.where(ok, disc, torch.zeros_like(disc)))
        inf = torch.full_like(b, INF)
        return [torch.where(ok, (-b - sq) / 2.0, inf), torch.where(ok, (-b + sq) / 2.0, inf)]
    if row.surf == C.SURF_CYLINDER:               # geom/primitives.py:201-231
        ox, oy, dx, dy = o[:, 0], o[:, 1], d[:, 0], d[:, 1]
        A = dx ** 2 + dy ** 2
        B = 2.0 * (ox * dx + oy * dy)
        Cq = (ox ** 2 + oy ** 2) - row.radius ** 2
        disc = B ** 2 - 4.0 * A * Cq
        ok = disc >= 0
        sq = _sqrt(torch.abs(disc))
        inf = torch.full_like(A, INF)
        return [torch.where(ok, (-B - sq) / (2.0 * A), inf), torch.where(ok, (-B + sq) / (2.0 * A), inf)]
    if row.surf == C.SURF_CONE:                   # geom/primitives.py:416-468 (slope in the c slot)
        ox, oy, oz, dx, dy, dz = o[:, 0], o[:, 1], o[:, 2], d[:, 0], d[:, 1], d[:, 2]
        k2 = row.c ** 2
        A = dz ** 2 - k2 * (dx ** 2 + dy ** 2)
        B = 2.0 * (oz * dz - k2 * (ox * dx + oy * dy))
        Cq = oz ** 2 - k2 * (ox ** 2 + oy ** 2)
        disc = B ** 2 - 4.0 * A * Cq
        ok = disc >= 0
        sq = _sqrt(torch.where(ok, disc, torch.zeros_like(disc)))
        lin = torch.abs(A) < EPS_S
        A_safe = torch.where(lin, torch.ones_like(A), A)
        t1 = (-B - sq) / (2.0 * A_safe)
        t2 = (-B + sq) / (2.0 * A_safe)
        B_safe = torch.where(torch.abs(B) < EPS_S, torch.full_like(B, EPS_S), B)
        t_lin = -Cq / B_safe
        inf = torch.full_like(A, INF)
        t1 = torch.where(lin, t_lin, torch.where(ok, t1, inf))
        t2 = torch.where(lin, t_lin, torch.where(ok, t2, inf))
        return [t1, t2]
    # conic sections: geom/primitives.py:266-320 and :356-376
    c, k = row.c, row.k
    oy, oz, dy, dz = o[:, 1], o[:, 2], d[:, 1], d[:, 2]
    if row.surf == C.SURF_QUADRIC:
        ox, dx = o[:, 0], d[:, 0]
        A = c * (dx ** 2 + dy ** 2) + c * (1 + k) * dz ** 2
        B = 2 * c * (ox * dx + oy * dy) + 2 * c * (1 + k) * oz * dz - 2 * dz
        Cq = c * (ox ** 2 + oy ** 2) + c * (1 + k) * oz ** 2 - 2 * oz
    else:
        A = c * dy ** 2 + c * (1 + k) * dz ** 2
        B = 2 * c * (oy * dy) + 2 * c * (1 + k) * oz * dz - 2 * dz
        Cq = c * oy ** 2 + c * (1 + k) * oz ** 2 - 2 * oz
    disc = B ** 2 - 4 * A * Cq
    ok = disc >= 0
    lin = torch.abs(A) < EPS_S
    sq = _sqrt(torch.abs(disc))
    A_safe = torch.where(lin, torch.ones_like(A), A)
    t1 = (-B - sq) / (2.0 * A_safe)
    t2 = (-B + sq) / (2.0 * A_safe)
    B_safe = torch.where(torch.abs(B) < EPS_S, torch.full_like(B, EPS_S), B)
    t_lin = -Cq / B_safe
    inf = torch.full_like(A, INF)
    t1 = torch.where(lin, t_lin, torch.where(ok, t1, inf))
    t2 = torch.where(lin, t_lin, torch.where(ok, t2, inf))
    return [t1, t2]


def _surface_in_bounds(row: Row, h):
    """Surface-level bound rule on local points [M,3] -> bool [M]."""
    sb = row.sb
    if row.bound == C.BOUND_DISK:                 # geom/bounded.py:60-64
        return h[:, 0] ** 2 + h[:, 1] ** 2 <= sb[0] ** 2
    if row.bound == C.BOUND_RECT:                 # geom/bounded.py:77-82
        return (torch.abs(h[:, 0]) <= sb[0]) & (torch.abs(h[:, 1]) <= sb[1])
    if row.bound == C.BOUND_ELLIPSE:              # geom/bounded.py:98-106
        u = h[:, 0] * sb[2] - h[:, 1] * sb[3]
        v = h[:, 0] * sb[3] + h[:, 1] * sb[2]
        return ((u / sb[0]) ** 2 + (v / sb[1]) ** 2) <= 1.0
    if row.bound in (C.BOUND_HALF, C.BOUND_HALF_DISK):   # geom/bounded.py:123-127, 171-174
        keep = torch.abs(h[:, 2] * row.c.detach()) < 1 + EPS_HALF
        if row.bound == C.BOUND_HALF_DISK:        # geom/bounded.py:151-159
            keep = keep & (h[:, 0] ** 2 + h[:, 1] ** 2 <= sb[0] ** 2)
        return keep
    if row.bound == C.BOUND_NAPPE:                # geom/bounded.py:208-217
        return (h[:, 2] * row.c.detach()) >= -1e-6
    return torch.ones(h.shape[0], dtype=torch.bool, device=h.device)


def _check_t(row: Row, roots, o, d):
    """Smallest admissible root; NaN propagates like torch.min."""
    t = torch.stack(roots)                        # [M,N]
    if row.bound == C.BOUND_NONE:                 # geom/primitives.py:28-36
        t = t.masked_fill(t <= EPS_T, INF)
    else:                                         # geom/bounded.py:20-36
        hits = o[None, :, :] + t[:, :, None] * d[None, :, :]
        M, N, _ = hits.shape
        keep = _surface_in_bounds(row, hits.view(-1, 3)).view(M, N)
        if row.invert:
            keep = ~keep
        t = t.masked_fill((t <= EPS_T) | ~keep, INF)
    return torch.min(t, dim=0)[0]


def _sag(c, h, tz):                               # geom/bounded.py:129-139, 176-186
    h2 = h ** 2
    return (c * h2) / (1.0 + _sqrt(torch.relu(1.0 - c ** 2 * h2))) + tz


def _shape_in_bounds(row: Row, h, rows: List[Row], r_idx: int):
    """Shape-level validity of element-frame points (only used by intersect tests)."""
    hb = row.hb
    x, y, z = h[:, 0], h[:, 1], h[:, 2]
    if row.shape == C.SHAPE_SPHERIC_FACE:         # geom/spherics.py:40-46
        return x ** 2 + y ** 2 <= hb[0] ** 2
    if row.shape == C.SHAPE_SPHERIC_EDGE:         # geom/spherics.py:34-39 (sags precomputed by the compiler)
        return (z >= hb[0]) & (z <= hb[1])
    if row.shape in (C.SHAPE_CYL_FACE, C.SHAPE_CYL_EDGE):     # geom/cylindrics.py:23-55
        ap = (x <= hb[1] + 1e-5) & (x >= hb[0] - 1e-5) & (y <= hb[3] + 1e-5) & (y >= hb[2] - 1e-5)
        if row.shape == C.SHAPE_CYL_FACE:
            return ap
        zf, zb = _sag(hb[4], y, hb[5]), _sag(hb[6], y, hb[7])
        return (z >= zf + 1e-4) & (z <= zb - 1e-4) & ap
    if row.shape == C.SHAPE_POLY:                 # geom/shape.py:122-132 (uses ROW 2 of each R)
        ok = torch.ones(h.shape[0], dtype=torch.bool, device=h.device)
        for m in range(row.poly_first, row.poly_first + row.poly_count):
            if m == r_idx:
                continue
            nrm, T = rows[m].Rs.detach()[2, :], rows[m].Ts.detach()
            ok = ok & (torch.sum(nrm[None, :] * (h - T[None, :]), dim=-1) < 1e-4)
        return ok
    raise ValueError("row has no shape-level rule")


def intersect_row(rows: List[Row], r_idx: int, p, d):
    """Distance to row r with all validity rules, inf/NaN = miss.

    Element with a Shape: geom/shape.py:25-59 (two poses, renormalised direction, shape
    validity).  Element whose shape is a bare Surface: geom/primitives.py:38-57."""
    row = rows[r_idx]
    if row.shape == C.SHAPE_NONE:
        o, dd = _pose(p, d, row.Rs, row.Ts)
        return _check_t(row, _roots(row, o, dd), o, dd)
    pe, de = _pose(p, d, row.Re, row.Te)
    de_n = F.normalize(de, p=2, dim=1)            # rays/ray.py:25 through with_coords
    o, dd = _pose(pe, de_n, row.Rs, row.Ts)
    t = _check_t(row, _roots(row, o, dd), o, dd)
    hit = pe + t[:, None] * de                    # un-normalised de: geom/shape.py:47
    valid = (t < INF) & _shape_in_bounds(row, hit, rows, r_idx)
    return torch.where(valid, t, torch.full_like(t, INF))


def _normal_local(row: Row, h):
    if row.surf == C.SURF_PLANE:                  # geom/primitives.py:138-143
        n = torch.zeros_like(h)
        n[:, 2] = 1.0
        return n
    if row.surf == C.SURF_SPHERE:                 # geom/primitives.py:186-187
        return h / row.radius
    if row.surf == C.SURF_CYLINDER:               # geom/primitives.py:233-241
        return torch.stack([h[:, 0] / row.radius, h[:, 1] / row.radius, torch.zeros_like(h[:, 0])], dim=1)
    if row.surf == C.SURF_CONE:                   # geom/primitives.py:470-494
        k2 = row.c ** 2
        raw = torch.stack([-k2 * h[:, 0], -k2 * h[:, 1], h[:, 2]], dim=1)
        ln = torch.norm(raw, dim=1, keepdim=True)
        up = torch.tensor([0.0, 0.0, 1.0], device=raw.device, dtype=raw.dtype)
        return torch.where(ln > 1e-8, raw / (ln + 1e-8), up)
    c, k = row.c, row.k                           # geom/primitives.py:330-343, 378-395
    nx = 2 * c * h[:, 0] if row.surf == C.SURF_QUADRIC else torch.zeros_like(h[:, 0])
    ny = 2 * c * h[:, 1]
    nz = 2 * c * (1 + k) * h[:, 2] - 2.0
    raw = torch.stack([nx, ny, nz], dim=1)
    return -(raw / (torch.norm(raw, dim=1, keepdim=True) + 1e-8))


def geometry_row(rows: List[Row], r_idx: int, p, d):
    """(t, hit_global, normal_global, hit_local) — geom/shape.py:61-87 or geom/primitives.py:59-96."""
    row = rows[r_idx]
    if row.shape == C.SHAPE_NONE:
        o, dd = _pose(p, d, row.Rs, row.Ts)
        t = _check_t(row, _roots(row, o, dd), o, dd)
        hit_local = o + t[:, None] * dd
        n_glob = _normal_local(row, hit_local) @ row.Rs.T
    else:
        pe, de = _pose(p, d, row.Re, row.Te)
        de_n = F.normalize(de, p=2, dim=1)
        o, dd = _pose(pe, de_n, row.Rs, row.Ts)
        t = _check_t(row, _roots(row, o, dd), o, dd)
        hit_local = o + t[:, None] * dd
        n_glob = (_normal_local(row, hit_local) @ row.Rs.T) @ row.Re.T
    return t, p + t[:, None] * d, n_glob, hit_local


def physics_row(row: Row, hit_local, d, n, ior=None, u=None):
    """(new_dir, intensity_mod).  ``ior`` optionally overrides (ior_in, ior_out) per ray; ``u`` = the uniform
    draws of a Fresnel row (the reference calls torch.rand_like, phys/std.py:190)."""
    ones = torch.ones_like(d[:, 0])
    if row.phys == C.PHYS_FRESNEL:                # phys/std.py:175-224
        ior_in, ior_out = (row.ior_in, row.ior_out) if ior is None else (ior[0][:, None], ior[1][:, None])
        dot = torch.sum(d * n, dim=1, keepdim=True)
        entering = dot < 0
        n_eff = torch.where(entering, n, -n)
        cos_i = torch.abs(dot)
        n1 = torch.where(entering, ior_in, ior_out)
        n2 = torch.where(entering, ior_out, ior_in)
        mu = n2 / n1
        sin2_t = (mu ** 2) * (1.0 - cos_i ** 2)
        is_tir = sin2_t > 1.0
        cos_t = _sqrt(torch.relu(1.0 - sin2_t))
        n1_ci, n2_ct, n1_ct, n2_ci = n1 * cos_i, n2 * cos_t, n1 * cos_t, n2 * cos_i
        rs = ((n1_ci - n2_ct) / (n1_ci + n2_ct + 1e-8)) ** 2
        rp = ((n1_ct - n2_ci) / (n1_ct + n2_ci + 1e-8)) ** 2
        R = torch.where(is_tir, torch.ones_like(rs), 0.5 * (rs + rp))
        reflect = u.to(R.dtype)[:, None] < R
        v_reflect = d - 2 * dot * n
        v_refract = mu * d + (mu * cos_i - cos_t) * n_eff
        return torch.where(reflect, v_reflect, v_refract), ones
    if row.phys == C.PHYS_TRANSMIT:               # phys/std.py:227-235
        return d, ones
    if row.phys == C.PHYS_BLOCK:                  # phys/std.py:243-254
        return torch.zeros_like(d), torch.zeros_like(d[:, 0])
    if row.phys == C.PHYS_REFLECT:                # phys/std.py:97-108
        cos = torch.sum(d * n, dim=1, keepdim=True)
        return d - 2 * cos * n, ones
    if row.phys == C.PHYS_APERTURE:               # phys/filter.py:24-33 (non-inverted bound)
        m = _surface_in_bounds(row, hit_local).to(d.dtype)
        return d * m[:, None], m
    if row.phys == C.PHYS_LINEAR:                 # phys/std.py:72-88 (Linear.transform = the plane's own pose)
        Cx, Cy, Dx, Dy = row.c, row.k, row.radius, row.ior_in
        dl = d @ row.Rs
        dl = dl / dl[:, 2][:, None]
        nx = Cx * hit_local[:, 0] + Dx * dl[:, 0]
        ny = Cy * hit_local[:, 1] + Dy * dl[:, 1]
        new_local = F.normalize(torch.stack([nx, ny, torch.ones_like(nx)], dim=1), p=2, dim=1)
        return new_local @ row.Rs.T, ones
    # Snell with TIR fallback: phys/std.py:123-145
    n_in, n_out = (row.ior_in, row.ior_out) if ior is None else (ior[0][:, None], ior[1][:, None])
    dot = torch.sum(d * n, dim=1, keepdim=True)
    entering = dot < 0
    n_eff = torch.where(entering, n, -n)
    c1 = torch.abs(dot)
    mu = torch.where(entering, n_out / n_in, n_in / n_out)
    term = 1.0 - mu ** 2 * (1.0 - c1 ** 2)
    c2 = _sqrt(torch.relu(term))
    refr = mu * d + (mu * c1 - c2) * n_eff
    refl = d - 2 * dot * n
    return torch.where(term < 0, refl, refr), ones


def make_rows(table_f: torch.Tensor, table_i) -> List[Row]:
    meta = table_i.tolist() if isinstance(table_i, torch.Tensor) else table_i
    return [Row(table_f[r], meta[r]) for r in range(table_f.shape[0])]


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Philox4x32-10 on numpy uint32 arrays (Salmon et al. 2011; the rounds of rtt_core.cuh::philox4x32_10)."""
    import numpy as np
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint64) for c in (c0, c1, c2, c3))
    k0, k1 = np.uint64(k0), np.uint64(k1)
    M0, M1, W0, W1, LO = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57), np.uint64(0x9E3779B9), np.uint64(0xBB67AE85), \
        np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0, p1 = M0 * c0, M1 * c2
        n0 = ((p1 >> np.uint64(32)) ^ c1 ^ k0) & LO
        n1 = p1 & LO
        n2 = ((p0 >> np.uint64(32)) ^ c3 ^ k1) & LO
        n3 = p0 & LO
        c0, c1, c2, c3 = n0, n1, n2, n3
        k0, k1 = (k0 + W0) & LO, (k1 + W1) & LO
    return c0, c1, c2, c3


def fresnel_u(seed: int, ray_index, row: int, bounce: int):
    """The uniform [0,1) draw of (ray, row, bounce) under ``seed`` (include/rtt_b200.h, RTT_PHYS_FRESNEL)."""
    import numpy as np
    i = np.asarray(ray_index, dtype=np.uint64)
    out = philox4x32_10(i & np.uint64(0xFFFFFFFF), i >> np.uint64(32), np.full_like(i, row + 256 * bounce),
                        np.full_like(i, 0x4672), seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)[0]
    return torch.from_numpy(((out >> np.uint64(8)).astype(np.float32) * np.float32(1.0 / 16777216.0)))


def table_seed(table_i) -> int:
    meta = table_i.tolist() if isinstance(table_i, torch.Tensor) else table_i
    return ((meta[0][C.I_RNG_HI] & 0xFFFFFFFF) << 32) | (meta[0][C.I_RNG_LO] & 0xFFFFFFFF)


def _lut_index(wavelength, lut_w):
    return torch.argmin(torch.abs(wavelength[:, None] - lut_w[None, :]), dim=1)


def element_step(rows: List[Row], r_idx: int, p, d, ior=None, u=None):
    """One Element.forward on rays assumed to hit (elements/parent.py:44-58)."""
    t, hit, n, hit_local = geometry_row(rows, r_idx, p, d)
    new_d, mod = physics_row(rows[r_idx], hit_local, d, n, ior, u)
    return hit, new_d, mod, hit_local, t, n


def _row_extras(rows, r, mask, lam, lut, seed, bounce):
    """(ior override, Fresnel draws) of the rays selected by ``mask`` at row r."""
    ior = u = None
    if lam is not None and rows[r].phys in (C.PHYS_SNELL, C.PHYS_FRESNEL):
        sel = lut[lam[mask], r]                    # [M,2]
        ior = (sel[:, 0], sel[:, 1])
    if rows[r].phys == C.PHYS_FRESNEL:
        u = fresnel_u(seed, torch.nonzero(mask)[:, 0].cpu().numpy(), r, bounce).to(mask.device)
    return ior, u


def trace_sequential(table_f, table_i, pos, dir_, intensity, *, wavelength=None, lut=None, lut_w=None):
    """scene/sequential.py:12-36 over the flat table.

    Returns dict(pos, dir, intensity, hit [N,S] bool, sensor={slot: (mask, hit_local, w)})."""
    rows = make_rows(table_f, table_i)
    seed = table_seed(table_i)
    N, S = pos.shape[0], len(rows)
    hit_log = torch.zeros(N, S, dtype=torch.bool, device=pos.device)
    sensor: Dict[int, tuple] = {}
    lam = _lut_index(wavelength, lut_w) if lut is not None else None
    for r in range(S):
        with torch.no_grad():
            t = intersect_row(rows, r, pos, dir_)
        mask = t < INF                             # NaN -> miss (Cylinder)
        hit_log[:, r] = mask
        if not bool(mask.any()):
            continue
        ior, u = _row_extras(rows, r, mask, lam, lut, seed, 0)
        hit, new_d, mod, hit_local, _, _ = element_step(rows, r, pos[mask], dir_[mask], ior, u)
        if rows[r].sensor >= 0:                    # elements/sensor.py:35-37: intensity BEFORE update
            sensor[rows[r].sensor] = (mask, hit_local, intensity[mask])
        idx = (mask,)
        pos = pos.index_put(idx, hit)
        dir_ = dir_.index_put(idx, new_d)
        intensity = intensity.index_put(idx, intensity[mask] * mod)
    return dict(pos=pos, dir=dir_, intensity=intensity, hit=hit_log, sensor=sensor)


def ray_cast(rows: List[Row], pos, dir_):
    """scene/base.py:144-178: (hit_mask, winner_row); NaN in any column -> no hit."""
    with torch.no_grad():
        tm = torch.stack([intersect_row(rows, r, pos, dir_) for r in range(len(rows))], dim=1)
        tmin, win = torch.min(tm, dim=1)
        return tmin < INF, win, tmin


def trace_nonsequential(table_f, table_i, pos, dir_, intensity, nbounces: int, *,
                        wavelength=None, lut=None, lut_w=None):
    """scene/base.py:129-235 over the flat table.

    Returns dict(pos, dir, intensity, seq [N,B] int (row per bounce, -1 none), nb [N],
    sensor_hits=[(ray_index, slot, hit_local, w)] in recording order)."""
    rows = make_rows(table_f, table_i)
    seed = table_seed(table_i)
    N = pos.shape[0]
    seq = torch.full((N, nbounces), -1, dtype=torch.long, device=pos.device)
    sensor_hits = []
    lam = _lut_index(wavelength, lut_w) if lut is not None else None
    for b in range(nbounces):
        if not bool((intensity > 0).any()):
            break
        hit_mask, win, _ = ray_cast(rows, pos, dir_)
        if not bool(hit_mask.any()):
            break
        active = hit_mask & (intensity > 0)
        if not bool(active.any()):
            break
        new_p, new_d, new_i = pos, dir_, intensity
        for r in range(len(rows)):
            m = active & (win == r)
            if not bool(m.any()):
                continue
            ior, u = _row_extras(rows, r, m, lam, lut, seed, b)
            hit, nd, mod, hit_local, _, _ = element_step(rows, r, pos[m], dir_[m], ior, u)
            if rows[r].sensor >= 0:
                sensor_hits.append((torch.nonzero(m)[:, 0], rows[r].sensor, hit_local, intensity[m]))
            new_p = new_p.index_put((m,), hit)
            new_d = new_d.index_put((m,), nd)
            new_i = new_i.index_put((m,), intensity[m] * mod)
            seq[m, b] = r
        pos, dir_, intensity = new_p, new_d, new_i
    return dict(pos=pos, dir=dir_, intensity=intensity, seq=seq, nb=(seq >= 0).sum(1), sensor_hits=sensor_hits)


def sensor_bins(hit_local, spec):
    """Bin indices of the fixed-range histogram (gui/workbench.py:615-624).

    spec = (H, W, x0, x1, y0, y1, channels).  Returns (iy, ix, inside) computed in the
    dtype of ``hit_local`` with ``floor((v - v0) / (v1 - v0) * n)``."""
    H, W, x0, x1, y0, y1 = spec[:6]
    x, y = hit_local[:, 0], hit_local[:, 1]
    fx = torch.floor((x - x0) / (x1 - x0) * W)
    fy = torch.floor((y - y0) / (y1 - y0) * H)
    inside = (fx >= 0) & (fx < W) & (fy >= 0) & (fy < H)
    return fy.long(), fx.long(), inside


def sensor_image(hit_local, w, spec, channel=None):
    """[C,H,W] float64 image (accumulated in double so the oracle is order independent)."""
    H, W = spec[0], spec[1]
    Cn = spec[6] if len(spec) > 6 else 1
    iy, ix, inside = sensor_bins(hit_local, spec)
    ch = torch.zeros_like(iy) if channel is None else channel.long()
    img = torch.zeros(Cn, H, W, dtype=torch.float64, device=hit_local.device)
    img.index_put_((ch[inside], iy[inside], ix[inside]), w[inside].double(), accumulate=True)
    return img
