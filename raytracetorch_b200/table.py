"""Scene compiler: Element/Shape/Surface objects -> flat surface table.

The kernels never see Python objects.  ``compile_elements`` walks the elements in the
reference's flattening order (``scene/base.py:116-123`` == the double loop of
``scene/sequential.py:17-19``) and emits one row per ``(element, surface index)``:

* ``SurfaceTable.f``  ``[S, ROW_F] float``  — poses as rotation matrices + translations,
  surface scalars, refractive indices, bound parameters.  Built with batched,
  differentiable torch ops from the live ``nn.Parameter`` objects, so parameter sharing
  (``geom/spherics.py:92-93``, ``elements/lens.py:41-56``), gradient-mask hooks
  (``geom/transform.py:29-35``) and ``matrix_exp`` (``geom/transform.py:58``) are handled
  by autograd *outside* the kernels; the adjoint kernel only returns ``d loss / d f``.
* ``SurfaceTable.i``  ``[S, ROW_I] int32`` — kinds, flags, sensor slots, sibling ranges.
* ``SurfaceTable.lut`` ``[L, S, 2] float`` — optional per-wavelength (ior_in, ior_out).

The compiler is duck-typed on class *names* (``type(obj).__mro__``), so it accepts this
package's description objects and the reference's own objects alike; the latter is how
``oracle/make_golden.py`` ties the table + oracle to the unmodified reference.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence

import torch

from . import codes as C
from .geom import rotation_from_vector


class UnsupportedSceneError(NotImplementedError):
    """Raised for objects outside the fused path (custom callables such as ``Fuzzy``, unfinished reference classes)."""


def _names(obj) -> set:
    return {k.__name__ for k in type(obj).__mro__}


def _surface_kind(surf) -> int:
    n = _names(surf)
    if "WedgeYZ" in n:
        raise UnsupportedSceneError("WedgeYZ is unfinished in the reference (geom/primitives.py:497-502)")
    if "Cone" in n:
        return C.SURF_CONE
    if "Sphere" in n:
        return C.SURF_SPHERE
    if "Cylinder" in n:
        return C.SURF_CYLINDER
    if "QuadricZY" in n:
        return C.SURF_QUADRIC_ZY
    if "Quadric" in n:
        return C.SURF_QUADRIC
    if "Plane" in n:
        return C.SURF_PLANE
    raise UnsupportedSceneError(f"unknown surface class {type(surf).__name__}")


def _bound_kind(surf) -> int:
    n = _names(surf)
    if "BoundedHalfSphere" in n:
        return C.BOUND_HALF_DISK
    if "HalfSphere" in n or "HalfCyl" in n:
        return C.BOUND_HALF
    if "SingleCone" in n:
        return C.BOUND_NAPPE
    if "Disk" in n:
        return C.BOUND_DISK
    if "Rectangle" in n:
        return C.BOUND_RECT
    if "Ellipse" in n:
        return C.BOUND_ELLIPSE
    if "SurfaceBounded" in n:
        raise UnsupportedSceneError(f"bounded surface {type(surf).__name__} has no kernel rule")
    return C.BOUND_NONE


def _phys_kind(sf) -> int:
    n = _names(sf)
    if "ApertureFilter" in n:
        return C.PHYS_APERTURE
    if "RefractFresnel" in n:
        return C.PHYS_FRESNEL
    if "RefractSnell" in n:
        return C.PHYS_SNELL
    if "Reflect" in n:
        return C.PHYS_REFLECT
    if "Block" in n:
        return C.PHYS_BLOCK
    if "Linear" in n:
        return C.PHYS_LINEAR
    if "Fuzzy" in n:
        raise UnsupportedSceneError(f"surface function {type(sf).__name__} is outside the fused path")
    if "Transmit" in n:
        return C.PHYS_TRANSMIT
    raise UnsupportedSceneError(f"unknown surface function {type(sf).__name__}")


class Ref:
    """Late-bound read of a live tensor: ``fn(tensor[index])`` evaluated at table-build
    time, so cached plans never hold stale views of a Parameter."""

    __slots__ = ("tensor", "index", "fn")

    def __init__(self, tensor, index=None, fn=None):
        self.tensor, self.index, self.fn = tensor, index, fn

    def get(self):
        t = self.tensor if self.index is None else self.tensor[self.index]
        return t if self.fn is None else self.fn(t)


@dataclass
class RowPlan:
    """Static description of one table row plus the live tensors it reads."""
    elem: int
    sidx: int
    meta: List[int]
    rot_e: Optional[torch.Tensor]
    trans_e: Optional[torch.Tensor]
    rot_s: torch.Tensor
    trans_s: torch.Tensor
    scal: List[Optional[torch.Tensor]]          # c, k, radius, ior_in, ior_out
    sb: List[Optional[torch.Tensor]]            # 4 surface-bound entries
    hb: List[Optional[torch.Tensor]]            # 8 shape-bound entries
    sag: Optional[tuple] = None                 # (c_lo, tz_lo, c_hi, tz_hi, h) for spheric edges


@dataclass
class SurfaceTable:
    f: torch.Tensor                 # [S, ROW_F]
    i: torch.Tensor                 # [S, ROW_I] int32 (same device as f)
    i_host: List[List[int]]
    lut: Optional[torch.Tensor] = None          # [L, S, 2]
    lut_wavelengths: Optional[torch.Tensor] = None  # [L]
    sensors: list = field(default_factory=list)     # Sensor elements by slot
    sensor_rows: List[int] = field(default_factory=list)
    elements: list = field(default_factory=list)

    @property
    def n_rows(self) -> int:
        return self.f.shape[0]

    @property
    def stochastic(self) -> bool:
        """True iff some row draws random numbers (RefractFresnel)."""
        return any(m[C.I_PHYS] == C.PHYS_FRESNEL for m in self.i_host)

    def with_seed(self, seed: int) -> "SurfaceTable":
        """Copy of this table whose Fresnel draws are keyed by ``seed`` (ints [I_RNG_LO, I_RNG_HI] of row 0).  The seed
        travels inside ``table.i`` through the forward AND the adjoint op, so both take the same branches."""
        import dataclasses
        seed &= (1 << 64) - 1
        lo, hi = seed & 0xFFFFFFFF, seed >> 32
        to_i32 = lambda v: v - (1 << 32) if v >= (1 << 31) else v
        meta = [list(m) for m in self.i_host]
        meta[0][C.I_RNG_LO], meta[0][C.I_RNG_HI] = to_i32(lo), to_i32(hi)
        i = self.i.clone()
        i[0, C.I_RNG_LO], i[0, C.I_RNG_HI] = to_i32(lo), to_i32(hi)
        return dataclasses.replace(self, i=i, i_host=meta)


class Dispersion:
    """Per-wavelength refractive indices (extension; the reference never reads
    ``Rays.wavelength`` — SURVEY.md section 0.3).

    ``wavelengths``: the L sample wavelengths, in the same unit as ``Rays.wavelength``.
    ``glasses``: ``{ior Parameter: [L] values}`` — any ``ior_in``/``ior_out`` Parameter found
    in this mapping is replaced, per ray, by the value of the nearest sample wavelength.
    Parity: tracing wavelength l must equal the reference with the scalar set to values[l].
    """

    def __init__(self, wavelengths: Sequence[float], glasses: Dict[torch.nn.Parameter, Sequence[float]]):
        self.wavelengths = torch.as_tensor(wavelengths, dtype=torch.float32)
        if self.wavelengths.numel() > C.MAX_WAVELENGTHS:
            raise ValueError(f"at most {C.MAX_WAVELENGTHS} sample wavelengths")
        self.values = {}
        for p, v in glasses.items():
            v = v if isinstance(v, torch.Tensor) else torch.as_tensor(v, dtype=torch.float32)
            if v.shape != self.wavelengths.shape:
                raise ValueError("one index value per sample wavelength required")
            self.values[id(p)] = v


def plan_elements(elements) -> tuple:
    """Walk the elements once; returns (rows, sensors, sensor_rows)."""
    rows: List[RowPlan] = []
    sensors, sensor_rows = [], []
    for e_idx, el in enumerate(elements):
        shape = el.shape
        sn = _names(shape)
        is_shape = "Shape" in sn
        members = list(shape.surfaces) if is_shape else [shape]
        if len(el.surface_functions) < len(members):
            raise UnsupportedSceneError(f"element {e_idx}: fewer surface functions than surfaces")
        first_row = len(rows)
        n_opt = int(getattr(shape, "N_optical", 0))
        is_sensor = "Sensor" in _names(el)
        for s_idx, surf in enumerate(members):
            sf = el.surface_functions[s_idx]
            kind, bound, phys = _surface_kind(surf), _bound_kind(surf), _phys_kind(sf)
            meta = [0] * C.ROW_I
            meta[C.I_SURF], meta[C.I_BOUND] = kind, bound
            meta[C.I_INVERT] = int(bool(getattr(surf, "invert", False)))
            meta[C.I_PHYS] = phys
            meta[C.I_SENSOR] = -1
            meta[C.I_ELEM], meta[C.I_SIDX] = e_idx, s_idx
            if phys == C.PHYS_APERTURE:
                owner = getattr(getattr(sf, "_inBounds", None), "__self__", None)
                if owner is not surf:
                    raise UnsupportedSceneError("ApertureFilter must filter on its own element's surface")
            scal = [getattr(surf, "slope", None) if kind == C.SURF_CONE else getattr(surf, "c", None),
                    getattr(surf, "k", None),
                    surf.radius if kind in (C.SURF_SPHERE, C.SURF_CYLINDER) else None,
                    getattr(sf, "ior_in", None) if phys in (C.PHYS_SNELL, C.PHYS_FRESNEL) else None,
                    getattr(sf, "ior_out", None) if phys in (C.PHYS_SNELL, C.PHYS_FRESNEL) else None]
            if phys == C.PHYS_LINEAR:
                # ray-transfer coefficients ride in the scalar slots a plane does not use (codes.py)
                if kind != C.SURF_PLANE or is_shape:
                    raise UnsupportedSceneError("Linear physics is defined on a bare plane (elements/ideal.py:47-54)")
                if getattr(sf, "transform", None) is not surf.transform:
                    raise UnsupportedSceneError("Linear.transform must be the plane's own transform (LinearElement)")
                scal = [sf.Cx, sf.Cy, sf.Dx, sf.Dy, None]
            sb: List[Optional[torch.Tensor]] = [None] * 4
            if bound == C.BOUND_DISK:
                sb[0] = surf.radius
            elif bound == C.BOUND_RECT:
                sb[0], sb[1] = surf.hx, surf.hy
            elif bound == C.BOUND_ELLIPSE:
                sb[0], sb[1] = surf.r_major, surf.r_minor
                sb[2], sb[3] = Ref(surf.rot, fn=torch.cos), Ref(surf.rot, fn=torch.sin)
            elif bound == C.BOUND_HALF_DISK:
                sb[0] = Ref(surf.diameter, fn=lambda d: d / 2.0)
            hb: List[Optional[torch.Tensor]] = [None] * 8
            sag = None
            if not is_shape:
                meta[C.I_SHAPE] = C.SHAPE_NONE
            elif "Spheric" in sn:
                if s_idx < n_opt:
                    meta[C.I_SHAPE] = C.SHAPE_SPHERIC_FACE
                    hb[0] = shape.radius
                else:
                    meta[C.I_SHAPE] = C.SHAPE_SPHERIC_EDGE
                    lo, hi = members[s_idx - n_opt], members[s_idx - n_opt + 1]
                    sag = (lo.c, Ref(lo.transform.trans, 2), hi.c, Ref(hi.transform.trans, 2), shape.radius)
            elif "Cylindric" in sn:
                meta[C.I_SHAPE] = C.SHAPE_CYL_FACE if s_idx < n_opt else C.SHAPE_CYL_EDGE
                hb[0] = Ref(members[3].transform.trans, 0)      # x_min   geom/cylindrics.py:31-34
                hb[1] = Ref(members[2].transform.trans, 0)      # x_max
                hb[2] = Ref(members[5].transform.trans, 1)      # y_min
                hb[3] = Ref(members[4].transform.trans, 1)      # y_max
                hb[4], hb[5] = members[0].c, Ref(members[0].transform.trans, 2)
                hb[6], hb[7] = members[1].c, Ref(members[1].transform.trans, 2)
            elif "CvxPolyhedron" in sn:
                meta[C.I_SHAPE] = C.SHAPE_POLY
                meta[C.I_POLY_FIRST], meta[C.I_POLY_COUNT] = first_row, len(members)
            else:
                raise UnsupportedSceneError(
                    f"shape class {type(shape).__name__} defines no in-bounds rule (geom/shape.py:89-94)")
            if is_sensor:
                if len(sensors) >= C.MAX_SENSORS and el not in sensors:
                    raise UnsupportedSceneError(f"at most {C.MAX_SENSORS} sensor rows")
                meta[C.I_SENSOR] = len(sensor_rows)
                sensors.append(el)
                sensor_rows.append(len(rows))
            tr_e = shape.transform if is_shape else None
            rows.append(RowPlan(e_idx, s_idx, meta,
                                tr_e.rot_vec if tr_e is not None else None,
                                tr_e.trans if tr_e is not None else None,
                                surf.transform.rot_vec, surf.transform.trans,
                                scal, sb, hb, sag))
    if len(rows) > C.MAX_ROWS:
        raise UnsupportedSceneError(f"{len(rows)} surface rows > MAX_ROWS={C.MAX_ROWS}")
    if len(sensor_rows) > C.MAX_SENSORS:
        raise UnsupportedSceneError(f"at most {C.MAX_SENSORS} sensor rows")
    return rows, sensors, sensor_rows


_ROT_CACHE: Dict[int, tuple] = {}


def _rotation_of(v: Optional[torch.Tensor], dtype, dev, eye) -> torch.Tensor:
    """R = expm(skew(rot_vec)) for ONE pose, unbatched on purpose.

    ``torch.linalg.matrix_exp`` picks its approximation degree per call; batching several
    poses changes the result by up to ~2e-6, while the reference evaluates each pose on its
    own (geom/transform.py:48-61).  Non-trainable rotations are cached per tensor version,
    like the reference's ``_cached_rot``; trainable ones are recomputed so autograd sees them.
    """
    if v is None:
        return eye
    if not v.requires_grad:
        hit = _ROT_CACHE.get(id(v))
        if hit is not None and hit[0]() is v and hit[1] == (v._version, v.device, dtype, dev):
            return hit[2]
    vv = v if (v.dtype == dtype and v.device == dev) else v.to(device=dev, dtype=dtype)
    R = rotation_from_vector(vv)
    if not v.requires_grad:
        import weakref
        _ROT_CACHE[id(v)] = (weakref.ref(v), (v._version, v.device, dtype, dev), R.detach())
    return R


def _rg(t) -> bool:
    return isinstance(t, torch.Tensor) and t.requires_grad


def _first_device(rows: List[RowPlan]) -> torch.device:
    return rows[0].trans_s.device if rows else torch.device("cpu")


_INT_TABLES: Dict[tuple, torch.Tensor] = {}


def _int_table(meta, S: int, dev) -> torch.Tensor:
    """Device copy of the int block, cached by content: it depends on the scene's structure and requires_grad flags
    only, so optimisation loops (and CUDA-graph captures, where a pageable host-to-device copy is illegal) reuse it."""
    key = (tuple(tuple(m) for m in meta), str(dev))
    hit = _INT_TABLES.get(key)
    if hit is None:
        if len(_INT_TABLES) > 256:
            _INT_TABLES.clear()
        hit = torch.tensor(meta, dtype=torch.int32).view(S, C.ROW_I).to(dev)
        _INT_TABLES[key] = hit
    return hit


def build_table(rows: List[RowPlan], sensors, sensor_rows, elements, *, dtype=torch.float32,
                device: Optional[torch.device] = None,
                dispersion: Optional[Dispersion] = None) -> SurfaceTable:
    """Assemble the float/int tables from a plan with a fixed, small number of torch ops."""
    dev = _first_device(rows) if device is None else torch.device(device)
    S = len(rows)
    zero = torch.zeros((), dtype=dtype, device=dev)
    zero3 = torch.zeros(3, dtype=dtype, device=dev)

    def cv(t, z):
        if t is None:
            return z
        if isinstance(t, Ref):
            t = t.get()
        if not isinstance(t, torch.Tensor):
            t = torch.as_tensor(t)
        if t.dtype != dtype or t.device != dev:
            t = t.to(device=dev, dtype=dtype)
        return t

    eye = torch.eye(3, dtype=dtype, device=dev)
    vecs = torch.stack([cv(x, zero3) for r in rows for x in (r.trans_e, r.trans_s)]).view(S, 2, 3)
    Re = torch.stack([_rotation_of(r.rot_e, dtype, dev, eye) for r in rows])
    Rs = torch.stack([_rotation_of(r.rot_s, dtype, dev, eye) for r in rows])
    scal = torch.stack([cv(x, zero) for r in rows for x in r.scal]).view(S, 5)
    with torch.no_grad():
        sb = torch.stack([cv(x, zero) for r in rows for x in r.sb]).view(S, 4)
        hb = torch.stack([cv(x, zero) for r in rows for x in r.hb]).view(S, 8)
        edge_rows = [n for n, r in enumerate(rows) if r.sag is not None]
        if edge_rows:
            # rim sag of the two neighbouring faces: geom/bounded.py:129-139, geom/spherics.py:34-39
            sg = torch.stack([cv(x, zero) for n in edge_rows for x in rows[n].sag]).view(-1, 5)
            h2 = sg[:, 4] ** 2

            def sag(c, tz):
                return (c * h2) / (1.0 + torch.sqrt(torch.relu(1.0 - c ** 2 * h2))) + tz

            zz = torch.stack([sag(sg[:, 0], sg[:, 1]), sag(sg[:, 2], sg[:, 3])], dim=1)
            hb = hb.clone()
            for k_edge, n_edge in enumerate(edge_rows):           # int indices: no host-to-device index tensor
                hb[n_edge, 0:2] = zz[k_edge]                      # (a pageable H2D copy is illegal in graph capture)
        pad = torch.zeros(S, C.ROW_F - C.F_HB - 8, dtype=dtype, device=dev)
    f = torch.cat([Re.reshape(S, 9), vecs[:, 0], Rs.reshape(S, 9), vecs[:, 1], scal, sb, hb, pad], dim=1)

    meta = []
    for r in rows:
        m = list(r.meta)
        fl = 0
        if _rg(r.rot_e) or _rg(r.trans_e):
            fl |= C.FLAG_GRAD_POSE_E
        if _rg(r.rot_s) or _rg(r.trans_s):
            fl |= C.FLAG_GRAD_POSE_S
        if _rg(r.scal[0]) or _rg(r.scal[1]):
            fl |= C.FLAG_GRAD_CK
        if _rg(r.scal[2]):
            fl |= C.FLAG_GRAD_RADIUS
        if _rg(r.scal[3]) or _rg(r.scal[4]):
            fl |= C.FLAG_GRAD_IOR
        m[C.I_FLAGS] = fl
        meta.append(m)
    i = _int_table(meta, S, dev)

    lut = lut_w = None
    if dispersion is not None:
        L = dispersion.wavelengths.numel()
        cols = []
        for r in rows:
            for p in (r.scal[3], r.scal[4]):
                v = dispersion.values.get(id(p)) if p is not None else None
                cols.append(cv(v, zero).expand(L) if v is not None else cv(p, zero).expand(L))
        lut = torch.stack(cols).view(S, 2, L).permute(2, 0, 1).contiguous()
        lut_w = dispersion.wavelengths.to(device=dev, dtype=dtype)
    return SurfaceTable(f=f, i=i, i_host=meta, lut=lut, lut_wavelengths=lut_w,
                        sensors=list(sensors), sensor_rows=list(sensor_rows), elements=list(elements))


def compile_elements(elements, *, dtype=torch.float32, device=None,
                     dispersion: Optional[Dispersion] = None) -> SurfaceTable:
    """One-shot: plan + build."""
    elements = list(elements)
    rows, sensors, sensor_rows = plan_elements(elements)
    return build_table(rows, sensors, sensor_rows, elements, dtype=dtype, device=device, dispersion=dispersion)


class SceneCompiler:
    """Caches the static plan of a scene; rebuilds the float table when needed.

    The plan is invalidated when the element list changes.  The float table itself is
    reused only if no involved tensor requires grad and none was modified in place
    (tensor ``_version`` counters), so optimiser steps are always seen."""

    def __init__(self):
        self._key = None
        self._plan = None
        self._table = None
        self._versions = None

    def _leaves(self, rows):
        for r in rows:
            for t in (r.rot_e, r.trans_e, r.rot_s, r.trans_s, *r.scal, *r.sb, *r.hb, *(r.sag or ())):
                if isinstance(t, Ref):
                    t = t.tensor
                if isinstance(t, torch.Tensor):
                    yield t

    def table(self, elements, *, dispersion=None, dtype=torch.float32) -> SurfaceTable:
        elements = list(elements)
        key = (tuple(id(e) for e in elements), tuple(len(e.shape) for e in elements), id(dispersion), dtype)
        if key != self._key:
            self._plan = plan_elements(elements)
            self._key, self._table = key, None
        rows, sensors, sensor_rows = self._plan
        leaves = list(self._leaves(rows))
        needs_grad = torch.is_grad_enabled() and any(t.requires_grad for t in leaves)
        versions = tuple((id(t), t._version, t.device) for t in leaves)
        if self._table is not None and not needs_grad and versions == self._versions \
                and not self._table.f.requires_grad:
            return self._table
        tab = build_table(rows, sensors, sensor_rows, elements, dtype=dtype, dispersion=dispersion)
        self._table, self._versions = tab, versions
        return tab
