"""Optical elements: a shape plus one surface function per member surface.

Mirror of the reference's ``elements`` package for the hot path
(``elements/parent.py``, ``lens.py``, ``mirror.py``, ``aperture.py``, ``sensor.py``).
Constructor arguments, attribute names and parameter sharing match the reference so that
scripts written against it keep working; the per-ray work is done by the CUDA ops.

Parity notes carried over verbatim from the reference:
* ``SingletLens`` binds ``ior_in=glass, ior_out=media`` on the FRONT face and the
  opposite on the back face (``elements/lens.py:41-49``) — the reverse of
  ``DoubletLens``/``TripletLens`` (``:261-276``).  Reproduced as is.
* ``CylSingletLens`` re-uses the third surface function for the 4 side planes
  (``elements/lens.py:206-208``).
* ``Sensor`` records the intensity *before* this surface's modulation
  (``elements/sensor.py:35-37``).
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn

from . import geom as G
from . import phys as P


class Element(nn.Module):
    """shape + surface_functions (elements/parent.py:8-58)."""

    def __init__(self):
        super().__init__()
        self.shape = G.Shape()
        self.surface_functions = nn.ModuleList()

    def intersectTest(self, rays):
        """[N,K] distances, inf = miss (elements/parent.py:30-42)."""
        return self.shape.intersectTest(rays)

    def forward(self, rays, surf_idx):
        """(new_pos, new_dir, intensity_mult) for rays hitting member ``surf_idx``
        (elements/parent.py:44-58).  One fused CUDA op with a hand-written adjoint."""
        from .ops import element_step
        new_pos, new_dir, mod, _hit_local = element_step(self, rays, int(surf_idx))
        return new_pos, new_dir, mod


class ElementCustom(Element):
    def __init__(self, shape, surface_function, device=None):
        super().__init__()
        self.shape = shape
        self.surface_functions.extend(len(shape) * [surface_function])


def _snell(ior_in: nn.Parameter, ior_out: nn.Parameter, fresnel: bool = False) -> P.RefractSnell:
    sf = (P.RefractFresnel if fresnel else P.RefractSnell)(0.0, 0.0)     # elements/lens.py:35-38
    sf.ior_in, sf.ior_out = ior_in, ior_out      # shared Parameters, as in elements/lens.py:41-47
    return sf


def _scalar_param(v, grad):
    return nn.Parameter(torch.as_tensor(float(v)), requires_grad=grad)


class SingletLens(Element):
    def __init__(self, c1: float, c2: float, d: float, t: float,
                 ior_glass: float, ior_media: float = 1.0,
                 c1_grad: bool = False, c2_grad: bool = False, t_grad: bool = False, d_grad: bool = False,
                 ior_glass_grad: bool = False, ior_media_grad: bool = False,
                 fresnel: bool = False, inked: bool = False, transform: Optional[G.RayTransform] = None):
        super().__init__()
        self.ior_glass = _scalar_param(ior_glass, ior_glass_grad)
        self.ior_media = _scalar_param(ior_media, ior_media_grad)
        self.shape = self._make_shape(c1, c2, d, t, c1_grad, c2_grad, t_grad, d_grad, transform)
        self.surface_functions.append(_snell(self.ior_glass, self.ior_media, fresnel))
        self.surface_functions.append(_snell(self.ior_media, self.ior_glass, fresnel))
        self.surface_functions.append(P.Block() if inked else _snell(self.ior_glass, self.ior_media, fresnel))

    def _make_shape(self, c1, c2, d, t, c1_grad, c2_grad, t_grad, d_grad, transform):
        return G.Singlet(C1=c1, C2=c2, D=d, T=t, C1_grad=c1_grad, C2_grad=c2_grad,
                         D_grad=d_grad, T_grad=t_grad, transform=transform)

    # paraxial conveniences (scalar parameter math, elements/lens.py:60-104)
    @property
    def power1(self):
        return self.shape.surfaces[0].c * (self.ior_glass - self.ior_media)

    @property
    def power2(self):
        return self.shape.surfaces[1].c * (self.ior_media - self.ior_glass)

    @property
    def Power(self):
        return self.power1 + self.power2 - self.power1 * self.power2 * (self.T / self.ior_glass)

    @property
    def f(self):
        return 1 / self.Power

    @property
    def f_bfl(self):
        return self.f * (1 - self.T * self.power1 / self.ior_glass)

    @property
    def f_ffl(self):
        return -self.f * (1 - self.T * self.power2 / self.ior_glass)

    @property
    def R1(self):
        return 1 / self.shape.surfaces[0].c

    @property
    def R2(self):
        return -1 / self.shape.surfaces[1].c

    @property
    def T(self):
        return self.shape.T

    @property
    def T_edge(self):
        return self.shape.T_edge


class CylSingletLens(SingletLens):
    """Cylindrical singlet; surfaces [front, back, +x, -x, +y, -y] (elements/lens.py:185-208)."""

    def __init__(self, c1, c2, height, width, t, ior_glass, ior_media=1.0,
                 c1_grad=False, c2_grad=False, t_grad=False, height_grad=False, width_grad=False,
                 ior_glass_grad=False, ior_media_grad=False,
                 fresnel=False, inked=False, transform: Optional[G.RayTransform] = None):
        self._cyl = dict(width=width, height=height, w_grad=width_grad, h_grad=height_grad)
        super().__init__(c1, c2, height, t, ior_glass, ior_media=ior_media,
                         c1_grad=c1_grad, c2_grad=c2_grad, t_grad=t_grad,
                         ior_glass_grad=ior_glass_grad, ior_media_grad=ior_media_grad,
                         fresnel=fresnel, inked=inked, transform=transform)
        for _ in range(3):
            self.surface_functions.append(self.surface_functions[-1])
        self.Nsurfaces = 6

    def _make_shape(self, c1, c2, d, t, c1_grad, c2_grad, t_grad, d_grad, transform):
        q = self._cyl
        return G.CylSinglet(C1=c1, C2=c2, width=q["width"], height=q["height"], T=t,
                            C1_grad=c1_grad, C2_grad=c2_grad, T_grad=t_grad,
                            w_grad=q["w_grad"], h_grad=q["h_grad"], transform=transform)


class _CementedLens(Element):
    """Shared construction of cemented stacks: faces refract media->g1->...->media,
    edges absorb (elements/lens.py:231-279, 325-387)."""

    def _finish(self, iors, n_edges, fresnel=False):
        for a, b in zip(iors[:-1], iors[1:]):
            self.surface_functions.append(_snell(a, b, fresnel))
        for _ in range(n_edges):
            self.surface_functions.append(P.Block())


class DoubletLens(_CementedLens):
    def __init__(self, c1, c2, c3, d, t1, t2, ior_glass1, ior_glass2, ior_media=1.0,
                 c1_grad=False, c2_grad=False, c3_grad=False, t1_grad=False, t2_grad=False, d_grad=False,
                 ior_glass1_grad=False, ior_glass2_grad=False, ior_media_grad=False,
                 fresnel=False, inked=True, transform: Optional[G.RayTransform] = None):
        super().__init__()
        self.ior_glass1 = _scalar_param(ior_glass1, ior_glass1_grad)
        self.ior_glass2 = _scalar_param(ior_glass2, ior_glass2_grad)
        self.ior_media = _scalar_param(ior_media, ior_media_grad)
        self.shape = G.Doublet(C1=c1, C2=c2, C3=c3, D=d, T1=t1, T2=t2,
                               C1_grad=c1_grad, C2_grad=c2_grad, C3_grad=c3_grad, D_grad=d_grad,
                               T1_grad=t1_grad, T2_grad=t2_grad, transform=transform)
        self._finish([self.ior_media, self.ior_glass1, self.ior_glass2, self.ior_media], 2, fresnel)

    @property
    def T1(self):
        return self.shape.T1

    @property
    def T2(self):
        return self.shape.T2


class TripletLens(_CementedLens):
    def __init__(self, c1, c2, c3, c4, d, t1, t2, t3, ior_glass1, ior_glass2, ior_glass3, ior_media=1.0,
                 c1_grad=False, c2_grad=False, c3_grad=False, c4_grad=False,
                 t1_grad=False, t2_grad=False, t3_grad=False, d_grad=False,
                 ior_glass1_grad=False, ior_glass2_grad=False, ior_glass3_grad=False, ior_media_grad=False,
                 fresnel=False, inked=True, transform: Optional[G.RayTransform] = None):
        super().__init__()
        self.ior_glass1 = _scalar_param(ior_glass1, ior_glass1_grad)
        self.ior_glass2 = _scalar_param(ior_glass2, ior_glass2_grad)
        self.ior_glass3 = _scalar_param(ior_glass3, ior_glass3_grad)
        self.ior_media = _scalar_param(ior_media, ior_media_grad)
        self.shape = G.Triplet(C1=c1, C2=c2, C3=c3, C4=c4, D=d, T1=t1, T2=t2, T3=t3,
                               C1_grad=c1_grad, C2_grad=c2_grad, C3_grad=c3_grad, C4_grad=c4_grad,
                               D_grad=d_grad, T1_grad=t1_grad, T2_grad=t2_grad, T3_grad=t3_grad,
                               transform=transform)
        self._finish([self.ior_media, self.ior_glass1, self.ior_glass2, self.ior_glass3, self.ior_media], 3, fresnel)

    @property
    def T1(self):
        return self.shape.T1

    @property
    def T2(self):
        return self.shape.T2

    @property
    def T3(self):
        return self.shape.T3


# ---- mirrors (elements/mirror.py): shape is a bare Surface, one Reflect function ----------
class Mirror(Element):
    def __init__(self):
        super().__init__()
        self.surface_functions.append(P.Reflect())

    @property
    def c1(self):
        return self.shape.c

    @property
    def R(self):
        return 1.0 / self.shape.c

    @property
    def f(self):
        return 1.0 / (2.0 * self.shape.c)


class SphericalMirror(Mirror):
    def __init__(self, c1: float, d: float, diameter: float = float("inf"),
                 c1_grad: bool = False, d_grad: bool = False, diameter_grad: bool = False,
                 transform: Optional[G.RayTransform] = None):
        super().__init__()
        self.shape = G.BoundedHalfSphere(curvature=c1, diameter=diameter, curvature_grad=c1_grad,
                                         diameter_grad=diameter_grad, transform=transform)
        self.d = _scalar_param(d, d_grad)


class CylindricalMirror(Mirror):
    def __init__(self, c1: float, d: float, c1_grad: bool = False, d_grad: bool = False,
                 transform: Optional[G.RayTransform] = None):
        super().__init__()
        self.shape = G.HalfCyl(curvature=c1, curvature_grad=c1_grad, transform=transform)
        self.d = _scalar_param(d, d_grad)


class ParabolicMirror(Mirror):
    def __init__(self, c1: float, d: float, c1_grad: bool = False, d_grad: bool = False,
                 transform: Optional[G.RayTransform] = None):
        super().__init__()
        self.shape = G.Quadric(c=c1, k=-1.0, c_grad=c1_grad, transform=transform)
        self.d = _scalar_param(d, d_grad)


class ParabolicMirrorXZ(Mirror):
    """QuadricZY rotated pi/2 about z so it focuses in x (elements/mirror.py:126-145)."""

    def __init__(self, c1: float, d: float, c1_grad: bool = False, d_grad: bool = False,
                 transform: Optional[G.RayTransform] = None):
        super().__init__()
        import math
        trans = transform.trans.detach().tolist() if transform is not None else None
        self.shape = G.QuadricZY(c=c1, k=-1.0, c_grad=c1_grad,
                                 transform=G.RayTransform(rotation=[0.0, 0.0, math.pi / 2.0], translation=trans))
        self.d = _scalar_param(d, d_grad)


# ---- apertures (elements/aperture.py) ------------------------------------------------------
class _Aperture(Element):
    def _set(self, surface):
        self.shape = surface
        self.surface_functions.append(P.ApertureFilter(surface.inBounds))


class CircularAperture(_Aperture):
    def __init__(self, radius: float, invert: bool = False, transform: Optional[G.RayTransform] = None):
        super().__init__()
        self._set(G.Disk(radius=radius, invert=invert, transform=transform))

    @property
    def radius(self):
        return self.shape.radius


class RectangularAperture(_Aperture):
    def __init__(self, half_x: float, half_y: float, invert: bool = False,
                 transform: Optional[G.RayTransform] = None):
        super().__init__()
        self._set(G.Rectangle(half_x=half_x, half_y=half_y, invert=invert, transform=transform))

    @property
    def half_x(self):
        return self.shape.hx

    @property
    def half_y(self):
        return self.shape.hy


class EllipticAperture(_Aperture):
    def __init__(self, r_major: float, r_minor: float, rot: float = 0.0, invert: bool = False,
                 transform: Optional[G.RayTransform] = None):
        super().__init__()
        self._set(G.Ellipse(r_major=r_major, r_minor=r_minor, rot=rot, invert=invert, transform=transform))

    @property
    def r_major(self):
        return self.shape.r_major

    @property
    def r_minor(self):
        return self.shape.r_minor


# ---- sensor (elements/sensor.py) -----------------------------------------------------------
class LinearElement(Element):
    """A plane with ``Linear`` physics bound to the plane's own pose (elements/ideal.py:47-61)."""

    def __init__(self, shape: G.Plane, linSurfFunc: P.Linear):
        super().__init__()
        self.shape = shape
        linSurfFunc.transform = shape.transform
        self.surface_functions.append(linSurfFunc)


def _ideal_plane(diameter, transform):
    return G.Plane(transform=transform) if diameter == float("inf") else G.Disk(radius=diameter / 2, transform=transform)


class IdealThinLens(LinearElement):
    """Thin lens of focal length ``focal``: P = -1/f on both axes (elements/ideal.py:64-87)."""

    def __init__(self, focal: float, focal_grad: bool = False, diameter: float = float("inf"),
                 transform: Optional[G.RayTransform] = None):
        super().__init__(shape=_ideal_plane(diameter, transform), linSurfFunc=P.Linear())
        self.P = nn.Parameter(torch.as_tensor(-1.0 / focal), requires_grad=focal_grad)
        self.surface_functions[0].Cx = self.P
        self.surface_functions[0].Cy = self.P

    @property
    def f(self):
        return -1 / self.P


class IdealCylThinLens(LinearElement):
    """Thin lens with separate focal lengths in x and y (elements/ideal.py:90-119).  The reference assigns ``Cy``
    to ``surface_functions[1]``, which does not exist (IndexError at construction); here both powers go to the
    element's one surface function, which is what the class documents."""

    def __init__(self, focal_x: float, focal_y: float, focal_x_grad: bool = False, focal_y_grad: bool = False,
                 diameter: float = float("inf"), transform: Optional[G.RayTransform] = None):
        super().__init__(shape=_ideal_plane(diameter, transform), linSurfFunc=P.Linear())
        self.Px = nn.Parameter(torch.as_tensor(-1.0 / focal_x), requires_grad=focal_x_grad)
        self.Py = nn.Parameter(torch.as_tensor(-1.0 / focal_y), requires_grad=focal_y_grad)
        self.surface_functions[0].Cx = self.Px
        self.surface_functions[0].Cy = self.Py

    @property
    def fx(self):
        return -1 / self.Px

    @property
    def fy(self):
        return -1 / self.Py


class IdealMirror(LinearElement):
    """Paraxial mirror of radii (Rx, Ry): Px = -2/Rx, Py = -2/Ry (elements/ideal.py:122-163).  Like the reference's
    ``Linear`` physics it keeps rays travelling towards +z of the plane (new local direction z = +1)."""

    def __init__(self, radius_x: float, radius_y: float, radius_x_grad: bool = False, radius_y_grad: bool = False,
                 diameter: float = float("inf"), transform: Optional[G.RayTransform] = None):
        super().__init__(shape=_ideal_plane(diameter, transform), linSurfFunc=P.Linear())
        self.Px = nn.Parameter(torch.as_tensor(-2.0 / radius_x), requires_grad=radius_x_grad)
        self.Py = nn.Parameter(torch.as_tensor(-2.0 / radius_y), requires_grad=radius_y_grad)
        self.surface_functions[0].Cx = self.Px
        self.surface_functions[0].Cy = self.Py

    @property
    def fx(self):
        return -1 / self.Px

    @property
    def fy(self):
        return -1 / self.Py

    @property
    def Rx(self):
        return -2 / self.Px

    @property
    def Ry(self):
        return -2 / self.Py


class Sensor(Element):
    """Transmitting surface that records ``(hit_local, intensity_before, id)`` per call
    (elements/sensor.py:9-65).  The fused scene kernels fill the same three lists, and —
    new relative to the reference — can also bin the hits into ``self.image`` using the
    histogram rule of ``gui/workbench.py:615-624`` (see ``set_image``)."""

    def __init__(self, shape):
        super().__init__()
        self.shape = shape
        self.surface_functions.extend([P.Transmit()] * len(shape))
        self._locs, self._w, self._ids, self._pending = [], [], [], []
        self.image_spec = None      # (H, W, x0, x1, y0, y1, n_channels)
        self.image = None

    # The fused scene kernels hand over (record [N,4], hit mask [N], ids [N]); compaction to the
    # reference's per-call lists needs a boolean gather (a host sync), so it is deferred until
    # somebody actually reads the lists.
    def _pend(self, record, hit, ids, overflow=None):
        """``overflow = (counts, depth)``: per-ray interaction counts of a non-sequential trace, handed over
        with the last kept ordinal so that dropped interactions are reported when the lists are read."""
        self._pending.append((record, hit, ids, overflow))

    def _flush(self):
        for record, hit, ids, overflow in self._pending:
            if overflow is not None:
                counts, depth = overflow
                most = int(counts.max()) if counts.numel() else 0
                if most > depth:
                    import warnings
                    warnings.warn(f"Sensor: a ray interacted {most} times with this sensor in one non-sequential "
                                  f"trace but only {depth} interactions per ray were kept; raise Scene.record_depth")
                if depth > 1 and most < depth:
                    continue                               # nobody got this far: no empty list entry
            hit = hit() if callable(hit) else hit          # masks / ids are built only when somebody reads the lists
            ids = ids() if callable(ids) else ids
            sel = record[hit]
            self._locs.append(sel[:, :3])
            self._w.append(sel[:, 3])
            self._ids.append(ids[hit])
        self._pending = []

    @property
    def hitLocs(self):
        self._flush()
        return self._locs

    @hitLocs.setter
    def hitLocs(self, v):
        self._pending, self._locs = [], v

    @property
    def hitIntensity(self):
        self._flush()
        return self._w

    @hitIntensity.setter
    def hitIntensity(self, v):
        self._w = v

    @property
    def hitID(self):
        self._flush()
        return self._ids

    @hitID.setter
    def hitID(self, v):
        self._ids = v

    def set_image(self, height: int, width: int, extent=None, channels: int = 1):
        """Ask the scene kernels to accumulate an intensity image on this sensor.

        ``extent = (x0, x1, y0, y1)`` in the sensor's local frame; default = the bounding
        box of the sensor surface (Rectangle half sizes / Disk radius).  ``channels`` > 1
        bins by the ray's wavelength index."""
        if extent is None:
            s = self.shape
            if hasattr(s, "hx"):
                hx, hy = float(s.hx.detach()), float(s.hy.detach())
                extent = (-hx, hx, -hy, hy)
            elif hasattr(s, "radius"):
                r = float(s.radius.detach())
                extent = (-r, r, -r, r)
            else:
                raise ValueError("extent required for this sensor shape")
        self.image_spec = (int(height), int(width), *map(float, extent), int(channels))
        self.image = None

    def forward(self, rays, surf_idx):
        from .ops import element_step
        new_pos, new_dir, mod, hit_local = element_step(self, rays, int(surf_idx))
        self.record(hit_local, rays.intensity, rays.id)
        return new_pos, new_dir, mod

    def record(self, hit_local, intensity, ids):
        self._flush()
        self._locs.append(hit_local)
        self._w.append(intensity)
        self._ids.append(ids)

    def reset(self):
        self._locs, self._w, self._ids, self._pending = [], [], [], []
        self.image = None

    # ---- per-id spot sizes (elements/sensor.py:67-176) ----------------------------------------------------------
    def _dense_records(self):
        """(rec [M,4], ids [M]) of every pending fused-kernel trace when ALL recorded hits are still in their dense,
        un-gathered form on a CUDA device (no boolean compaction, no host sync); None otherwise."""
        if self._locs or not self._pending:
            return None
        recs, ids = [], []
        for record, _hit, rid, overflow in self._pending:
            if overflow is not None or not record.is_cuda:
                return None
            recs.append(record.reshape(-1, 4))
            rid = rid() if callable(rid) else rid
            ids.append(rid.reshape(-1))
        return (recs[0], ids[0]) if len(recs) == 1 else (torch.cat(recs, 0), torch.cat(ids, 0))

    def getSpotSizeParallel_xy(self, query_ids, target_xy=None, norm_ord=2):
        """Per-ray-id spot sizes, all ids at once (elements/sensor.py:87-176): for each id in ``query_ids``
        sum_i w_i (|x_i - cx|^p + |y_i - cy|^p) / (2 sum_i w_i) about the id's intensity centroid or its row of
        ``target_xy`` [K,2] (given in the order of ``query_ids``).  Returns ``(spot_sizes [K], intensity_sum [K])``
        like the reference: spot sizes in the order of ``query_ids``, the intensity sums in the order of the SORTED
        ids (the reference un-sorts only its first output).

        After a fused CUDA trace this is two reduction kernels over the dense sensor records (rtt_spot_id_*), no hit
        lists are built; otherwise the same sums are taken with torch index_add on the hit lists."""
        q = torch.as_tensor(query_ids).reshape(-1)
        order = torch.argsort(q.to(torch.int64))
        dense = self._dense_records()
        if dense is not None:
            from . import ops
            rec, ids = dense
            size, wsum = ops.spot_size_per_id(rec, ids, q, target_xy, norm_ord)
            return size, wsum[order.to(wsum.device)]
        locs, w, ids = self.getHitsTensors()
        K = q.numel()
        lut = torch.full((256,), -1, dtype=torch.int64, device=ids.device)
        lut[q.to(torch.int64).to(ids.device) + 128] = torch.arange(K, device=ids.device)
        grp = lut[ids.to(torch.int64) + 128]
        keep = grp >= 0
        xy, w, grp = locs[keep, :2], w[keep], grp[keep]
        wsum = torch.zeros(K, dtype=w.dtype, device=w.device).index_add(0, grp, w)
        safe = torch.where(wsum == 0, torch.ones_like(wsum), wsum)
        if target_xy is None:
            centres = torch.zeros(K, 2, dtype=xy.dtype, device=xy.device).index_add(0, grp, xy * w[:, None]) / safe[:, None]
        else:
            centres = torch.as_tensor(target_xy, dtype=xy.dtype, device=xy.device).reshape(K, 2)
        moment = (w[:, None] * (xy - centres[grp]).abs() ** norm_ord).sum(1)
        size = torch.zeros(K, dtype=xy.dtype, device=xy.device).index_add(0, grp, moment) / (2 * safe)
        return size, wsum[order.to(wsum.device)]

    def getSpotSizeID_xy(self, ray_id, target_xy=None, norm_ord=2):
        """Per-axis weighted moment [2] of ONE ray id about its centroid / ``target_xy`` [2]:
        sum_{i in id} w_i |xy_i - c|^p / sum_i w_i (elements/sensor.py:67-85).  The reference's method raises at this
        snapshot (it indexes a 0-dim centroid with [None, :] and multiplies the un-masked intensities into the masked
        hits); this is what its formula states once the id mask is applied consistently: centroid and moment over the
        id's hits, normalised by the TOTAL recorded intensity as written there."""
        locs, w, ids = self.getHitsTensors()
        keep = ids == int(ray_id)
        xy, wi = locs[keep, :2], w[keep]
        total = w.sum()
        c = (xy * wi[:, None]).sum(0) / total if target_xy is None else \
            torch.as_tensor(target_xy, dtype=xy.dtype, device=xy.device).reshape(2)
        return (wi[:, None] * (xy - c[None, :]).abs() ** norm_ord).sum(0) / total

    def getHitsTensors(self, ray_id=None):
        """(locs [M,3], intensities [M], ids [M]) over all recorded hits
        (elements/sensor.py:46-65)."""
        locs = torch.cat(self.hitLocs, dim=0)
        w = torch.cat(self.hitIntensity, dim=0)
        ids = torch.cat(self.hitID, dim=0)
        if ray_id is not None:
            keep = ids == int(ray_id)
            locs, w, ids = locs[keep], w[keep], ids[keep]
        return locs, w, ids
