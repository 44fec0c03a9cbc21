"""Geometry description objects: poses, analytic surfaces, bounded surfaces, shapes.

These classes mirror the *names, constructor arguments and parameter attributes* of the
reference's ``geom`` package (``geom/transform.py``, ``geom/primitives.py``,
``geom/bounded.py``, ``geom/shape.py``, ``geom/spherics.py``, ``geom/cylindrics.py``) so
user scripts keep working, but they hold no ray arithmetic at all: they are parameter
containers that the scene compiler (``table.py``) flattens into surface-table rows, and
every per-ray computation happens in the CUDA kernels.  ``intersectTest`` / ``forward``
route to the single-surface CUDA ops in ``ops.py``.

Each class carries a ``KIND`` / ``BOUND`` / ``SHAPE`` code used by the compiler; the
compiler also accepts the reference's own objects (it dispatches on class names), which
is how the golden fixtures are produced.
"""
from __future__ import annotations

import math
from typing import Optional, Sequence

import torch
import torch.nn as nn

from . import codes as C

Vector3 = Optional[Sequence[float]]
Bool3 = Optional[Sequence[bool]]

intersectEpsilon = 1e-6  # geom/primitives.py:6


def _skew(v: torch.Tensor) -> torch.Tensor:
    """[...,3] rotation vectors -> [...,3,3] skew matrices (geom/transform.py:52-57)."""
    x, y, z = v.unbind(-1)
    o = torch.zeros_like(x)
    return torch.stack([torch.stack([o, -z, y], -1),
                        torch.stack([z, o, -x], -1),
                        torch.stack([-y, x, o], -1)], -2)


def rotation_from_vector(v: torch.Tensor) -> torch.Tensor:
    """R = expm(skew(v)); batched.  Same primitive as geom/transform.py:58."""
    return torch.linalg.matrix_exp(_skew(v))


class RayTransform(nn.Module):
    """Pose = translation + rotation vector (geom/transform.py:10-46).

    global->local is ``(p - trans) @ rot`` and ``d @ rot`` (geom/transform.py:90-93).
    ``trans_mask`` / ``rot_mask`` multiply the gradient (geom/transform.py:29-35,44-46).
    """

    bundle_pose = False

    def __init__(self, rotation: Vector3 = None, translation: Vector3 = None,
                 dtype: torch.dtype = torch.float32,
                 trans_grad: bool = False, trans_mask: Bool3 = None,
                 rot_grad: bool = False, rot_mask: Bool3 = None):
        super().__init__()
        t0 = torch.zeros(3, dtype=dtype) if translation is None else \
            torch.as_tensor(translation).detach().clone().to(dtype)
        r0 = torch.zeros(3, dtype=dtype) if rotation is None else \
            torch.as_tensor(rotation).detach().clone().to(dtype)
        self.trans = nn.Parameter(t0, requires_grad=trans_grad)
        self.rot_vec = nn.Parameter(r0, requires_grad=rot_grad)
        if trans_grad and trans_mask is not None:
            self.register_buffer("trans_mask", torch.as_tensor(trans_mask, dtype=dtype))
            self.trans.register_hook(lambda g: g * self.trans_mask)
        if rot_grad and rot_mask is not None:
            self.register_buffer("rot_mask", torch.as_tensor(rot_mask, dtype=dtype))
            self.rot_vec.register_hook(lambda g: g * self.rot_mask)

    @property
    def rot(self) -> torch.Tensor:
        return rotation_from_vector(self.rot_vec)

    # Plain tensor pose helpers (host-side conveniences used by ray sources; tiny tensors)
    def transform_(self, pos, dir_):
        R = self.rot
        return (pos - self.trans[None, :]) @ R, dir_ @ R

    def invTransform_(self, pos, dir_):
        R = self.rot
        return pos @ R.T + self.trans[None, :], dir_ @ R.T

    def transform(self, rays):
        return self.transform_(rays.pos, rays.dir)

    def invTransform(self, rays):
        return self.invTransform_(rays.pos, rays.dir)


class RayTransformBundle(RayTransform):
    """Source pose: ``transform_`` maps local->global (geom/transform.py:245-276)."""

    bundle_pose = True

    def transform_(self, pos, dir_):
        R = self.rot
        return pos @ R.T + self.trans[None, :], dir_ @ R.T

    def invTransform_(self, pos, dir_):
        R = self.rot
        return (pos - self.trans[None, :]) @ R, dir_ @ R


# ----------------------------------------------------------------------------------------
# Surfaces
# ----------------------------------------------------------------------------------------
class Surface(nn.Module):
    """One analytic surface with its own pose (geom/primitives.py:9-117)."""

    KIND = C.SURF_PLANE
    BOUND = C.BOUND_NONE

    def __init__(self, transform: Optional[RayTransform] = None):
        super().__init__()
        self.epsilon = nn.Parameter(torch.as_tensor(1e-6), requires_grad=False)
        self.transform = RayTransform() if transform is None else transform
        self.invert = False

    def __len__(self):
        return 1

    @property
    def surfaces(self):
        return (self,)

    @property
    def z(self):
        return self.transform.trans[2]

    # --- single-surface CUDA ops -----------------------------------------------------
    def intersectTest(self, rays):
        """[N,1] distance, inf on miss (geom/primitives.py:38-57); CUDA op."""
        from .ops import surface_intersect_test
        return surface_intersect_test(self, rays)

    def forward(self, rays, *unused):
        """(t, hit_global, normal_global, hit_local) (geom/primitives.py:59-96); CUDA op."""
        from .ops import surface_geometry
        return surface_geometry(self, rays, 0)


class Plane(Surface):
    KIND = C.SURF_PLANE


class Sphere(Surface):
    KIND = C.SURF_SPHERE

    def __init__(self, radius: float, radius_grad: bool = False, transform: Optional[RayTransform] = None):
        super().__init__(transform)
        self.radius = nn.Parameter(torch.tensor(float(radius)), requires_grad=radius_grad)


class Cylinder(Surface):
    KIND = C.SURF_CYLINDER

    def __init__(self, radius: float, transform: Optional[RayTransform] = None, radius_grad: bool = False):
        super().__init__(transform)
        self.radius = nn.Parameter(torch.tensor(float(radius)), requires_grad=radius_grad)


class Quadric(Surface):
    """c(x^2+y^2) + c(1+k)z^2 - 2z = 0 (geom/primitives.py:244-343)."""

    KIND = C.SURF_QUADRIC

    def __init__(self, c: float, k: float, transform: Optional[RayTransform] = None,
                 c_grad: bool = False, k_grad: bool = False):
        super().__init__(transform)
        self.c = nn.Parameter(torch.as_tensor(float(c)), requires_grad=c_grad)
        self.k = nn.Parameter(torch.as_tensor(float(k)), requires_grad=k_grad)


class QuadricZY(Quadric):
    """x-invariant quadric (geom/primitives.py:346-395)."""

    KIND = C.SURF_QUADRIC_ZY


class Cone(Surface):
    """Double cone z^2 = slope^2 (x^2 + y^2), vertex at the local origin; slope = dz/dr, 0 = plane
    (geom/primitives.py:398-494)."""

    KIND = C.SURF_CONE

    def __init__(self, slope: float, slope_grad: bool = False, transform: Optional[RayTransform] = None):
        super().__init__(transform)
        self.slope = nn.Parameter(torch.as_tensor(float(slope)), requires_grad=slope_grad)


class SurfaceBounded(Surface):
    """Per-root bound test with optional inversion (geom/bounded.py:9-48)."""

    def __init__(self, transform: Optional[RayTransform] = None, invert: bool = False):
        super().__init__(transform)
        self.invert = bool(invert)

    def inBounds(self, local_pos):
        from .ops import surface_in_bounds
        return surface_in_bounds(self, local_pos)


class SingleCone(Cone, SurfaceBounded):
    """One nappe of the cone: hits with z * slope >= -1e-6 (geom/bounded.py:189-217; like the reference, the
    ``invert`` argument is accepted and ignored)."""

    BOUND = C.BOUND_NAPPE

    def __init__(self, slope: float, slope_grad: bool = False, invert: bool = False,
                 transform: Optional[RayTransform] = None):
        Cone.__init__(self, slope=slope, slope_grad=slope_grad, transform=transform)


class Disk(Plane, SurfaceBounded):
    BOUND = C.BOUND_DISK

    def __init__(self, radius: float, invert: bool = False, transform: Optional[RayTransform] = None):
        super().__init__(transform, invert)
        self.radius = nn.Parameter(torch.as_tensor(float(radius)), requires_grad=False)


class Rectangle(Plane, SurfaceBounded):
    BOUND = C.BOUND_RECT

    def __init__(self, half_x: float, half_y: float, invert: bool = False,
                 transform: Optional[RayTransform] = None):
        super().__init__(transform, invert)
        self.hx = nn.Parameter(torch.as_tensor(half_x, dtype=torch.float32))
        self.hy = nn.Parameter(torch.as_tensor(half_y, dtype=torch.float32))


class Ellipse(Plane, SurfaceBounded):
    BOUND = C.BOUND_ELLIPSE

    def __init__(self, r_major: float, r_minor: float, rot: float,
                 r_major_grad: bool = False, r_minor_grad: bool = False, rot_grad: bool = False,
                 invert: bool = False, transform: Optional[RayTransform] = None):
        super().__init__(transform, invert)
        self.r_minor = nn.Parameter(torch.as_tensor(float(r_minor)), requires_grad=r_minor_grad)
        self.r_major = nn.Parameter(torch.as_tensor(float(r_major)), requires_grad=r_major_grad)
        self.rot = nn.Parameter(torch.as_tensor(float(rot)), requires_grad=rot_grad)


def _sag(c: torch.Tensor, h, tz: torch.Tensor) -> torch.Tensor:
    """Vertex sag at height h plus vertex z (geom/bounded.py:129-139,176-186)."""
    h2 = h ** 2
    return (c * h2) / (1.0 + torch.sqrt(torch.relu(1.0 - c ** 2 * h2))) + tz


class HalfSphere(Quadric, SurfaceBounded):
    """Quadric k=0 clipped to |z c| < 1 + 1e-6 (geom/bounded.py:109-139)."""

    BOUND = C.BOUND_HALF

    def __init__(self, curvature: float, curvature_grad: bool, transform: Optional[RayTransform] = None):
        super().__init__(c=curvature, k=0.0, transform=transform, c_grad=curvature_grad, k_grad=False)

    def sagittalZ(self, radius):
        return _sag(self.c, radius, self.transform.trans[2])


class BoundedHalfSphere(HalfSphere):
    """HalfSphere additionally clipped to a circular aperture (geom/bounded.py:142-159)."""

    BOUND = C.BOUND_HALF_DISK

    def __init__(self, curvature: float, diameter: float, curvature_grad: bool = False,
                 diameter_grad: bool = False, transform: Optional[RayTransform] = None):
        super().__init__(curvature, curvature_grad, transform)
        self.diameter = nn.Parameter(torch.as_tensor(float(diameter)), requires_grad=diameter_grad)


class HalfCyl(QuadricZY, SurfaceBounded):
    """QuadricZY k=0 clipped like HalfSphere (geom/bounded.py:162-186)."""

    BOUND = C.BOUND_HALF

    def __init__(self, curvature: float, curvature_grad: bool, transform: Optional[RayTransform] = None):
        super().__init__(c=curvature, k=0.0, transform=transform, c_grad=curvature_grad, k_grad=False)

    def sagittalZ(self, y_height):
        return _sag(self.c, y_height, self.transform.trans[2])


# ----------------------------------------------------------------------------------------
# Shapes (several surfaces under one element pose)
# ----------------------------------------------------------------------------------------
class Shape(nn.Module):
    """Surfaces sharing an element pose; shape-level validity per surface index
    (geom/shape.py:8-102).  ``SHAPE`` says which rule the kernels apply."""

    SHAPE = C.SHAPE_OPEN

    def __init__(self, transform: Optional[RayTransform] = None):
        super().__init__()
        self.epsilon = nn.Parameter(torch.as_tensor(1e-7), requires_grad=False)
        self.surfaces = nn.ModuleList()
        self.transform = RayTransform() if transform is None else transform

    def __len__(self):
        return len(self.surfaces)

    @property
    def z(self):
        return self.transform.trans[2]

    def intersectTest(self, rays):
        """[N,K] distances with shape-level validity applied (geom/shape.py:25-59); CUDA op."""
        from .ops import shape_intersect_test
        return shape_intersect_test(self, rays)

    def forward(self, rays, surf_idx):
        """Geometry of one member surface (geom/shape.py:61-87); CUDA op."""
        from .ops import surface_geometry
        return surface_geometry(self, rays, int(surf_idx))


class CvxPolyhedron(Shape):
    """Intersection of half spaces; a hit is valid if it is behind every *other* plane
    by 1e-4 (geom/shape.py:104-132)."""

    SHAPE = C.SHAPE_POLY

    def __init__(self, planes_list=None, transform: Optional[RayTransform] = None):
        super().__init__(transform)
        for s in planes_list or ():
            self.surfaces.append(s)


def _face(pos, rot_vec, grad):
    mask = [abs(v) > 1e-5 for v in pos]
    return Plane(RayTransform(rotation=rot_vec, translation=pos, trans_grad=grad, trans_mask=mask))


_HALF_PI = math.pi / 2
# (axis, sign, rotation vector) of the side planes, in the reference's order (geom/shape.py:188-210)
_SIDE_FACES = ((0, +1, (0.0, -_HALF_PI, 0.0)), (0, -1, (0.0, _HALF_PI, 0.0)),
               (1, +1, (_HALF_PI, 0.0, 0.0)), (1, -1, (-_HALF_PI, 0.0, 0.0)))


def _side_planes(width, height, w_grad, h_grad):
    out = []
    for axis, sign, rv in _SIDE_FACES:
        pos = [0.0, 0.0, 0.0]
        pos[axis] = sign * (width if axis == 0 else height) / 2
        out.append(_face(pos, rv, w_grad if axis == 0 else h_grad))
    return out


class Box(CvxPolyhedron):
    """Six planes: +z, -z, +x, -x, +y, -y (geom/shape.py:135-210)."""

    def __init__(self, length: float, width: float, height: float,
                 transform: Optional[RayTransform] = None,
                 l_grad: bool = False, w_grad: bool = False, h_grad: bool = False):
        super().__init__(transform=transform)
        self.surfaces.append(_face([0.0, 0.0, length / 2], (0.0, 0.0, 0.0), l_grad))
        self.surfaces.append(_face([0.0, 0.0, -length / 2], (0.0, math.pi, 0.0), l_grad))
        for p in _side_planes(width, height, w_grad, h_grad):
            self.surfaces.append(p)

    @property
    def length(self):
        return self.surfaces[0].transform.trans[2] - self.surfaces[1].transform.trans[2]

    @property
    def width(self):
        return self.surfaces[2].transform.trans[0] - self.surfaces[3].transform.trans[0]

    @property
    def height(self):
        return self.surfaces[4].transform.trans[1] - self.surfaces[5].transform.trans[1]


class Box4Side(CvxPolyhedron):
    """The four side planes only (geom/shape.py:213-276)."""

    def __init__(self, width: float, height: float, transform: Optional[RayTransform] = None,
                 w_grad: bool = False, h_grad: bool = False):
        super().__init__(transform=transform)
        for p in _side_planes(width, height, w_grad, h_grad):
            self.surfaces.append(p)

    @property
    def width(self):
        return self.surfaces[0].transform.trans[0] - self.surfaces[1].transform.trans[0]

    @property
    def height(self):
        return self.surfaces[2].transform.trans[1] - self.surfaces[3].transform.trans[1]


class Spheric(Shape):
    """Stack of HalfSphere faces followed by Cylinder edges (geom/spherics.py:10-54).

    Faces are valid inside the lens radius; edge j is valid between the rim sags of
    faces j and j+1 (geom/spherics.py:27-46)."""

    SHAPE = C.SHAPE_SPHERIC_FACE
    N_optical = 0

    def _build(self, curvs, z_vertices, c_grads, z_grad, D, D_grad):
        self.N_optical = len(curvs)
        self.radius = nn.Parameter(torch.as_tensor(D / 2.0), requires_grad=D_grad)
        for cv, zv, cg in zip(curvs, z_vertices, c_grads):
            pose = RayTransform(translation=[0.0, 0.0, zv], trans_grad=z_grad,
                                trans_mask=[False, False, True])
            self.surfaces.append(HalfSphere(curvature=cv, curvature_grad=cg, transform=pose))
        for _ in range(len(curvs) - 1):
            edge = Cylinder(D / 2)
            edge.radius = self.radius          # shared parameter (geom/spherics.py:92-93)
            self.surfaces.append(edge)
        self._validate(curvs, z_vertices, D)

    def _validate(self, curvs, z_vertices, D):
        # constructor checks of geom/spherics.py:100-111,176-199,266-283
        for i, cv in enumerate(curvs):
            if abs(0.5 * cv) > 1 / D:
                raise ValueError(f"|R{i + 1}| must be larger than D/2")
        for i in range(len(z_vertices) - 1):
            if z_vertices[i + 1] - z_vertices[i] <= 1e-6:
                raise ValueError(f"Thickness T{i + 1} must be positive" if len(curvs) > 2
                                 else "Thickness T must be positive")
        with torch.no_grad():
            rim = [float(self.surfaces[i].sagittalZ(self.radius)) for i in range(self.N_optical)]
        for i in range(len(rim) - 1):
            if rim[i] > rim[i + 1]:
                raise ValueError("Intersecting optical surfaces" if len(curvs) == 2
                                 else f"Optical surfaces {i + 1} and {i + 2} intersect")

    @property
    def T(self):
        return self.surfaces[self.N_optical - 1].transform.trans[2] - self.surfaces[0].transform.trans[2]

    @property
    def T_edge(self):
        return (self.surfaces[self.N_optical - 1].sagittalZ(self.radius)
                - self.surfaces[0].sagittalZ(self.radius))


class Singlet(Spheric):
    """[front, back, edge] (geom/spherics.py:56-111)."""

    def __init__(self, C1: float, C2: float, D: float, T: float,
                 C1_grad: bool = True, C2_grad: bool = True, D_grad: bool = False, T_grad: bool = True,
                 transform: Optional[RayTransform] = None):
        super().__init__(transform=transform)
        self._build([C1, C2], [-T / 2, T / 2], [C1_grad, C2_grad], T_grad, D, D_grad)


class Doublet(Spheric):
    """[s1, s2, s3, e1, e2] (geom/spherics.py:116-206)."""

    def __init__(self, C1: float, C2: float, C3: float, D: float, T1: float, T2: float,
                 C1_grad: bool = True, C2_grad: bool = True, C3_grad: bool = True, D_grad: bool = False,
                 T1_grad: bool = True, T2_grad: bool = True, transform: Optional[RayTransform] = None):
        super().__init__(transform=transform)
        z1 = -(T1 + T2) / 2.0
        self._build([C1, C2, C3], [z1, z1 + T1, z1 + T1 + T2], [C1_grad, C2_grad, C3_grad],
                    T1_grad or T2_grad, D, D_grad)

    @property
    def T1(self):
        return self.surfaces[1].transform.trans[2] - self.surfaces[0].transform.trans[2]

    @property
    def T2(self):
        return self.surfaces[2].transform.trans[2] - self.surfaces[1].transform.trans[2]


class Triplet(Spheric):
    """[s1..s4, e1, e2, e3] (geom/spherics.py:209-297)."""

    def __init__(self, C1: float, C2: float, C3: float, C4: float, D: float,
                 T1: float, T2: float, T3: float,
                 C1_grad: bool = True, C2_grad: bool = True, C3_grad: bool = True, C4_grad: bool = True,
                 D_grad: bool = False, T1_grad: bool = True, T2_grad: bool = True, T3_grad: bool = True,
                 transform: Optional[RayTransform] = None):
        super().__init__(transform=transform)
        z1 = -(T1 + T2 + T3) / 2.0
        z2 = z1 + T1
        z3 = z2 + T2
        self._build([C1, C2, C3, C4], [z1, z2, z3, z3 + T3], [C1_grad, C2_grad, C3_grad, C4_grad],
                    T1_grad or T2_grad or T3_grad, D, D_grad)

    @property
    def T1(self):
        return self.surfaces[1].transform.trans[2] - self.surfaces[0].transform.trans[2]

    @property
    def T2(self):
        return self.surfaces[2].transform.trans[2] - self.surfaces[1].transform.trans[2]

    @property
    def T3(self):
        return self.surfaces[3].transform.trans[2] - self.surfaces[2].transform.trans[2]


class Cylindric(Shape):
    """HalfCyl faces + four side planes; rectangular aperture with 1e-5 slack, side
    planes valid between the face sags with 1e-4 slack (geom/cylindrics.py:10-55)."""

    SHAPE = C.SHAPE_CYL_FACE
    N_optical = 2


class CylSinglet(Cylindric):
    """[front, back, +x, -x, +y, -y] (geom/cylindrics.py:58-118)."""

    def __init__(self, C1: float, C2: float, width: float, height: float, T: float,
                 C1_grad: bool = True, C2_grad: bool = True, T_grad: bool = True,
                 w_grad: bool = False, h_grad: bool = False, transform: Optional[RayTransform] = None):
        super().__init__(transform=transform)
        for cv, zv, cg in ((C1, -T / 2, C1_grad), (C2, T / 2, C2_grad)):
            if abs(0.5 * cv) > 1 / height:
                raise ValueError("|R1| must be larger than Height/2" if zv < 0
                                 else "|R2| must be larger than Height/2")
            pose = RayTransform(translation=[0.0, 0.0, zv], trans_grad=T_grad,
                                trans_mask=[False, False, True])
            self.surfaces.append(HalfCyl(curvature=cv, curvature_grad=cg, transform=pose))
        for p in _side_planes(width, height, w_grad, h_grad):
            self.surfaces.append(p)
        with torch.no_grad():
            if float(self.surfaces[0].sagittalZ(height / 2)) > float(self.surfaces[1].sagittalZ(height / 2)):
                raise ValueError("Front and back surfaces intersecting")

    @property
    def width(self):
        return self.surfaces[2].transform.trans[0] - self.surfaces[3].transform.trans[0]

    @property
    def height(self):
        return self.surfaces[4].transform.trans[1] - self.surfaces[5].transform.trans[1]
