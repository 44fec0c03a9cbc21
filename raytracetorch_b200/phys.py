"""Surface interaction descriptors (mirror of the reference's ``phys`` package).

Like ``geom.py`` these are parameter holders: the arithmetic of each interaction
(``phys/std.py:97-108`` Reflect, ``:123-145`` RefractSnell, ``:227-235`` Transmit,
``:243-254`` Block, ``phys/filter.py:24-33`` ApertureFilter) lives in the CUDA kernels;
``PHYS`` is the code the scene compiler writes into the surface table.

``Linear`` (``phys/std.py:35-88``, the ray-transfer physics of the ideal elements) is code ``PHYS_LINEAR``.

``RefractFresnel`` (``phys/std.py:146-224``) is code ``PHYS_FRESNEL``: the reflect / refract choice is drawn from a
counter-based generator keyed by a seed stored in the table (``table.with_seed``), so parity with the reference's
``torch.rand_like`` is statistical.

Not provided: ``Fuzzy`` (arbitrary Python callable).
"""
from __future__ import annotations

from typing import Callable

import torch
import torch.nn as nn

from . import codes as C


class SurfaceFunction(nn.Module):
    PHYS = C.PHYS_TRANSMIT

    def forward(self, local_intersect, ray_dir, normal, **kwargs):
        """(new_dir [N,3], intensity_mod [N]) — evaluated by the CUDA physics op."""
        from .ops import physics_apply
        return physics_apply(self, local_intersect, ray_dir, normal)


class Transmit(SurfaceFunction):
    PHYS = C.PHYS_TRANSMIT


class Linear(SurfaceFunction):
    """Paraxial ray-transfer physics (phys/std.py:35-88): in the frame of ``transform`` (the plane's own pose —
    ``LinearElement`` rebinds it, elements/ideal.py:54), with slopes (u, v) = (dx, dy)/dz,
    new slopes = (Cx x + Dx u, Cy y + Dy v), new direction = normalize(u', v', 1) rotated back."""

    PHYS = C.PHYS_LINEAR

    def __init__(self, Cx: float = 0, Cy: float = 0, Dx: float = 1, Dy=1,
                 Cx_grad=False, Cy_grad=False, Dx_grad=False, Dy_grad=False, transform=None):
        super().__init__()
        self.Cx = nn.Parameter(torch.as_tensor(float(Cx)), requires_grad=Cx_grad)
        self.Cy = nn.Parameter(torch.as_tensor(float(Cy)), requires_grad=Cy_grad)
        self.Dx = nn.Parameter(torch.as_tensor(float(Dx)), requires_grad=Dx_grad)
        self.Dy = nn.Parameter(torch.as_tensor(float(Dy)), requires_grad=Dy_grad)
        if transform is None:
            from .geom import RayTransform
            transform = RayTransform()
        self.transform = transform


class Reflect(SurfaceFunction):
    PHYS = C.PHYS_REFLECT


class Block(SurfaceFunction):
    PHYS = C.PHYS_BLOCK


class RefractSnell(SurfaceFunction):
    """ior_in: medium on the -normal side, ior_out: medium on the +normal side
    (phys/std.py:126-132).  Lenses rebind these attributes to shared Parameters."""

    PHYS = C.PHYS_SNELL

    def __init__(self, ior_in, ior_out, ior_in_grad: bool = False, ior_out_grad: bool = False):
        super().__init__()
        self.ior_in = nn.Parameter(torch.as_tensor(float(ior_in)), requires_grad=ior_in_grad)
        self.ior_out = nn.Parameter(torch.as_tensor(float(ior_out)), requires_grad=ior_out_grad)


class RefractFresnel(RefractSnell):
    """Stochastic Fresnel interface (phys/std.py:146-224): reflects with probability R, refracts otherwise."""

    PHYS = C.PHYS_FRESNEL


class ApertureFilter(Transmit):
    """Passes rays whose local hit lies inside the (non-inverted) bound of the surface
    it was built from, zeroes direction and intensity otherwise (phys/filter.py:10-33)."""

    PHYS = C.PHYS_APERTURE

    def __init__(self, inBounds: Callable):
        super().__init__()
        self._inBounds = inBounds
        # the bounded surface whose rule the kernels evaluate (not registered as a submodule)
        object.__setattr__(self, "_bound_surface", getattr(inBounds, "__self__", None))
