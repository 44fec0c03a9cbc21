"""Ray container and ray sources (mirror of ``rays/ray.py`` and ``rays/bundle.py``).

``Rays`` is the I/O contract of the hot path: ``pos[N,3] f32``, ``dir[N,3] f32``,
``intensity[N] f32``, ``id[N] i8``, ``wavelength[N] f32`` (rays/ray.py:7-19).  The
reference builds it on ``tensordict.tensorclass``; here it is a small plain class with the
same constructor keywords and methods, because the only behaviours the hot path relies on
are: direction renormalisation at construction (rays/ray.py:22-25), mask indexing that does
NOT renormalise (used by scene/sequential.py:29), ``with_coords`` and ``scatter_update``.

Ray sources draw from torch's generator in the same order as the reference so that a
seeded script produces the same bundle.
"""
from __future__ import annotations

import math
from typing import Optional, Union

import torch
import torch.nn as nn
import torch.nn.functional as F

from .geom import RayTransformBundle

_FIELDS = ("pos", "dir", "intensity", "id", "wavelength")


class Rays:
    def __init__(self, *, pos, dir, intensity, id, wavelength, batch_size=None):
        self.pos = pos
        self.dir = F.normalize(dir, p=2, dim=1)           # rays/ray.py:25
        self.intensity = intensity
        self.id = id
        self.wavelength = wavelength
        self.batch_size = torch.Size(batch_size if batch_size is not None else pos.shape[:1])

    @classmethod
    def _wrap(cls, **fields):
        """Build without renormalising (what tensorclass indexing / .to / .clone do)."""
        self = object.__new__(cls)
        for k in _FIELDS:
            setattr(self, k, fields[k])
        self.batch_size = torch.Size(fields["pos"].shape[:1])
        return self

    def _map(self, fn):
        return Rays._wrap(**{k: fn(getattr(self, k)) for k in _FIELDS})

    def __getitem__(self, idx):
        return self._map(lambda t: t[idx])

    def to(self, *a, **k):
        return self._map(lambda t: t.to(*a, **k))

    def clone(self):
        return self._map(lambda t: t.clone())

    def __len__(self):
        return int(self.batch_size[0])

    @property
    def device(self):
        return self.pos.device

    def with_coords(self, new_pos, new_dir):
        """New Rays sharing metadata; renormalises like the reference (rays/ray.py:84-97)."""
        return Rays(pos=new_pos, dir=new_dir, intensity=self.intensity, id=self.id,
                    wavelength=self.wavelength, batch_size=self.batch_size)

    def scatter_update(self, mask, new_pos, new_dir, intensity_mod):
        """Masked write-back; ``intensity *= mod`` on the masked rays (rays/ray.py:29-40)."""
        idx = (mask,)
        self.pos = self.pos.index_put(idx, new_pos)
        self.dir = self.dir.index_put(idx, new_dir)
        self.intensity = self.intensity.index_put(idx, self.intensity[mask] * intensity_mod)

    @classmethod
    def initialize(cls, origins, directions, wavelengths=None, intensities=None, ray_id: int = 0,
                   device: Union[str, torch.device] = "cpu", dtype: torch.dtype = torch.float32):
        """Factory with broadcasting and defaults (rays/ray.py:43-82)."""
        o = torch.as_tensor(origins, device=device, dtype=dtype)
        d = torch.as_tensor(directions, device=device, dtype=dtype)
        o = o.unsqueeze(0) if o.ndim == 1 else o
        d = d.unsqueeze(0) if d.ndim == 1 else d
        n = o.shape[0]
        w = torch.ones(n, device=device, dtype=dtype) if intensities is None else \
            torch.as_tensor(intensities, device=device, dtype=dtype)
        lam = torch.zeros(n, device=device, dtype=dtype) if wavelengths is None else \
            torch.as_tensor(wavelengths, device=device, dtype=dtype)
        ids = torch.full((n,), ray_id, dtype=torch.int8, device=device)
        return cls(pos=o, dir=d, intensity=w, id=ids, wavelength=lam, batch_size=[n])


def _uniform(n, lo, hi):
    """Uniform(lo, hi).sample((n,)) for 1-element tensors: rand[n,1]*(hi-lo)+lo."""
    return lo + torch.rand((n,) + tuple(lo.shape), dtype=lo.dtype, device=lo.device) * (hi - lo)


class Bundle(nn.Module):
    """Source = local sampler + local->global pose (rays/bundle.py:9-37)."""

    def __init__(self, ray_id: int, device: Union[str, torch.device] = "cpu",
                 dtype: torch.dtype = torch.float32, transform: Optional[RayTransformBundle] = None):
        super().__init__()
        self.ray_id, self.device, self.dtype = ray_id, device, dtype
        self.transform = RayTransformBundle(dtype=dtype) if transform is None else transform

    def sample_dir(self, N: int):
        return torch.tensor([[0, 0, 1]], device=self.device, dtype=self.dtype).repeat(N, 1)

    def sample_pos(self, N: int):
        return torch.zeros((N, 3), device=self.device, dtype=self.dtype)

    def sample(self, N: int) -> Rays:
        p, d = self.transform.transform_(self.sample_pos(N), self.sample_dir(N))
        return Rays.initialize(p, d, ray_id=self.ray_id, device=self.device, dtype=self.dtype)


class DiskSample:
    """Area-uniform annulus sector sampler (rays/bundle.py:40-56): theta first, then r."""

    def __init__(self, radius_inner_2, radius_outer_2, theta_min, theta_max):
        self.r2 = (radius_inner_2, radius_outer_2)
        self.th = (theta_min, theta_max)

    def sample(self, N: int):
        theta = _uniform(N, *self.th).squeeze()
        r = torch.sqrt(_uniform(N, *self.r2)).squeeze()
        return torch.stack([r * torch.cos(theta), r * torch.sin(theta), torch.zeros_like(r)], dim=1)


class SolidAngleSample:
    """Uniform-in-solid-angle cone sampler (rays/bundle.py:58-80): phi first, then theta."""

    def __init__(self, F_phi_min, F_phi_max, theta_min, theta_max):
        self.Fp = (F_phi_min, F_phi_max)
        self.th = (theta_min, theta_max)

    @classmethod
    def invCDF_phi(cls, Fv):
        return torch.acos(-2 * Fv + 1)

    @classmethod
    def CDF_phi(cls, phi):
        return (1 - torch.cos(phi)) / torch.pi

    def sample(self, N: int):
        phi = self.invCDF_phi(_uniform(N, *self.Fp)).squeeze()
        theta = _uniform(N, *self.th).squeeze()
        return phi, theta


class CollimatedDisk(Bundle):
    def __init__(self, radius: float, ray_id: int, device="cpu", dtype=torch.float32,
                 transform: Optional[RayTransformBundle] = None):
        super().__init__(ray_id, device, dtype, transform)
        self.radius2 = torch.as_tensor(radius * radius, device=device, dtype=dtype)
        self.zero = torch.tensor([0.0], device=device, dtype=dtype)
        self.tmax = torch.tensor([2 * math.pi], device=device, dtype=dtype)
        self.disk = DiskSample(self.zero, self.radius2, self.zero, self.tmax)

    def sample_pos(self, N: int):
        return self.disk.sample(N)


class CollimatedLine(Bundle):
    def __init__(self, length: float, ray_id: int, device="cpu", dtype=torch.float32,
                 transform: Optional[RayTransformBundle] = None):
        super().__init__(ray_id, device, dtype, transform)
        self.length_2 = torch.tensor([length], device=device, dtype=dtype)

    def sample_pos(self, N: int):
        x = _uniform(N, -self.length_2, self.length_2)
        return torch.cat([x, torch.zeros((N, 2), device=self.device, dtype=self.dtype)], dim=1)


class Fan(Bundle):
    """2-D fan spreading in y (rays/bundle.py:121-140)."""

    def __init__(self, angle: float, ray_id: int, device="cpu", dtype=torch.float32,
                 transform: Optional[RayTransformBundle] = None):
        super().__init__(ray_id, device, dtype, transform)
        self.angle_2 = torch.tensor([angle / 2], device=device, dtype=dtype)

    def sample_dir(self, N):
        th = _uniform(N, -self.angle_2, self.angle_2).squeeze()
        return torch.stack([torch.zeros_like(th), torch.sin(th), torch.cos(th)], dim=1)


class PointSource(Bundle):
    """Cone of half-angle asin(NA) (rays/bundle.py:143-170)."""

    def __init__(self, NA: float, ray_id: int, device="cpu", dtype=torch.float32,
                 transform: Optional[RayTransformBundle] = None):
        super().__init__(ray_id, device, dtype, transform)
        self.zero = torch.tensor([0.0], device=device, dtype=dtype)
        self.twopi = torch.tensor([2 * math.pi], device=device, dtype=dtype)
        self.F_phi_max = SolidAngleSample.CDF_phi(torch.arcsin(torch.tensor(NA, device=device, dtype=dtype)))
        self.angle_dist = SolidAngleSample(self.zero, self.F_phi_max, self.zero, self.twopi)

    def sample_dir(self, N):
        phi, theta = self.angle_dist.sample(N)
        dr = torch.sin(phi)
        return torch.stack([torch.cos(theta) * dr, torch.sin(theta) * dr, torch.cos(phi)], dim=1)
