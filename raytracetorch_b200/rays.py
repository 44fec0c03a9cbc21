"""Ray container and ray sources (mirror of ``rays/ray.py`` and ``rays/bundle.py``).

``Rays`` is the I/O contract of the hot path: ``pos[N,3] f32``, ``dir[N,3] f32``,
``intensity[N] f32``, ``id[N] i8``, ``wavelength[N] f32`` (rays/ray.py:7-19).  The
reference builds it on ``tensordict.tensorclass``; here it is a small plain class with the
same constructor keywords and methods, because the only behaviours the hot path relies on
are: direction renormalisation at construction (rays/ray.py:22-25), mask indexing that does
NOT renormalise (used by scene/sequential.py:29), ``with_coords`` and ``scatter_update``.

Ray sources draw from torch's generator in the same order as the reference so that a
seeded script produces the same bundle.

On a CUDA device ``Bundle.sample(N)`` does not run the reference's dozen eager ops: it returns
``SourceRays`` — a description (source kind, pose, Philox key and counter) of rays that the
trace kernels GENERATE in registers (rtt_source_t, include/rtt_b200.h).  A trace of such rays
reads no ray input from HBM, and its adjoint regenerates the same rays from the 16-byte state.
Any other consumer that touches ``.pos`` / ``.dir`` / ``.intensity`` gets them materialised by
``rtt_sample_bundle`` (same generator, same rays).
"""
from __future__ import annotations

import math
from typing import Optional, Union

import torch
import torch.nn as nn
import torch.nn.functional as F

from .geom import RayTransformBundle

_FIELDS = ("pos", "dir", "intensity", "id", "wavelength")

# Bundle.sample on a CUDA device returns SourceRays (in-kernel generation).  Set to False to draw the
# samples with torch ops in the reference's order instead (e.g. to reproduce a torch-seeded script).
use_device_sources = True


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


class Paths:
    """Proxy around a ``Rays`` that keeps the position of every ray after each bounce, for the GUI overlay
    (rays/ray.py:100-225).  ``scene.rays = Paths(scene.rays)``; every ``scene.step()`` that moved a ray appends one
    ``[N,3]`` CPU snapshot (``scene.simulate()`` then runs its bounces as single-bounce launches so the history has
    one entry per bounce, like the reference's step loop); ``get_history()`` / ``unwrap()`` as in the reference."""

    def __init__(self, rays: "Rays"):
        self._rays = rays
        self._history = [rays.pos.clone().detach().cpu()]

    pos = property(lambda self: self._rays.pos, lambda self, v: setattr(self._rays, "pos", v))
    dir = property(lambda self: self._rays.dir, lambda self, v: setattr(self._rays, "dir", v))
    intensity = property(lambda self: self._rays.intensity, lambda self, v: setattr(self._rays, "intensity", v))
    id = property(lambda self: self._rays.id)
    wavelength = property(lambda self: self._rays.wavelength)
    batch_size = property(lambda self: self._rays.batch_size)

    def __getitem__(self, idx):
        return self._rays[idx]

    def __len__(self):
        return len(self._rays)

    def scatter_update(self, mask, new_pos, new_dir, intensity_mod):
        self._rays.scatter_update(mask, new_pos, new_dir, intensity_mod)
        self.snapshot()

    def snapshot(self):
        """Append the current positions (called by Scene.step after a bounce that hit something)."""
        self._history.append(self._rays.pos.clone().detach().cpu())

    def unwrap(self) -> "Rays":
        return self._rays

    def get_history(self) -> list:
        return self._history

    def to(self, device):
        self._rays = self._rays.to(device)
        return self


# ---- device ray sources ------------------------------------------------------------------------
_SRC_STATE = {}      # device -> (seed the state was built from, int64[2] tensor {Philox key, counter})


def source_state(device) -> torch.Tensor:
    """Per-device {key, counter} of the in-kernel ray generator, keyed to torch's CUDA seed
    (``torch.manual_seed`` re-keys it).  Lives on the device so that captured CUDA graphs draw
    fresh rays on every replay."""
    device = torch.device(device)
    if device.index is None:
        device = torch.device("cuda", torch.cuda.current_device())
    seed = int(torch.cuda.default_generators[device.index].initial_seed())
    hit = _SRC_STATE.get(device)
    if hit is None or hit[0] != seed:
        hit = (seed, torch.tensor([seed & (2 ** 63 - 1), 0], dtype=torch.int64, device=device))
        _SRC_STATE[device] = hit
    return hit[1]


_RANK_STRIDE = 1 << 40   # Philox counter offset between ranks: 1e12 rays per rank before two shards could overlap


def _rank_first(n: int) -> int:
    """Counter offset (``rtt_source_t.first``) of this rank's shard.  Ranks that share a seed — the usual DDP
    setup — would otherwise generate identical rays and all-reduce world_size copies of one sample."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        return dist.get_rank() * _RANK_STRIDE
    return 0


_IDENT_POSE = {}


def _identity_pose12(device) -> torch.Tensor:
    device = torch.device(device)
    hit = _IDENT_POSE.get(device)
    if hit is None:
        hit = _IDENT_POSE[device] = torch.tensor([1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0], dtype=torch.float32,
                                                 device=device)
    return hit


class SourceRays(Rays):
    """Rays of a device ray source, generated on demand (see the module docstring).

    ``source`` = dict(kind, a, width, height, intensity, wavelength, first); ``pose`` = [12] device
    tensor (R row-major, T); ``state`` = int64[2] snapshot {key, counter} owned by this object."""

    def __init__(self, source: dict, pose: torch.Tensor, state: torch.Tensor, n: int, ray_id: int = 0):
        self.source, self.pose, self.state, self.n, self.ray_id = source, pose, state, int(n), int(ray_id)
        self._fields = {}
        self.batch_size = torch.Size([self.n])

    @property
    def generated(self) -> bool:
        """True while no field has been materialised or overwritten (the kernels can generate)."""
        return not any(k in self._fields for k in ("pos", "dir", "intensity", "wavelength"))

    def _materialise(self):
        from . import ops
        pos, dir_, inten, wav = ops.sample_source(self)
        for k, v in (("pos", pos), ("dir", dir_), ("intensity", inten), ("wavelength", wav)):
            self._fields.setdefault(k, v)

    def _get(self, k):
        if k not in self._fields:
            if k == "id":
                self._fields[k] = torch.full((self.n,), self.ray_id, dtype=torch.int8, device=self.pose.device)
            else:
                self._materialise()
        return self._fields[k]

    pos = property(lambda self: self._get("pos"), lambda self, v: self._fields.__setitem__("pos", v))
    dir = property(lambda self: self._get("dir"), lambda self, v: self._fields.__setitem__("dir", v))
    intensity = property(lambda self: self._get("intensity"), lambda self, v: self._fields.__setitem__("intensity", v))
    wavelength = property(lambda self: self._get("wavelength"), lambda self, v: self._fields.__setitem__("wavelength", v))
    id = property(lambda self: self._get("id"), lambda self, v: self._fields.__setitem__("id", v))

    @property
    def device(self):
        return self.pose.device

    def to(self, *a, **k):
        dev = k.get("device", a[0] if a and isinstance(a[0], (str, torch.device)) else None)
        if dev is not None and torch.device(dev).type == "cuda" and len(a) + len(k) == 1:
            d = torch.device(dev)
            if d.index is None or d.index == self.pose.device.index:
                return self
        return super().to(*a, **k)


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

    # (kind, a[4]) of the device ray source equivalent to sample_pos/sample_dir, or None
    def _source(self):
        from . import _cabi
        if type(self).sample_pos is Bundle.sample_pos and type(self).sample_dir is Bundle.sample_dir:
            return _cabi.SRC_LINE, [0.0, 0.0, 0.0, 0.0]
        return None

    def _pose12(self, device) -> torch.Tensor:
        tr = self.transform
        key = (tr.rot_vec._version, tr.trans._version, tr.rot_vec.device, torch.device(device))
        hit = getattr(self, "_pose_cache", None)
        if hit is None or hit[0] != key:
            with torch.no_grad():
                pose = torch.cat([tr.rot.reshape(9), tr.trans.reshape(3)]).to(device=device, dtype=torch.float32)
            hit = (key, pose.contiguous())
            self._pose_cache = hit
        return hit[1]

    def sample(self, N: int) -> Rays:
        dev = torch.device(self.device)
        src = self._source() if (dev.type == "cuda" and self.dtype == torch.float32 and use_device_sources) else None
        if src is not None and isinstance(self.transform, RayTransformBundle):
            state = source_state(dev)
            snap = state.clone()
            state[1:].add_(int(N))
            spec = dict(kind=src[0], a=list(src[1]), width=0, height=0, intensity=1.0, wavelength=0.0,
                        first=_rank_first(N))
            if torch.is_grad_enabled() and any(p.requires_grad for p in self.transform.parameters()):
                # The source pose is being optimised (rays/bundle.py:30-37 with geom/transform.py:245-276; the
                # reference's tests/test_ideal.py:142-168 differentiates w.r.t. the source origin).  In-kernel
                # generated rays have no autograd edge to the pose, so: same Philox samples, drawn in the LOCAL
                # frame by the kernel (identity pose), then the pose applied with differentiable torch ops.  The
                # trace's adjoint returns d/d pos, d/d dir of these rays and autograd chains them to rot_vec / trans.
                local = SourceRays(spec, _identity_pose12(dev), snap, N, self.ray_id)
                p, d = self.transform.transform_(local.pos, local.dir)
                return Rays.initialize(p, d, ray_id=self.ray_id, device=self.device, dtype=self.dtype)
            return SourceRays(spec, self._pose12(dev), snap, N, self.ray_id)
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
        self._src_a = [0.0, float(radius) * float(radius), 0.0, 2 * math.pi]

    def sample_pos(self, N: int):
        return self.disk.sample(N)

    def _source(self):
        from . import _cabi
        return _cabi.SRC_DISK, self._src_a


class CollimatedLine(Bundle):
    def __init__(self, length: float, ray_id: int, device="cpu", dtype=torch.float32,
                 transform: Optional[RayTransformBundle] = None):
        super().__init__(ray_id, device, dtype, transform)
        self.length_2 = torch.tensor([length], device=device, dtype=dtype)
        self._src_a = [float(length), 0.0, 0.0, 0.0]

    def _source(self):
        from . import _cabi
        return _cabi.SRC_LINE, self._src_a

    def sample_pos(self, N: int):
        x = _uniform(N, -self.length_2, self.length_2)
        return torch.cat([x, torch.zeros((N, 2), device=self.device, dtype=self.dtype)], dim=1)


class Fan(Bundle):
    """2-D fan spreading in y (rays/bundle.py:121-140)."""

    def __init__(self, angle: float, ray_id: int, device="cpu", dtype=torch.float32,
                 transform: Optional[RayTransformBundle] = None):
        super().__init__(ray_id, device, dtype, transform)
        self.angle_2 = torch.tensor([angle / 2], device=device, dtype=dtype)
        self._src_a = [float(angle) / 2, 0.0, 0.0, 0.0]

    def _source(self):
        from . import _cabi
        return _cabi.SRC_FAN, self._src_a

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
        self._src_a = [0.0, float(self.F_phi_max), 0.0, 2 * math.pi]

    def _source(self):
        from . import _cabi
        return _cabi.SRC_POINT, self._src_a

    def sample_dir(self, N):
        phi, theta = self.angle_dist.sample(N)
        dr = torch.sin(phi)
        return torch.stack([torch.cos(theta) * dr, torch.sin(theta) * dr, torch.cos(phi)], dim=1)
