"""Pinhole camera ray source (mirror of ``render/camera.py:16-72``).

Only ``Camera.generate_rays`` is on the hot path (it feeds the sequential trace of BASELINE
config 4).  ``Renderer.render_3d`` (a single nearest-hit bounce plus Lambert shading) is listed as
"next" in SURVEY section 8(f) and is not provided.

Extension: ``generate_rays(samples=k, seed=...)`` draws k jittered sub-pixel samples per pixel
(sample 0 is the reference's pixel-centre ray), which is how config 4 reaches ~1e9 rays.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from .rays import Rays


class Camera:
    def __init__(self, position, look_at, up_vector, fov_deg, width, height, device="cpu"):
        self.device, self.width, self.height, self.fov_deg = device, width, height, fov_deg
        self.origin = torch.tensor(position, dtype=torch.float32, device=device)
        target = torch.tensor(look_at, dtype=torch.float32, device=device)
        up = torch.tensor(up_vector, dtype=torch.float32, device=device)
        self.forward = F.normalize(target - self.origin, dim=0)
        self.right = F.normalize(torch.linalg.cross(self.forward, up), dim=0)
        self.up_cam = torch.linalg.cross(self.right, self.forward)

    def _grids(self):
        scale_y = torch.tan(torch.deg2rad(torch.tensor(self.fov_deg * 0.5)))
        scale_x = scale_y * (self.width / self.height)
        y_grid = torch.linspace(scale_y, -scale_y, self.height, device=self.device)
        x_grid = torch.linspace(-scale_x, scale_x, self.width, device=self.device)
        return x_grid, y_grid, float(scale_x), float(scale_y)

    # ---- device ray source (rtt_source_t kind CAMERA): the trace kernels generate these rays in registers ----
    def source_spec(self, samples: int = 1, first: int = 0) -> dict:
        _xg, _yg, sx, sy = self._grids()
        return dict(kind=4, a=[sx, sy, 0.0, 0.0], width=self.width, height=self.height, intensity=1.0,
                    wavelength=0.0, first=int(first))

    def source_pose(self) -> torch.Tensor:
        """[12]: rows right / up / forward, then the pinhole position."""
        return torch.cat([self.right, self.up_cam, self.forward, self.origin]).float().contiguous()

    def generate_source_rays(self, samples: int = 1, seed: int = 0, first: int = 0, count=None):
        """``samples`` rays per pixel as SourceRays (CUDA cameras): ray g = first + i covers pixel g mod (W*H),
        sample g div (W*H); sample 0 is the reference's pixel-centre ray, the others are jittered by up to half a
        pixel.  ``first`` / ``count`` select a shard (multi-GPU)."""
        from .rays import SourceRays
        n = self.width * self.height * int(samples) - int(first) if count is None else int(count)
        state = torch.tensor([int(seed), 0], dtype=torch.int64, device=self.device)
        return SourceRays(self.source_spec(samples, first), self.source_pose().to(self.device), state, n, 0)

    def generate_rays(self, samples: int = 1, seed: int = 0, pixel_range=None) -> Rays:
        """One ray per pixel (render/camera.py:39-72); ``samples`` > 1 adds jittered sub-pixel rays.
        ``pixel_range=(lo, hi)`` restricts to flat pixel indices [lo, hi) (multi-GPU sharding)."""
        x_grid, y_grid, sx, sy = self._grids()
        yy, xx = torch.meshgrid(y_grid, x_grid, indexing="ij")
        xx, yy = xx.reshape(-1), yy.reshape(-1)
        if pixel_range is not None:
            xx, yy = xx[pixel_range[0]:pixel_range[1]], yy[pixel_range[0]:pixel_range[1]]
        if samples > 1:
            g = torch.Generator(device=self.device).manual_seed(seed)
            dx = 2.0 * sx / max(self.width - 1, 1)
            dy = 2.0 * sy / max(self.height - 1, 1)
            jx = torch.rand((samples, xx.shape[0]), device=self.device, generator=g) - 0.5
            jy = torch.rand((samples, xx.shape[0]), device=self.device, generator=g) - 0.5
            jx[0], jy[0] = 0.0, 0.0                                    # sample 0 = the reference's ray
            xx = (xx.unsqueeze(0) + jx * dx).reshape(-1)
            yy = (yy.unsqueeze(0) + jy * dy).reshape(-1)
        dirs = xx.unsqueeze(1) * self.right + yy.unsqueeze(1) * self.up_cam + self.forward
        origins = self.origin.expand_as(dirs)
        return Rays.initialize(origins, dirs, device=self.device)
