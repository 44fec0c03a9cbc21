"""Pinhole camera ray source (mirror of ``render/camera.py:16-72``).

``Camera.generate_rays`` feeds the sequential trace of BASELINE config 4.  ``Renderer.render_3d``
(``render/camera.py:191-257``: a single nearest-hit bounce over the non-aperture elements plus Lambert shading)
is built from the fused ops: one ``rtt_trace_nonseq_fwd`` launch finds the winning row of every pixel ray, one
``rtt_surface_step_fwd`` launch per row that won somewhere yields the normals, the shading is elementwise torch.

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


def _is_aperture(el) -> bool:
    return any("ApertureFilter" in type(f).__name__ or "Fuzzy" in type(f).__name__ for f in el.surface_functions)


def _base_color(phys_func) -> torch.Tensor:
    """Colour rules of render/camera.py:259-301 (class-name based, like the reference)."""
    name = type(phys_func).__name__
    if "Reflect" in name:
        return torch.tensor([1.0, 0.6, 0.0])
    if "Block" in name:
        return torch.tensor([0.2, 0.2, 0.2])
    if "Transmit" in name:
        return torch.tensor([0.0, 0.8, 0.2])
    if "RefractSnell" in name or "RefractFresnel" in name:
        n1, n2 = float(getattr(phys_func, "ior_in", 1.5)), float(getattr(phys_func, "ior_out", 1.5))
        n = max(n1, n2)
        white, cyan = torch.tensor([0.9, 0.9, 0.9]), torch.tensor([0.0, 1.0, 1.0])
        blue, navy, purp = torch.tensor([0.3, 0.6, 1.0]), torch.tensor([0.0, 0.0, 0.5]), torch.tensor([0.3, 0.0, 0.3])
        if n <= 1.0:
            return white
        if n <= 1.3:
            return torch.lerp(white, cyan, (n - 1.0) / 0.3)
        if n <= 1.4:
            return torch.lerp(cyan, blue, (n - 1.3) / 0.1)
        if n <= 1.7:
            return torch.lerp(blue, navy, (n - 1.4) / 0.3)
        return torch.lerp(navy, purp, min((n - 1.7) / 0.3, 1.0))
    return torch.tensor([1.0, 0.0, 1.0])


class Renderer:
    """Visual ray cast of a scene (render/camera.py:173-301)."""

    def __init__(self, scene, background_color=(1.0, 1.0, 1.0), light_dir=(-0.5, 1.0, -1.0)):
        self.scene = scene
        dev = scene.map_to_element.device
        self.bg_color = torch.tensor(background_color, dtype=torch.float32, device=dev)
        self.light_dir = F.normalize(torch.as_tensor(light_dir, dtype=torch.float32, device=dev), dim=0)

    def render_3d(self, camera) -> torch.Tensor:
        """[H, W, 3] image on the CPU: nearest hit of every pixel ray over the non-aperture elements, base colour by
        surface physics, 0.3 ambient + 0.7 |n . light| shading, background elsewhere (render/camera.py:191-257).

        ONE kernel launch (rtt_render_shade): the search over all rows, the winner's normal and the shading run in
        the thread that owns the pixel ray; a CUDA camera's rays are generated in the kernel (no ray tensors at all)."""
        from . import ops
        from .table import compile_elements
        self.scene._build_index_maps()
        n = camera.width * camera.height
        dev = self.bg_color.device
        renderable = [el for el in self.scene.elements if not _is_aperture(el)]
        if not renderable:
            return self.bg_color.expand(n, 3).clone().reshape(camera.height, camera.width, 3).cpu()
        with torch.no_grad():
            table = compile_elements(renderable, dispersion=getattr(self.scene, "dispersion", None))
            base = torch.stack([_base_color(el.surface_functions[j]) for el in renderable
                                for j in range(len(el.shape))]).to(device=dev, dtype=torch.float32).contiguous()
            light, bg = self.light_dir.cpu().tolist(), self.bg_color.cpu().tolist()
            if torch.device(camera.device).type == "cuda":
                src = camera.generate_source_rays(samples=1)
                rgb, self.last_rows = torch.ops.rtt_b200.render_shade(
                    None, None, ops.source_cfg_of(src), src.pose, src.state, n, table.f, table.i, base, light, bg)
            else:
                rays = camera.generate_rays().to(dev)
                rgb, self.last_rows = torch.ops.rtt_b200.render_shade(
                    rays.pos.contiguous(), rays.dir.contiguous(), [], None, None, n, table.f, table.i, base, light, bg)
        return rgb.reshape(camera.height, camera.width, 3).cpu()
