"""Scenes: the two trace loops of the reference, each executed by ONE fused CUDA kernel.

* ``SequentialScene.simulate(rays)``  — scene/sequential.py:12-36.  The reference runs a
  Python double loop with ~10^4 eager ops and two host syncs per surface; here the whole
  surface stack is one launch (plus one adjoint launch in backward).
* ``Scene.simulate()`` / ``step()`` / ``ray_cast(rays)`` — scene/base.py:129-235, the
  non-sequential nearest-hit bounce loop.

Same public attributes as the reference (``elements``, ``bundles``, ``rays``, ``Nbounces``,
``map_to_element`` / ``map_to_surface``, ``total_surfaces``).  Sensors receive the same three
hit lists; they are materialised lazily (``Sensor.hitLocs`` etc.) so that ``simulate`` itself
never synchronises with the host.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn

from . import codes as C
from . import ops
from .rays import Bundle, Rays, SourceRays
from .table import Dispersion, SceneCompiler


class Scene(nn.Module):
    """Non-sequential scene (scene/base.py:8-289)."""

    def __init__(self):
        super().__init__()
        self.elements = nn.ModuleList()
        self.bundles = nn.ModuleList()
        self._bundle_N_rays: list[int] = []
        self.rays: Optional[Rays] = None
        self.Nbounces = 100
        self.dispersion: Optional[Dispersion] = None
        self.record_hits = True          # fill Sensor.hitLocs/hitIntensity/hitID like the reference
        # Non-sequential traces: sensor interactions kept per ray.  The reference appends one hit-list
        # entry per interaction, and a ray routinely re-hits the sensor plane it has just left (t > 1e-6
        # at an fp32 ulp of ~8e-6), so 2 is the smallest depth that reproduces its lists on
        # transmitting sensors; Sensor warns when a ray had more interactions than were kept.
        self.record_depth = 2
        # False: a sequential trace (and a non-sequential trace of in-kernel generated rays) skips the final
        # pos/dir/intensity outputs and leaves the Rays as they were — what the optimisation goals do while they
        # evaluate, since they only read sensor records
        self.final_rays = True
        self.mode: Optional[int] = None  # None = ops default (FAST)
        self.last_trace = None           # raw kernel outputs of the latest simulate()/step()
        self._compiler = SceneCompiler()
        self._build_index_maps()

    # ---- population (scene/base.py:25-52) ---------------------------------------------------
    def add_element(self, element):
        self.elements.append(element)

    def add_bundle(self, bundle: Bundle, N_rays: int = 200):
        self.bundles.append(bundle)
        self._bundle_N_rays.append(N_rays)

    def clear_elements(self):
        self.elements = nn.ModuleList()
        self._build_index_maps()

    def clear_bundles(self):
        self.bundles = nn.ModuleList()
        self._bundle_N_rays = []
        self.rays = None

    def clear_rays(self):
        self.rays = None

    def set_dispersion(self, dispersion: Optional[Dispersion]):
        """Per-wavelength refractive indices (extension, see table.Dispersion)."""
        self.dispersion = dispersion

    # ---- ray construction (scene/base.py:57-90) ----------------------------------------------
    def _build_rays(self):
        if len(self.bundles) == 0:
            self.rays = None
            return
        batches = [b.sample(n) for b, n in zip(self.bundles, self._bundle_N_rays)]
        if len(batches) == 1:
            self.rays = batches[0]
            return
        cat = lambda k: torch.cat([getattr(r, k) for r in batches], dim=0)
        self.rays = Rays._wrap(pos=cat("pos"), dir=cat("dir"), intensity=cat("intensity"), id=cat("id"),
                               wavelength=cat("wavelength"))

    # ---- flattening (scene/base.py:96-123): row order == table row order -----------------------
    def _build_index_maps(self):
        dev = self.map_to_element.device if isinstance(getattr(self, "map_to_element", None), torch.Tensor) \
            else torch.device("cpu")
        e_idx, s_idx = [], []
        for k, el in enumerate(self.elements):
            n = len(el.shape)
            e_idx += [k] * n
            s_idx += list(range(n))
        self.register_buffer("map_to_element", torch.tensor(e_idx, dtype=torch.long, device=dev))
        self.register_buffer("map_to_surface", torch.tensor(s_idx, dtype=torch.long, device=dev))
        self.total_surfaces = len(e_idx)

    # ---- kernels --------------------------------------------------------------------------------
    def table(self):
        """The compiled surface table.  Scenes with RefractFresnel rows get a fresh seed per call (one per simulate /
        step, like the reference's torch.rand_like per interaction): drawn from torch's generator, so
        ``torch.manual_seed`` makes a run reproducible.  ``self.rng_seed = int`` pins it instead."""
        tab = self._compiler.table(self.elements, dispersion=self.dispersion)
        if tab.stochastic:
            seed = getattr(self, "rng_seed", None)
            if seed is None:
                seed = int(torch.randint(0, 2 ** 62, (1,)).item())
            tab = tab.with_seed(int(seed))
        return tab

    def _deliver_to_sensors(self, table, records, hit_of_slot, rays_before: Rays, images):
        for slot, sensor in enumerate(table.sensors):
            if images[slot] is not None:
                sensor.image = images[slot] if sensor.image is None else sensor.image + images[slot]
            if self.record_hits and records.numel():
                sensor._pend(records[slot], (lambda s=slot: hit_of_slot(s)), lambda: rays_before.id)

    def _trace(self, rays: Rays, nbounces: int):
        table = self._last_table = self.table()
        depth = max(1, min(int(self.record_depth), nbounces))
        src = rays if (isinstance(rays, SourceRays) and rays.generated) else None
        if src is not None:
            out = ops.trace_nonsequential(table, None, None, None, nbounces, want_record=self.record_hits,
                                          mode=self.mode, record_depth=depth, source=src, want_rays=self.final_rays)
        elif table.f.is_cuda and not rays.pos.is_cuda:
            # rays in host memory, scene on the GPU: the host->device copies are pipelined with the bounce loop; the
            # Rays object ends up on the device, like after ``rays.to(device)`` followed by ``simulate``; forward only
            out = ops.trace_nonsequential_host(table, rays.pos, rays.dir, rays.intensity, nbounces, rays.wavelength,
                                               want_record=self.record_hits, mode=self.mode, record_depth=depth,
                                               ids=rays.id)
            rays.id = out["in_id"]
            rays.wavelength = out["in_wavelength"] if out["in_wavelength"] is not None \
                else rays.wavelength.to(table.f.device, non_blocking=True)
        else:
            out = ops.trace_nonsequential(table, rays.pos, rays.dir, rays.intensity, nbounces, rays.wavelength,
                                          want_record=self.record_hits, mode=self.mode, record_depth=depth)
        self.last_trace = out
        records, counts = out["records"], out["sensor_counts"]
        for slot, sensor in enumerate(table.sensors):
            if out["images"][slot] is not None:
                sensor.image = out["images"][slot] if sensor.image is None else sensor.image + out["images"][slot]
            if self.record_hits and records.numel():
                # one list entry per interaction ordinal (the reference: one per bounce; same multiset)
                for k in range(depth):
                    sensor._pend(records[slot, k], (lambda c=counts[slot], k=k: c > k), lambda: rays.id,
                                 overflow=(counts[slot], depth) if k == depth - 1 else None)
        if src is None or self.final_rays:
            rays.pos, rays.dir, rays.intensity = out["pos"], out["dir"], out["intensity"]
        return out

    def simulate(self):
        """Propagate ``self.rays`` for up to ``Nbounces`` bounces in one kernel (scene/base.py:129-142)."""
        if self.rays is None:
            self._build_rays()
        if self.rays is None:
            return
        self._build_index_maps()
        if hasattr(self.rays, "snapshot"):
            # Paths proxy (visualisation): one launch per bounce so every bounce lands in the history,
            # like the reference's loop (scene/base.py:139-142)
            for _ in range(min(int(self.Nbounces), C.MAX_BOUNCES)):
                if not bool((self.rays.intensity > 0).any()):
                    break
                self.step()
            return
        rays = self.rays.unwrap() if hasattr(self.rays, "unwrap") else self.rays
        self._trace(rays, min(int(self.Nbounces), C.MAX_BOUNCES))

    def step(self):
        """One bounce for all active rays (scene/base.py:180-235)."""
        if self.rays is None:
            return
        rays = self.rays.unwrap() if hasattr(self.rays, "unwrap") else self.rays
        out = self._trace(rays, 1)
        if hasattr(self.rays, "snapshot") and bool((out["n_hits"] > 0).any()):
            self.rays.snapshot()              # the reference records in scatter_update, i.e. only when a ray was hit

    def ray_cast(self, rays):
        """(hit_mask, winner_element_ids, winner_surf_ids) or None (scene/base.py:144-178)."""
        if len(self.elements) == 0:
            return None
        raw = rays.unwrap() if hasattr(rays, "unwrap") else rays
        table = self.table()
        with torch.no_grad():
            out = ops.trace_nonsequential(table, raw.pos, raw.dir, torch.ones_like(raw.intensity), 1,
                                          raw.wavelength, want_record=False, sensor_cfg=[], mode=self.mode)
            win = out["hit_seq"][:, 0].long()
            hit_mask = win != 255
            if not bool(hit_mask.any()):
                return None
            win = torch.where(hit_mask, win, torch.zeros_like(win))
            if self.map_to_element.device != win.device:
                self._build_index_maps()
                self.map_to_element = self.map_to_element.to(win.device)
                self.map_to_surface = self.map_to_surface.to(win.device)
            return hit_mask, self.map_to_element[win], self.map_to_surface[win]

    # ---- conversions (scene/base.py:261-289, scene/sequential.py:80-105) ------------------------
    def to_sequential(self):
        ordered = sorted(self.elements, key=lambda el: el.shape.transform.trans[2].item())
        seq = SequentialScene(ordered)
        seq.Nbounces = self.Nbounces
        for b, n in zip(self.bundles, self._bundle_N_rays):
            seq.add_bundle(b, n)
        seq.rays, seq.dispersion = self.rays, self.dispersion
        return seq


class SequentialScene(Scene):
    """Fixed traversal order (scene/sequential.py:7-36)."""

    def __init__(self, elements):
        super().__init__()
        self.elements = nn.ModuleList(elements)
        self._build_index_maps()

    def simulate(self, rays: Optional[Rays] = None):
        """Propagate ``rays`` through every surface of every element, in order; mutates and
        returns the same ``Rays`` object.  ``rays=None`` uses ``self.rays`` (so the goals in
        ``optim`` work with sequential scenes too — the reference's own call raises)."""
        if rays is None:
            if self.rays is None:
                self._build_rays()
            rays = self.rays
        if rays is None:
            return None
        table = self._last_table = self.table()
        src = rays if (isinstance(rays, SourceRays) and rays.generated) else None
        if src is None and table.f.is_cuda and not rays.pos.is_cuda:
            return self._simulate_host_bundle(table, rays)
        if src is not None:     # rays generated in the kernel: no ray input read from HBM
            out = ops.trace_sequential(table, want_record=self.record_hits, mode=self.mode, source=src,
                                       want_rays=self.final_rays)
        else:
            out = ops.trace_sequential(table, rays.pos, rays.dir, rays.intensity, rays.wavelength,
                                       want_record=self.record_hits, mode=self.mode, want_rays=self.final_rays)
        self.last_trace = out
        mask = out["hitmask"]

        def hit_of_slot(slot):
            return ((mask >> table.sensor_rows[slot]) & 1).bool()

        self._deliver_to_sensors(table, out["records"], hit_of_slot, rays, out["images"])
        if self.final_rays:
            rays.pos, rays.dir, rays.intensity = out["pos"], out["dir"], out["intensity"]
        return rays

    def _simulate_host_bundle(self, table, rays: Rays):
        """Rays in host memory, scene on the GPU: host->device copies are pipelined with the trace
        (ops.trace_sequential_host).  The Rays object ends up on the device, like after ``rays.to(device)``
        followed by ``simulate``; forward only."""
        dev = table.f.device
        out = ops.trace_sequential_host(table, rays.pos, rays.dir, rays.intensity, rays.wavelength,
                                        want_record=self.record_hits, mode=self.mode, ids=rays.id)
        self.last_trace = out
        mask = out["hitmask"]
        rays.id = out["in_id"]
        rays.wavelength = out["in_wavelength"] if out["in_wavelength"] is not None \
            else rays.wavelength.to(dev, non_blocking=True)

        def hit_of_slot(slot):
            return ((mask >> table.sensor_rows[slot]) & 1).bool()

        self._deliver_to_sensors(table, out["records"], hit_of_slot, rays, out["images"])
        rays.pos, rays.dir, rays.intensity = out["pos"], out["dir"], out["intensity"]
        return rays

    def to_base(self):
        base = Scene()
        base.Nbounces = self.Nbounces
        for el in self.elements:
            base.add_element(el)
        for b, n in zip(self.bundles, self._bundle_N_rays):
            base.add_bundle(b, n)
        base.rays, base.dispersion = self.rays, self.dispersion
        base._build_index_maps()
        return base
