"""Goals that drive forward + adjoint for lens optimisation (mirror of ``optim/goals.py``).

``SpotTargetLoss`` (optim/goals.py:42-96) and ``SpotSizeLoss`` (:99-187) keep the reference's
constructor signatures and formulas.  They call ``scene.simulate()`` with ``scene.rays`` set, which
works for both scene classes here (the reference's own call raises for ``SequentialScene``,
SURVEY section 0.7).  Differences are in execution only:

* the fused kernels hand over one sensor record per ray (zero weight = no hit), and the goals reduce
  those records with dedicated kernels (``ops.spot_moments`` / ``ops.spot_size``: two reduction launches
  forward, one elementwise launch backward) — no boolean gather, no host synchronisation, no eager
  elementwise chain; bundles sampled on the device are generated inside the trace kernels and the trace
  skips its final-ray outputs, so a goal evaluation moves 24 B/ray forward and 40 B/ray backward;
* with ``torch.distributed`` initialised the moments (W, sum w x, sum w y) and the per-bundle sums
  are all-reduced, so a bundle sharded over ranks gives the same loss on every rank; after
  ``loss.backward()`` one ``dist.allreduce_scene_results(sensors, params)`` sums the per-shard
  parameter gradients into the full gradient.

``FocalLengthLoss`` works on the paraxial 5x5 matrices (no rays) and is outside this package's
scope.
"""
from __future__ import annotations

from typing import List, Optional

import torch
import torch.nn as nn

from . import ops
from .elements import Sensor
from .rays import Bundle


class Goal(nn.Module):
    """Base class for optimisation goals (optim/goals.py:11-13)."""


class _SumOverRanks(torch.autograd.Function):
    """all_reduce(SUM) whose backward passes the cotangent through unchanged.

    Every rank evaluates the SAME loss from the reduced moments, so the cotangent is identical on
    all ranks; with an identity backward each rank's parameter gradient is its shard's partial
    derivative, and the one SUM all-reduce of the gradients (dist.allreduce_scene_results) yields
    the full gradient — no 1/world_size bookkeeping."""

    @staticmethod
    def forward(ctx, t):
        import torch.distributed as dist
        out = t.clone()
        dist.all_reduce(out, op=dist.ReduceOp.SUM)
        return out

    @staticmethod
    def backward(ctx, g):
        return g


def _dist_sum(t: torch.Tensor) -> torch.Tensor:
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        return _SumOverRanks.apply(t)
    return t


def _sensor_hits(scene, sensor: Sensor):
    """(xy [M,2], w [M]) with zero weight on rays that did not reach the sensor.

    Fast path: the raw per-ray record of the latest fused trace (no compaction).  Fallback: the
    reference's semantics on the hit lists / final rays (optim/goals.py:76-82)."""
    tr = getattr(scene, "last_trace", None)
    if tr is not None and tr.get("records") is not None and tr["records"].numel():
        table = getattr(scene, "_last_table", None) or scene.table()
        if sensor in table.sensors:
            slot = table.sensors.index(sensor)
            rec = tr["records"][slot]
            if "hitmask" in tr:                             # sequential: [N,4], one interaction per ray
                hit = ((tr["hitmask"] >> table.sensor_rows[slot]) & 1).bool()
            else:                                           # non-sequential: [K,N,4], every kept interaction counts
                cnt = tr["sensor_counts"][slot]
                hit = (torch.arange(rec.shape[0], device=cnt.device)[:, None] < cnt[None, :]).reshape(-1)
                rec = rec.reshape(-1, 4)
            return rec[:, :2], torch.where(hit, rec[:, 3], torch.zeros_like(rec[:, 3]))
    if sensor.hitLocs:
        locs, w, _ = sensor.getHitsTensors()
        return locs[:, :2], w
    return scene.rays.pos[:, :2], scene.rays.intensity


def _sensor_records(scene, sensor: Sensor):
    """Raw records [M,4] of `sensor` from the latest fused trace (M = rays, or depth x rays for the
    non-sequential trace; entries that were not hit are all zero), or None."""
    tr = getattr(scene, "last_trace", None)
    if tr is None or tr.get("records") is None or not tr["records"].numel():
        return None
    table = getattr(scene, "_last_table", None) or scene.table()
    if sensor not in table.sensors:
        return None
    recs = tr["records"]
    if recs.shape[0] == 1:          # one sensor: a view whose backward is a view too (select's backward would
        return recs.reshape(-1, 4)  # allocate and fill a zero tensor of the records' size first)
    return recs[table.sensors.index(sensor)].reshape(-1, 4)


class _goal_trace:
    """While a goal evaluates: keep sensor records, skip the final-ray outputs of generated bundles."""

    def __init__(self, scene):
        self.scene = scene

    def __enter__(self):
        sc = self.scene
        self.saved = (getattr(sc, "record_hits", None), getattr(sc, "final_rays", None))
        if self.saved[0] is not None:
            sc.record_hits = True
        if self.saved[1] is not None:
            sc.final_rays = False

    def __exit__(self, *exc):
        sc = self.scene
        if self.saved[0] is not None:
            sc.record_hits = self.saved[0]
        if self.saved[1] is not None:
            sc.final_rays = self.saved[1]


def _place(scene, rays):
    if hasattr(scene, "parameters"):
        dev = next(iter(scene.parameters()), torch.zeros(1)).device
        rays = rays.to(dev)
    scene.rays = rays
    return rays


def _mean_over_active(losses, flags):
    """Mean of the per-bundle terms over the bundles that reached the sensor: the reference `continue`s on a bundle
    without (active) hits and averages the rest (optim/goals.py:72-74, 165-167, 185-187).  On the device path the
    emptiness test is a device flag per bundle (no host synchronisation): sum(term * flag) / max(sum(flag), 1), which
    is the plain mean when every bundle is active and 0 when none is.  The eager path (`flags` empty) keeps the mean."""
    if len(flags) != len(losses):
        return torch.stack(losses).mean()
    f = torch.stack(flags)
    return (torch.stack(losses) * f).sum() / f.sum().clamp(min=1.0)


class SpotTargetLoss(Goal):
    """Squared distance between each bundle's intensity centroid and a target (optim/goals.py:42-96)."""

    def __init__(self, sensor: Sensor, target_xy: torch.Tensor):
        super().__init__()
        self.sensor = sensor
        target_xy = torch.as_tensor(target_xy, dtype=torch.float32)
        if target_xy.ndim == 1:
            target_xy = target_xy.unsqueeze(0)
        self.register_buffer("target_xy", target_xy)

    def forward(self, scene, bundles: List[Bundle], N_rays: int = 128) -> torch.Tensor:
        losses, flags = [], []
        for i, bundle in enumerate(bundles):
            self.sensor.reset()
            _place(scene, bundle.sample(N_rays))
            with _goal_trace(scene):
                scene.simulate()
            rec = _sensor_records(scene, self.sensor)
            if rec is not None and rec.is_cuda:
                # never skipped, even on an empty local shard: spot_moments holds a collective every rank must enter
                mom = ops.spot_moments(rec, active_only=False)       # every recorded hit (optim/goals.py:76-88)
                xy = rec
                flags.append((mom[3] > 0).to(torch.float32).detach())   # bundles that never reached the sensor drop
            else:
                xy, w = _sensor_hits(scene, self.sensor)
                if w.shape[0] == 0:
                    continue
                mom = _dist_sum(torch.stack([w.sum(), (xy[:, 0] * w).sum(), (xy[:, 1] * w).sum()]))
            w_sum = mom[0].clamp(min=1e-12)
            cx, cy = mom[1] / w_sum, mom[2] / w_sum
            tidx = min(i, self.target_xy.shape[0] - 1)
            tx, ty = self.target_xy[tidx, 0].to(xy.device), self.target_xy[tidx, 1].to(xy.device)
            losses.append((cx - tx) ** 2 + (cy - ty) ** 2)
        if not losses:
            return torch.tensor(0.0)
        return _mean_over_active(losses, flags)


class SpotSizeLoss(Goal):
    """Mean over bundles of  sum_i sqrt(((x_i-cx)^2 + (y_i-cy)^2) * w_i / W)  (optim/goals.py:99-187;
    note: a sum of square roots, exactly as the reference writes it)."""

    def __init__(self, sensor: Sensor, bundles: List[Bundle], N_rays: int = 128,
                 target_xy: Optional[torch.Tensor] = None):
        super().__init__()
        self.sensor, self.bundles, self.N_rays = sensor, bundles, N_rays
        if target_xy is not None:
            target_xy = torch.as_tensor(target_xy, dtype=torch.float32)
            if target_xy.ndim == 1:
                target_xy = target_xy.unsqueeze(0).expand(len(bundles), 2)
        self._target_xy = target_xy

    def forward(self, scene) -> torch.Tensor:
        losses, flags = [], []
        for i, bundle in enumerate(self.bundles):
            self.sensor.reset()
            _place(scene, bundle.sample(self.N_rays))
            with _goal_trace(scene):
                scene.simulate()
            rec = _sensor_records(scene, self.sensor)
            if rec is not None and rec.is_cuda:             # fused reductions (rtt_goals.cu)
                tgt = None
                if self._target_xy is not None:
                    if self._target_xy.device != rec.device:   # once: no per-step host-to-device copy
                        self._target_xy = self._target_xy.to(rec.device)
                    tgt = self._target_xy[min(i, self._target_xy.shape[0] - 1)]
                term, active = ops.spot_size_active(rec, tgt)
                losses.append(term)
                flags.append(active)
                continue
            xy, w = _sensor_hits(scene, self.sensor)
            active = w > 0                                  # optim/goals.py:165 (as a mask: no gather)
            wa = torch.where(active, w, torch.zeros_like(w))
            mom = _dist_sum(torch.stack([wa.sum(), (xy[:, 0] * wa).sum(), (xy[:, 1] * wa).sum()]))
            w_sum = mom[0].clamp(min=1e-12)
            if self._target_xy is not None:
                tidx = min(i, self._target_xy.shape[0] - 1)
                cx, cy = self._target_xy[tidx, 0].to(xy.device), self._target_xy[tidx, 1].to(xy.device)
            else:
                cx, cy = mom[1] / w_sum, mom[2] / w_sum
            r2w = ((xy[:, 0] - cx) ** 2 + (xy[:, 1] - cy) ** 2) * (wa / w_sum)
            # sqrt only where active: d sqrt(0) would poison the gradients of the masked rays
            rms = torch.sqrt(torch.where(active, r2w, torch.ones_like(r2w)))
            losses.append(_dist_sum(torch.where(active, rms, torch.zeros_like(rms)).sum()))
        if not losses:
            return torch.tensor(0.0)
        return _mean_over_active(losses, flags)


class GraphedStep:
    """One optimisation step — ``zero_grad -> goal(scene) -> backward -> optimizer.step`` — captured in a CUDA graph.

    A step of the singlet optimisation is ~25 launches (two trace kernels, three goal reductions, the table-building
    elementwise ops, the optimiser) for ~1.3 ms of GPU work at 1e7 rays; replaying one graph removes the launch gaps.
    Device-generated bundles advance their {key, counter} state on the device, so every replay traces fresh rays.

    Requirements: CUDA scene and bundles, a capturable optimiser (e.g. ``torch.optim.Adam(..., capturable=True)``), no
    trainable rotation vectors (``matrix_exp`` synchronises) and no stochastic rows (their seed is drawn on the host).
    ``GraphedStep.try_build`` returns None instead of raising when the capture is refused (reason in ``last_error``)."""

    last_error = None

    def __init__(self, scene, goal, optimizer, warmup: int = 3, after_backward=None):
        """``after_backward``: optional callable run between ``backward`` and ``optimizer.step`` — with a bundle sharded
        over ranks, the one all-reduce of the parameter gradients (``dist.allreduce_scene_results``); NCCL collectives
        are captured into the graph like kernels."""
        self.scene, self.goal, self.optimizer = scene, goal, optimizer
        dev = next(p for p in scene.parameters()).device

        def drop_graph_refs():
            # Autograd keeps ONE AccumulateGrad node per Parameter alive as long as any graph references it, and the
            # node remembers the stream it was created on.  A table cached by the scene (or a loss kept by the caller)
            # from an earlier eager call would pin nodes of the default stream, which invalidates the capture.
            comp = getattr(scene, "_compiler", None)
            if comp is not None:
                comp._table = None
            scene._last_table = None
            scene.last_trace = None
            for el in getattr(scene, "elements", []):
                if hasattr(el, "reset"):
                    el.reset()
            optimizer.zero_grad(set_to_none=True)

        drop_graph_refs()
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):                       # eager warm-up: allocator, caches, optimiser state
            for _ in range(max(1, warmup)):
                optimizer.zero_grad(set_to_none=True)
                loss = goal(scene)
                loss.backward()
                if after_backward is not None:
                    after_backward()
                optimizer.step()
                del loss
                drop_graph_refs()
        torch.cuda.current_stream(dev).wait_stream(side)
        from . import _cabi
        lib = _cabi.load()
        self.graph = torch.cuda.CUDAGraph()
        optimizer.zero_grad(set_to_none=True)
        before = lib.launch_count()
        with torch.cuda.graph(self.graph):
            self.loss = goal(scene)
            self.loss.backward()
            if after_backward is not None:
                after_backward()
            optimizer.step()
        self.launches_per_step = lib.launch_count() - before     # this library's kernels inside one replay

    def __call__(self) -> torch.Tensor:
        """Run one step; returns the (static) loss tensor of that step."""
        self.graph.replay()
        return self.loss

    @classmethod
    def try_build(cls, scene, goal, optimizer, warmup: int = 3, after_backward=None):
        cls.last_error = None
        try:
            return cls(scene, goal, optimizer, warmup, after_backward)
        except Exception as exc:                             # capture refused: callers fall back to eager steps
            cls.last_error = f"{type(exc).__name__}: {exc}"
            torch.cuda.synchronize()
            return None
