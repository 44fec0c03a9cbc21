"""torch.library custom ops over the C ABI, and the autograd Functions that pair each
forward kernel with its hand-written adjoint kernel.

Layering:  scene / element objects  ->  ``trace_sequential`` / ``trace_nonsequential`` /
``element_step`` (this file, autograd.Function)  ->  ``torch.ops.rtt_b200.*`` (opaque custom
ops, registered below)  ->  ctypes  ->  ``librtt_b200.so`` (CUDA, sm_100a).

There is no CPU implementation: every op checks that its tensors live on a CUDA device and
raises otherwise; a missing shared library raises ``RttLibraryMissing`` at first use.
"""
from __future__ import annotations

import ctypes as ct
from typing import List, Optional, Sequence, Tuple

import torch

from . import _cabi, codes as C
from .table import SurfaceTable, compile_elements

MODE_FAST, MODE_EXACT = _cabi.MODE_FAST, _cabi.MODE_EXACT
# Arithmetic defaults.  Sequential traces: FAST (FMA contraction; masks verified identical to
# the reference on every fixture, points within 1e-5).  Non-sequential traces: EXACT — the
# reference's t > 1e-6 self-intersection rule sits at the fp32 ulp of scene-scale coordinates,
# so which surface a ray hits next depends on the reference's exact rounding sequence
# (SURVEY 0.10); only arithmetic that rounds like the reference reproduces its hit sequences.
_default_mode = MODE_FAST
_default_mode_nonseq = MODE_EXACT


def set_default_mode(mode: int, nonseq: Optional[int] = None):
    """Set the arithmetic of the sequential/element ops (and, if given, of the non-sequential op)."""
    global _default_mode, _default_mode_nonseq
    _default_mode = int(mode)
    if nonseq is not None:
        _default_mode_nonseq = int(nonseq)


def get_default_mode() -> int:
    return _default_mode


class NoCpuPathError(RuntimeError):
    pass


def _need_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise NoCpuPathError(
                "rtt_b200 kernels run on CUDA tensors only (sm_100a); this package has no CPU path. "
                "Move the rays and the scene to a CUDA device.")


def _f32c(t: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
    if t is None:
        return None
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def _ptr(t: Optional[torch.Tensor]) -> int:
    return 0 if t is None or t.numel() == 0 else t.data_ptr()


def _stream(t: torch.Tensor):
    return ct.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def _table_req(table_f, table_i, lut, lut_w):
    n_lut = 0 if lut is None or lut.numel() == 0 else lut.shape[0]
    return _cabi.make_table(table_f.data_ptr(), table_i.data_ptr(), table_f.shape[0],
                            _ptr(lut) if n_lut else 0, _ptr(lut_w) if n_lut else 0, n_lut)


SENSOR_CFG = 7   # floats per sensor slot in `sensor_cfg`: H, W, C, x0, y0, sx, sy  (H == 0: no image)


def _sensor_reqs(sensor_cfg: Sequence[float], n: int, records, images, counts=None, record_hits: int = 1):
    reqs = []
    ns = len(sensor_cfg) // SENSOR_CFG
    off = 0
    for s in range(ns):
        H, W, Cn, x0, y0, sx, sy = sensor_cfg[s * SENSOR_CFG:(s + 1) * SENSOR_CFG]
        H, W, Cn = int(H), int(W), int(Cn)
        r = dict(record=(records[s].data_ptr() if records is not None else 0), record_hits=record_hits,
                 count=(counts[s].data_ptr() if counts is not None else 0))
        if H > 0:
            r.update(image=images.data_ptr() + 4 * off, height=H, width=W, channels=Cn, x0=x0, y0=y0, sx=sx, sy=sy)
            off += H * W * Cn
        reqs.append(r)
    return _cabi.make_sensors(reqs)


def _image_numel(sensor_cfg: Sequence[float]) -> int:
    tot = 0
    for s in range(len(sensor_cfg) // SENSOR_CFG):
        H, W, Cn = (int(v) for v in sensor_cfg[s * SENSOR_CFG:s * SENSOR_CFG + 3])
        tot += H * W * Cn
    return tot


# =============================================================================================
# custom ops (opaque to autograd; the Functions below wire the adjoints)
# =============================================================================================
@torch.library.custom_op("rtt_b200::trace_seq_fwd", mutates_args=())
def _trace_seq_fwd(pos: torch.Tensor, dir: torch.Tensor, intensity: torch.Tensor,
                   wavelength: Optional[torch.Tensor], table_f: torch.Tensor, table_i: torch.Tensor,
                   lut: Optional[torch.Tensor], lut_w: Optional[torch.Tensor],
                   sensor_cfg: List[float], want_record: bool, mode: int) -> List[torch.Tensor]:
    """-> [out_pos, out_dir, out_intensity, hitmask(int64), records [ns,N,4], images (flat)]"""
    _need_cuda(pos, dir, intensity, table_f, table_i)
    lib = _cabi.load()
    n = pos.shape[0]
    ns = len(sensor_cfg) // SENSOR_CFG
    opos, odir, oint = torch.empty_like(pos), torch.empty_like(dir), torch.empty_like(intensity)
    hitmask = torch.empty(n, dtype=torch.int64, device=pos.device)
    records = torch.zeros((ns, n, 4), dtype=torch.float32, device=pos.device) if (want_record and ns) \
        else torch.empty((0, n, 4), dtype=torch.float32, device=pos.device)
    images = torch.zeros(_image_numel(sensor_cfg), dtype=torch.float32, device=pos.device)
    sens, cnt = _sensor_reqs(sensor_cfg, n, records if (want_record and ns) else None, images)
    req = _table_req(table_f, table_i, lut, lut_w)
    with torch.cuda.device(pos.device):
        lib.call("rtt_trace_seq_fwd", pos.data_ptr(), dir.data_ptr(), intensity.data_ptr(), _ptr(wavelength),
                 opos.data_ptr(), odir.data_ptr(), oint.data_ptr(), hitmask.data_ptr(),
                 ct.byref(req), sens, cnt, n, mode, _stream(pos))
    return [opos, odir, oint, hitmask, records, images]


@_trace_seq_fwd.register_fake
def _(pos, dir, intensity, wavelength, table_f, table_i, lut, lut_w, sensor_cfg, want_record, mode):
    n = pos.shape[0]
    ns = len(sensor_cfg) // SENSOR_CFG
    return [torch.empty_like(pos), torch.empty_like(dir), torch.empty_like(intensity),
            pos.new_empty(n, dtype=torch.int64), pos.new_empty(((ns if want_record else 0), n, 4)),
            pos.new_empty(_image_numel(sensor_cfg))]


@torch.library.custom_op("rtt_b200::trace_seq_bwd", mutates_args=())
def _trace_seq_bwd(pos: torch.Tensor, dir: torch.Tensor, intensity: torch.Tensor,
                   wavelength: Optional[torch.Tensor], hitmask: torch.Tensor,
                   g_pos: Optional[torch.Tensor], g_dir: Optional[torch.Tensor], g_int: Optional[torch.Tensor],
                   g_records: Optional[torch.Tensor],
                   table_f: torch.Tensor, table_i: torch.Tensor,
                   lut: Optional[torch.Tensor], lut_w: Optional[torch.Tensor],
                   need_rays: bool, need_table: bool, mode: int) -> List[torch.Tensor]:
    """-> [g_in_pos, g_in_dir, g_in_intensity, g_table [S,ROW_G], g_lut [L,S,2]]"""
    _need_cuda(pos, dir, intensity, table_f, table_i, hitmask)
    lib = _cabi.load()
    n, S = pos.shape[0], table_f.shape[0]
    dev = pos.device
    gp = torch.empty_like(pos) if need_rays else pos.new_empty(0)
    gd = torch.empty_like(dir) if need_rays else pos.new_empty(0)
    gi = torch.empty_like(intensity) if need_rays else pos.new_empty(0)
    gt = torch.zeros((S, C.ROW_G), dtype=torch.float32, device=dev) if need_table else pos.new_empty(0)
    has_lut = lut is not None and lut.numel() > 0
    gl = torch.zeros_like(lut) if (need_table and has_lut) else pos.new_empty(0)
    ns = 0 if g_records is None else g_records.shape[0]
    rec_arr = None
    if ns:
        rec_arr = (ct.c_void_p * ns)(*[g_records[s].data_ptr() for s in range(ns)])
    req = _table_req(table_f, table_i, lut, lut_w)
    with torch.cuda.device(dev):
        lib.call("rtt_trace_seq_bwd", pos.data_ptr(), dir.data_ptr(), intensity.data_ptr(), _ptr(wavelength),
                 hitmask.data_ptr(), _ptr(g_pos), _ptr(g_dir), _ptr(g_int), rec_arr,
                 _ptr(gp), _ptr(gd), _ptr(gi), _ptr(gt), _ptr(gl), ct.byref(req), ns, n, mode, _stream(pos))
    return [gp, gd, gi, gt, gl]


@_trace_seq_bwd.register_fake
def _(pos, dir, intensity, wavelength, hitmask, g_pos, g_dir, g_int, g_records, table_f, table_i, lut, lut_w,
      need_rays, need_table, mode):
    e = pos.new_empty(0)
    return [torch.empty_like(pos) if need_rays else e, torch.empty_like(dir) if need_rays else e,
            torch.empty_like(intensity) if need_rays else e,
            pos.new_empty((table_f.shape[0], C.ROW_G)) if need_table else e,
            torch.empty_like(lut) if (need_table and lut is not None) else e]


@torch.library.custom_op("rtt_b200::trace_nonseq_fwd", mutates_args=())
def _trace_nonseq_fwd(pos: torch.Tensor, dir: torch.Tensor, intensity: torch.Tensor,
                      wavelength: Optional[torch.Tensor], table_f: torch.Tensor, table_i: torch.Tensor,
                      lut: Optional[torch.Tensor], lut_w: Optional[torch.Tensor],
                      sensor_cfg: List[float], want_record: bool, nbounces: int, mode: int,
                      record_depth: int = 1) -> List[torch.Tensor]:
    """-> [out_pos, out_dir, out_intensity, hit_seq uint8 [N,B], n_hits uint8 [N], records [ns,K,N,4],
    images (flat), counts uint8 [ns,N]]; K = record_depth = sensor interactions kept per ray."""
    _need_cuda(pos, dir, intensity, table_f, table_i)
    lib = _cabi.load()
    n = pos.shape[0]
    ns = len(sensor_cfg) // SENSOR_CFG
    opos, odir, oint = torch.empty_like(pos), torch.empty_like(dir), torch.empty_like(intensity)
    seq = torch.empty((n, nbounces), dtype=torch.uint8, device=pos.device)
    nh = torch.empty(n, dtype=torch.uint8, device=pos.device)
    K = max(1, int(record_depth))
    rec_on = bool(want_record and ns)
    records = torch.zeros((ns, K, n, 4), dtype=torch.float32, device=pos.device) if rec_on \
        else torch.empty((0, K, n, 4), dtype=torch.float32, device=pos.device)
    counts = torch.zeros((ns if rec_on else 0, n), dtype=torch.uint8, device=pos.device)
    images = torch.zeros(_image_numel(sensor_cfg), dtype=torch.float32, device=pos.device)
    sens, cnt = _sensor_reqs(sensor_cfg, n, records if rec_on else None, images, counts if rec_on else None, K)
    req = _table_req(table_f, table_i, lut, lut_w)
    with torch.cuda.device(pos.device):
        lib.call("rtt_trace_nonseq_fwd", pos.data_ptr(), dir.data_ptr(), intensity.data_ptr(), _ptr(wavelength),
                 opos.data_ptr(), odir.data_ptr(), oint.data_ptr(), seq.data_ptr(), nh.data_ptr(),
                 ct.byref(req), sens, cnt, nbounces, n, mode, _stream(pos))
    return [opos, odir, oint, seq, nh, records, images, counts]


@_trace_nonseq_fwd.register_fake
def _(pos, dir, intensity, wavelength, table_f, table_i, lut, lut_w, sensor_cfg, want_record, nbounces, mode,
      record_depth=1):
    n = pos.shape[0]
    ns = len(sensor_cfg) // SENSOR_CFG
    K = max(1, int(record_depth))
    return [torch.empty_like(pos), torch.empty_like(dir), torch.empty_like(intensity),
            pos.new_empty((n, nbounces), dtype=torch.uint8), pos.new_empty(n, dtype=torch.uint8),
            pos.new_empty(((ns if want_record else 0), K, n, 4)), pos.new_empty(_image_numel(sensor_cfg)),
            pos.new_empty(((ns if want_record else 0), n), dtype=torch.uint8)]


@torch.library.custom_op("rtt_b200::trace_nonseq_bwd", mutates_args=())
def _trace_nonseq_bwd(pos: torch.Tensor, dir: torch.Tensor, intensity: torch.Tensor,
                      wavelength: Optional[torch.Tensor], hit_seq: torch.Tensor,
                      g_pos: Optional[torch.Tensor], g_dir: Optional[torch.Tensor], g_int: Optional[torch.Tensor],
                      g_records: Optional[torch.Tensor],
                      table_f: torch.Tensor, table_i: torch.Tensor,
                      lut: Optional[torch.Tensor], lut_w: Optional[torch.Tensor],
                      need_rays: bool, need_table: bool, mode: int) -> List[torch.Tensor]:
    _need_cuda(pos, dir, intensity, table_f, table_i, hit_seq)
    lib = _cabi.load()
    n, S = pos.shape[0], table_f.shape[0]
    dev = pos.device
    gp = torch.empty_like(pos) if need_rays else pos.new_empty(0)
    gd = torch.empty_like(dir) if need_rays else pos.new_empty(0)
    gi = torch.empty_like(intensity) if need_rays else pos.new_empty(0)
    gt = torch.zeros((S, C.ROW_G), dtype=torch.float32, device=dev) if need_table else pos.new_empty(0)
    has_lut = lut is not None and lut.numel() > 0
    gl = torch.zeros_like(lut) if (need_table and has_lut) else pos.new_empty(0)
    ns = 0 if g_records is None else g_records.shape[0]          # g_records: [ns, K, N, 4]
    rec_arr = (ct.c_void_p * ns)(*[g_records[s].data_ptr() for s in range(ns)]) if ns else None
    depth = (ct.c_int32 * ns)(*([g_records.shape[1]] * ns)) if ns else None
    req = _table_req(table_f, table_i, lut, lut_w)
    with torch.cuda.device(dev):
        lib.call("rtt_trace_nonseq_bwd", pos.data_ptr(), dir.data_ptr(), intensity.data_ptr(), _ptr(wavelength),
                 hit_seq.data_ptr(), hit_seq.shape[1], _ptr(g_pos), _ptr(g_dir), _ptr(g_int), rec_arr, depth,
                 _ptr(gp), _ptr(gd), _ptr(gi), _ptr(gt), _ptr(gl), ct.byref(req), ns, n, mode, _stream(pos))
    return [gp, gd, gi, gt, gl]


@_trace_nonseq_bwd.register_fake
def _(pos, dir, intensity, wavelength, hit_seq, g_pos, g_dir, g_int, g_records, table_f, table_i, lut, lut_w,
      need_rays, need_table, mode):
    e = pos.new_empty(0)
    return [torch.empty_like(pos) if need_rays else e, torch.empty_like(dir) if need_rays else e,
            torch.empty_like(intensity) if need_rays else e,
            pos.new_empty((table_f.shape[0], C.ROW_G)) if need_table else e,
            torch.empty_like(lut) if (need_table and lut is not None) else e]


@torch.library.custom_op("rtt_b200::intersect_test", mutates_args=())
def _intersect_test(pos: torch.Tensor, dir: torch.Tensor, table_f: torch.Tensor, table_i: torch.Tensor,
                    row0: int, k: int, mode: int) -> torch.Tensor:
    _need_cuda(pos, dir, table_f, table_i)
    lib = _cabi.load()
    n = pos.shape[0]
    t = torch.empty((n, k), dtype=torch.float32, device=pos.device)
    req = _table_req(table_f, table_i, None, None)
    with torch.cuda.device(pos.device):
        lib.call("rtt_intersect_test", pos.data_ptr(), dir.data_ptr(), t.data_ptr(), ct.byref(req), row0, k, n,
                 mode, _stream(pos))
    return t


@_intersect_test.register_fake
def _(pos, dir, table_f, table_i, row0, k, mode):
    return pos.new_empty((pos.shape[0], k))


@torch.library.custom_op("rtt_b200::surface_step_fwd", mutates_args=())
def _surface_step_fwd(pos: torch.Tensor, dir: torch.Tensor, wavelength: Optional[torch.Tensor],
                      table_f: torch.Tensor, table_i: torch.Tensor,
                      lut: Optional[torch.Tensor], lut_w: Optional[torch.Tensor],
                      row: int, mode: int) -> List[torch.Tensor]:
    """-> [new_pos, new_dir, mod, hit_local, t, normal]"""
    _need_cuda(pos, dir, table_f, table_i)
    lib = _cabi.load()
    n = pos.shape[0]
    npos, ndir, hl, nrm = (torch.empty_like(pos) for _ in range(4))
    mod = torch.empty(n, dtype=torch.float32, device=pos.device)
    t = torch.empty(n, dtype=torch.float32, device=pos.device)
    req = _table_req(table_f, table_i, lut, lut_w)
    with torch.cuda.device(pos.device):
        lib.call("rtt_surface_step_fwd", pos.data_ptr(), dir.data_ptr(), _ptr(wavelength),
                 npos.data_ptr(), ndir.data_ptr(), mod.data_ptr(), hl.data_ptr(), t.data_ptr(), nrm.data_ptr(),
                 ct.byref(req), row, n, mode, _stream(pos))
    return [npos, ndir, mod, hl, t, nrm]


@_surface_step_fwd.register_fake
def _(pos, dir, wavelength, table_f, table_i, lut, lut_w, row, mode):
    n = pos.shape[0]
    return [torch.empty_like(pos), torch.empty_like(pos), pos.new_empty(n), torch.empty_like(pos),
            pos.new_empty(n), torch.empty_like(pos)]


@torch.library.custom_op("rtt_b200::surface_step_bwd", mutates_args=())
def _surface_step_bwd(pos: torch.Tensor, dir: torch.Tensor, wavelength: Optional[torch.Tensor],
                      g_npos: Optional[torch.Tensor], g_ndir: Optional[torch.Tensor],
                      g_hl: Optional[torch.Tensor], g_t: Optional[torch.Tensor], g_n: Optional[torch.Tensor],
                      table_f: torch.Tensor, table_i: torch.Tensor,
                      lut: Optional[torch.Tensor], lut_w: Optional[torch.Tensor],
                      row: int, mode: int) -> List[torch.Tensor]:
    """-> [g_pos, g_dir, g_table, g_lut]"""
    _need_cuda(pos, dir, table_f, table_i)
    lib = _cabi.load()
    n, S = pos.shape[0], table_f.shape[0]
    gp, gd = torch.empty_like(pos), torch.empty_like(dir)
    gt = torch.zeros((S, C.ROW_G), dtype=torch.float32, device=pos.device)
    has_lut = lut is not None and lut.numel() > 0
    gl = torch.zeros_like(lut) if has_lut else pos.new_empty(0)
    req = _table_req(table_f, table_i, lut, lut_w)
    with torch.cuda.device(pos.device):
        lib.call("rtt_surface_step_bwd", pos.data_ptr(), dir.data_ptr(), _ptr(wavelength),
                 _ptr(g_npos), _ptr(g_ndir), _ptr(g_hl), _ptr(g_t), _ptr(g_n),
                 gp.data_ptr(), gd.data_ptr(), gt.data_ptr(), _ptr(gl), ct.byref(req), row, n, mode, _stream(pos))
    return [gp, gd, gt, gl]


@_surface_step_bwd.register_fake
def _(pos, dir, wavelength, g_npos, g_ndir, g_hl, g_t, g_n, table_f, table_i, lut, lut_w, row, mode):
    return [torch.empty_like(pos), torch.empty_like(dir), pos.new_empty((table_f.shape[0], C.ROW_G)),
            torch.empty_like(lut) if lut is not None else pos.new_empty(0)]


# =============================================================================================
# autograd Functions
# =============================================================================================
def _pad_table_grad(gt: torch.Tensor, like: torch.Tensor) -> torch.Tensor:
    return gt if gt.shape == like.shape else torch.nn.functional.pad(gt, (0, like.shape[1] - gt.shape[1]))


def _cg(g: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
    return None if g is None else g.contiguous()


class _TraceSeq(torch.autograd.Function):
    """forward = rtt_trace_seq_fwd, backward = rtt_trace_seq_bwd (hand-written adjoint)."""

    @staticmethod
    def forward(ctx, pos, dir_, intensity, wavelength, table_f, table_i, lut, lut_w, sensor_cfg, want_record, mode):
        outs = torch.ops.rtt_b200.trace_seq_fwd(pos, dir_, intensity, wavelength, table_f, table_i, lut, lut_w,
                                                sensor_cfg, want_record, mode)
        opos, odir, oint, hitmask, records, images = outs
        ctx.save_for_backward(pos, dir_, intensity, wavelength, hitmask, table_f, table_i, lut, lut_w)
        ctx.mode = mode
        ctx.mark_non_differentiable(hitmask, images)
        return opos, odir, oint, hitmask, records, images

    @staticmethod
    def backward(ctx, g_pos, g_dir, g_int, _g_mask, g_records, _g_images):
        pos, dir_, intensity, wavelength, hitmask, table_f, table_i, lut, lut_w = ctx.saved_tensors
        need_rays = any(ctx.needs_input_grad[:3])
        need_table = ctx.needs_input_grad[4] or ctx.needs_input_grad[6]
        if g_records is not None and g_records.numel() == 0:
            g_records = None
        gp, gd, gi, gt, gl = torch.ops.rtt_b200.trace_seq_bwd(
            pos, dir_, intensity, wavelength, hitmask, _cg(g_pos), _cg(g_dir), _cg(g_int), _cg(g_records),
            table_f, table_i, lut, lut_w, need_rays, need_table, ctx.mode)
        return (gp if ctx.needs_input_grad[0] else None, gd if ctx.needs_input_grad[1] else None,
                gi if ctx.needs_input_grad[2] else None, None,
                _pad_table_grad(gt, table_f) if ctx.needs_input_grad[4] else None, None,
                gl if ctx.needs_input_grad[6] else None, None, None, None, None)


class _TraceNonseq(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pos, dir_, intensity, wavelength, table_f, table_i, lut, lut_w, sensor_cfg, want_record,
                nbounces, mode, record_depth):
        outs = torch.ops.rtt_b200.trace_nonseq_fwd(pos, dir_, intensity, wavelength, table_f, table_i, lut, lut_w,
                                                   sensor_cfg, want_record, nbounces, mode, record_depth)
        opos, odir, oint, seq, nh, records, images, counts = outs
        ctx.save_for_backward(pos, dir_, intensity, wavelength, seq, table_f, table_i, lut, lut_w)
        ctx.mode = mode
        ctx.mark_non_differentiable(seq, nh, images, counts)
        return opos, odir, oint, seq, nh, records, images, counts

    @staticmethod
    def backward(ctx, g_pos, g_dir, g_int, _g_seq, _g_nh, g_records, _g_images, _g_counts):
        pos, dir_, intensity, wavelength, seq, table_f, table_i, lut, lut_w = ctx.saved_tensors
        need_rays = any(ctx.needs_input_grad[:3])
        need_table = ctx.needs_input_grad[4] or ctx.needs_input_grad[6]
        if g_records is not None and g_records.numel() == 0:
            g_records = None
        gp, gd, gi, gt, gl = torch.ops.rtt_b200.trace_nonseq_bwd(
            pos, dir_, intensity, wavelength, seq, _cg(g_pos), _cg(g_dir), _cg(g_int), _cg(g_records),
            table_f, table_i, lut, lut_w, need_rays, need_table, ctx.mode)
        return (gp if ctx.needs_input_grad[0] else None, gd if ctx.needs_input_grad[1] else None,
                gi if ctx.needs_input_grad[2] else None, None,
                _pad_table_grad(gt, table_f) if ctx.needs_input_grad[4] else None, None,
                gl if ctx.needs_input_grad[6] else None, None, None, None, None, None, None)


class _SurfaceStep(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pos, dir_, wavelength, table_f, table_i, lut, lut_w, row, mode):
        npos, ndir, mod, hl, t, nrm = torch.ops.rtt_b200.surface_step_fwd(pos, dir_, wavelength, table_f, table_i,
                                                                          lut, lut_w, row, mode)
        ctx.save_for_backward(pos, dir_, wavelength, table_f, table_i, lut, lut_w)
        ctx.row, ctx.mode = row, mode
        ctx.mark_non_differentiable(mod)
        return npos, ndir, mod, hl, t, nrm

    @staticmethod
    def backward(ctx, g_npos, g_ndir, _g_mod, g_hl, g_t, g_n):
        pos, dir_, wavelength, table_f, table_i, lut, lut_w = ctx.saved_tensors
        gp, gd, gt, gl = torch.ops.rtt_b200.surface_step_bwd(
            pos, dir_, wavelength, _cg(g_npos), _cg(g_ndir), _cg(g_hl), _cg(g_t), _cg(g_n),
            table_f, table_i, lut, lut_w, ctx.row, ctx.mode)
        return (gp if ctx.needs_input_grad[0] else None, gd if ctx.needs_input_grad[1] else None, None,
                _pad_table_grad(gt, table_f) if ctx.needs_input_grad[3] else None, None,
                gl if ctx.needs_input_grad[5] else None, None, None, None)


# =============================================================================================
# public functional API
# =============================================================================================
def sensor_cfg_of(table: SurfaceTable) -> List[float]:
    """Flatten the image requests of the table's sensors (Sensor.set_image) for the op."""
    cfg: List[float] = []
    for el in table.sensors:
        spec = getattr(el, "image_spec", None)
        if spec is None:
            cfg += [0.0] * SENSOR_CFG
        else:
            H, W, x0, x1, y0, y1, ch = spec
            sx = float(torch.tensor(W / (x1 - x0), dtype=torch.float32))
            sy = float(torch.tensor(H / (y1 - y0), dtype=torch.float32))
            cfg += [float(H), float(W), float(ch), float(x0), float(y0), sx, sy]
    return cfg


def split_images(images: torch.Tensor, sensor_cfg: Sequence[float]) -> List[Optional[torch.Tensor]]:
    out, off = [], 0
    for s in range(len(sensor_cfg) // SENSOR_CFG):
        H, W, Cn = (int(v) for v in sensor_cfg[s * SENSOR_CFG:s * SENSOR_CFG + 3])
        if H > 0:
            out.append(images[off:off + H * W * Cn].view(Cn, H, W))
            off += H * W * Cn
        else:
            out.append(None)
    return out


def _prep_rays(pos, dir_, intensity, wavelength, table: SurfaceTable):
    _need_cuda(pos, dir_, intensity, table.f)
    pos, dir_, intensity = _f32c(pos), _f32c(dir_), _f32c(intensity)
    wav = _f32c(wavelength) if table.lut is not None else None
    return pos, dir_, intensity, wav


def trace_sequential(table: SurfaceTable, pos, dir_, intensity, wavelength=None, *, want_record=True,
                     sensor_cfg: Optional[List[float]] = None, mode: Optional[int] = None):
    """Fused SequentialScene.simulate (scene/sequential.py:12-36).

    Returns dict(pos, dir, intensity, hitmask [N] int64 (bit r = interacted with row r),
    records [n_sensors,N,4] (hit_local xyz, weight-before), images [per sensor: [C,H,W] or None])."""
    pos, dir_, intensity, wav = _prep_rays(pos, dir_, intensity, wavelength, table)
    cfg = sensor_cfg_of(table) if sensor_cfg is None else list(sensor_cfg)
    mode = _default_mode if mode is None else mode
    opos, odir, oint, hitmask, records, images = _TraceSeq.apply(
        pos, dir_, intensity, wav, table.f, table.i, table.lut, table.lut_wavelengths, cfg, bool(want_record), mode)
    return dict(pos=opos, dir=odir, intensity=oint, hitmask=hitmask, records=records,
                images=split_images(images, cfg))


def trace_nonsequential(table: SurfaceTable, pos, dir_, intensity, nbounces: int, wavelength=None, *,
                        want_record=True, sensor_cfg: Optional[List[float]] = None, mode: Optional[int] = None,
                        record_depth: int = 1):
    """Fused Scene.simulate bounce loop (scene/base.py:129-235).

    Returns dict(pos, dir, intensity, hit_seq [N,B] uint8 (255 = none), n_hits [N] uint8,
    records [n_sensors, K, N, 4] (the k-th interaction of ray i with the sensor, K = record_depth),
    sensor_counts [n_sensors, N] uint8 (interactions per ray, may exceed K), images)."""
    pos, dir_, intensity, wav = _prep_rays(pos, dir_, intensity, wavelength, table)
    if not 0 <= nbounces <= C.MAX_BOUNCES:
        raise ValueError(f"nbounces must be in [0, {C.MAX_BOUNCES}]")
    cfg = sensor_cfg_of(table) if sensor_cfg is None else list(sensor_cfg)
    mode = _default_mode_nonseq if mode is None else mode
    opos, odir, oint, seq, nh, records, images, counts = _TraceNonseq.apply(
        pos, dir_, intensity, wav, table.f, table.i, table.lut, table.lut_wavelengths, cfg, bool(want_record),
        int(nbounces), mode, max(1, int(record_depth)))
    return dict(pos=opos, dir=odir, intensity=oint, hit_seq=seq, n_hits=nh, records=records,
                sensor_counts=counts, images=split_images(images, cfg))


def intersect_rows(table: SurfaceTable, pos, dir_, row0: int, k: int, mode: Optional[int] = None) -> torch.Tensor:
    _need_cuda(pos, dir_, table.f)
    mode = _default_mode if mode is None else mode
    return torch.ops.rtt_b200.intersect_test(_f32c(pos).detach(), _f32c(dir_).detach(), table.f.detach(), table.i,
                                             int(row0), int(k), mode)


def step_row(table: SurfaceTable, pos, dir_, row: int, wavelength=None, mode: Optional[int] = None):
    """(new_pos, new_dir, mod, hit_local, t, normal) of one row, differentiable."""
    _need_cuda(pos, dir_, table.f)
    mode = _default_mode if mode is None else mode
    wav = _f32c(wavelength) if table.lut is not None else None
    return _SurfaceStep.apply(_f32c(pos), _f32c(dir_), wav, table.f, table.i, table.lut, table.lut_wavelengths,
                              int(row), mode)


# ---- object-level helpers used by geom.py / elements.py / phys.py --------------------------
class _Holder:
    """Minimal element wrapper so a bare Surface / Shape can be compiled on its own."""

    def __init__(self, shape, surface_functions):
        self.shape, self.surface_functions = shape, surface_functions


def _single_table(shape_or_element) -> SurfaceTable:
    from . import phys as P
    obj = shape_or_element
    if hasattr(obj, "surface_functions") and hasattr(obj, "shape"):
        return compile_elements([obj])
    return compile_elements([_Holder(obj, [P.Transmit() for _ in range(len(obj))])])


def element_step(element, rays, surf_idx: int):
    """Element.forward: (new_pos, new_dir, intensity_mult, hit_local)."""
    tab = _single_table(element)
    npos, ndir, mod, hl, _t, _n = step_row(tab, rays.pos, rays.dir, surf_idx, getattr(rays, "wavelength", None))
    return npos, ndir, mod, hl


def surface_geometry(shape, rays, surf_idx: int):
    """Shape.forward / Surface.forward: (t, hit_global, normal_global, hit_local)."""
    tab = _single_table(shape)
    npos, _ndir, _mod, hl, t, nrm = step_row(tab, rays.pos, rays.dir, surf_idx)
    return t, npos, nrm, hl


def shape_intersect_test(shape, rays):
    tab = _single_table(shape)
    return intersect_rows(tab, rays.pos, rays.dir, 0, tab.n_rows)


def surface_intersect_test(surface, rays):
    tab = _single_table(surface)
    return intersect_rows(tab, rays.pos, rays.dir, 0, 1)


def surface_in_bounds(surface, local_pos):
    """SurfaceBounded.inBounds on points already in the surface frame (boolean mask).

    Evaluated by the aperture-filter physics of a pose-free copy of the bound."""
    raise NotImplementedError(
        "inBounds on raw points is evaluated inside the kernels (root selection, ApertureFilter); "
        "call Element.forward / intersectTest instead")


def physics_apply(surface_function, local_intersect, ray_dir, normal):
    raise NotImplementedError(
        "surface functions are evaluated inside the fused kernels; call Element.forward(rays, surf_idx)")
