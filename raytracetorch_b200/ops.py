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
MODE_SCALAR_GRADS = _cabi.MODE_SCALAR_GRADS
MODE_NONSEQ_FAST = _cabi.MODE_NONSEQ_FAST
MODE_NO_FINAL_RAYS = 0x200    # trace_seq_fwd op only: skip the final pos / dir / intensity outputs (goal evaluations)
# Arithmetic defaults.  Sequential traces: FAST (FMA contraction; masks verified identical to
# the reference on every fixture, points within 1e-5).  Non-sequential traces: EXACT — the
# reference's t > 1e-6 self-intersection rule sits at the fp32 ulp of scene-scale coordinates,
# so which surface a ray hits next depends on the reference's exact rounding sequence
# (SURVEY 0.10); only arithmetic that rounds like the reference reproduces its hit sequences.
_default_mode = MODE_FAST
_default_mode_nonseq = MODE_EXACT

# Kernel-build selectors (include/rtt_b200.h RTT_MODE_TUNE_*): 0 = the library's default for the table.  The C library
# itself reads no environment; for sweeps the Python layer honours RTT_FWD_TILE / RTT_BWD_MINB once at import
# (RTT_FWD_TILE=0, the per-ray kernel, is selector 9).
MODE_TUNE_SHIFT, MODE_TUNE_MASK = 16, 0xFF0000


def _env_tune(name: str, zero_means: int = 0) -> int:
    import os
    v = os.environ.get(name)
    if v is None or not v.strip().lstrip("-").isdigit():
        return 0
    v = int(v)
    return zero_means if v == 0 else max(0, min(255, v))


_tune_fwd = _env_tune("RTT_FWD_TILE", zero_means=9)
_tune_bwd = _env_tune("RTT_BWD_MINB")
_tune_nonseq = _env_tune("RTT_NS_TUNE")


def set_tuning(fwd: Optional[int] = None, bwd: Optional[int] = None):
    """Select the compiled build of the sequential forward / adjoint kernel (0 = default); see RTT_MODE_TUNE_* in
    include/rtt_b200.h.  Results do not depend on it."""
    global _tune_fwd, _tune_bwd
    if fwd is not None:
        _tune_fwd = int(fwd) & 0xFF
    if bwd is not None:
        _tune_bwd = int(bwd) & 0xFF


def _with_tune(mode: int, tune: int) -> int:
    return mode if (mode & MODE_TUNE_MASK) else (mode | (tune << MODE_TUNE_SHIFT))


def _fwd_build_hint(i_host) -> int:
    """Kernel build of the sequential forward trace chosen from the host copy of the table: 0 = the library's default
    (since the block-size A/B of round 2 — profiles/r2_block_size_ab.md — one build wins on every BASELINE table: the
    tile kernel with 2 rays per thread in one 1024-thread block per SM; the packed-pair streaming kernel (16) that short
    walks used to get stays selectable with ``set_tuning(fwd=16)`` / ``RTT_FWD_TILE=16``)."""
    return 0


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
SRC_CFG = 10   # floats per ray source in `src_cfg`: kind, a0..a3, width, height, intensity, wavelength, first


def source_cfg_of(rays) -> List[float]:
    sp = rays.source
    return [float(sp["kind"]), *map(float, sp["a"]), float(sp["width"]), float(sp["height"]),
            float(sp["intensity"]), float(sp["wavelength"]), float(sp["first"])]


def _source_req(src_cfg: Sequence[float], pose: torch.Tensor, state: torch.Tensor):
    kind, a0, a1, a2, a3, width, height, inten, wav, first = src_cfg
    return _cabi.make_source(int(kind), [a0, a1, a2, a3], pose.data_ptr(), first=int(first), state_ptr=state.data_ptr(),
                             width=int(width), height=int(height), intensity=inten, wavelength=wav)


def _seq_fwd_body(dev, n, pos, dir, intensity, wavelength, src, table_f, table_i, lut, lut_w, sensor_cfg,
                  want_record, want_rays, mode):
    lib = _cabi.load()
    ns = len(sensor_cfg) // SENSOR_CFG
    f32 = dict(dtype=torch.float32, device=dev)
    opos = torch.empty((n, 3) if want_rays else (0, 3), **f32)
    odir = torch.empty((n, 3) if want_rays else (0, 3), **f32)
    oint = torch.empty(n if want_rays else 0, **f32)
    hitmask = torch.empty(n, dtype=torch.int64, device=dev)
    records = torch.zeros((ns, n, 4), **f32) if (want_record and ns) else torch.empty((0, n, 4), **f32)
    images = torch.zeros(_image_numel(sensor_cfg), **f32)
    sens, cnt = _sensor_reqs(sensor_cfg, n, records if (want_record and ns) else None, images)
    req = _table_req(table_f, table_i, lut, lut_w)
    with torch.cuda.device(dev):
        lib.call("rtt_trace_seq_fwd", _ptr(pos), _ptr(dir), _ptr(intensity), _ptr(wavelength),
                 ct.byref(src) if src is not None else None,
                 _ptr(opos), _ptr(odir), _ptr(oint), hitmask.data_ptr(),
                 ct.byref(req), sens, cnt, n, _with_tune(mode, _tune_fwd), _stream(table_f))
    return [opos, odir, oint, hitmask, records, images]


def _seq_bwd_body(dev, n, pos, dir, intensity, wavelength, src, hitmask, g_pos, g_dir, g_int, g_records,
                  table_f, table_i, lut, lut_w, need_rays, need_table, mode):
    lib = _cabi.load()
    S = table_f.shape[0]
    f32 = dict(dtype=torch.float32, device=dev)
    # every unrequested output is its own empty tensor: custom-op returns must not alias each other
    gp = torch.empty((n, 3) if need_rays else 0, **f32)
    gd = torch.empty((n, 3) if need_rays else 0, **f32)
    gi = torch.empty(n if need_rays else 0, **f32)
    gt = torch.zeros((S, C.ROW_G), **f32) if need_table else torch.empty(0, **f32)
    has_lut = lut is not None and lut.numel() > 0
    gl = torch.zeros_like(lut) if (need_table and has_lut) else torch.empty(0, **f32)
    ns = 0 if g_records is None else g_records.shape[0]
    rec_arr = (ct.c_void_p * ns)(*[g_records[s].data_ptr() for s in range(ns)]) if ns else None
    req = _table_req(table_f, table_i, lut, lut_w)
    with torch.cuda.device(dev):
        lib.call("rtt_trace_seq_bwd", _ptr(pos), _ptr(dir), _ptr(intensity), _ptr(wavelength),
                 ct.byref(src) if src is not None else None,
                 hitmask.data_ptr(), _ptr(g_pos), _ptr(g_dir), _ptr(g_int), rec_arr,
                 _ptr(gp), _ptr(gd), _ptr(gi), _ptr(gt), _ptr(gl), ct.byref(req), ns, n,
                 _with_tune(mode & ~MODE_TUNE_MASK, _tune_bwd), _stream(table_f))
    return [gp, gd, gi, gt, gl]


def _nonseq_fwd_body(dev, n, pos, dir, intensity, wavelength, src, table_f, table_i, lut, lut_w, sensor_cfg,
                     want_record, want_rays, nbounces, mode, record_depth):
    lib = _cabi.load()
    ns = len(sensor_cfg) // SENSOR_CFG
    f32 = dict(dtype=torch.float32, device=dev)
    opos = torch.empty((n, 3) if want_rays else (0, 3), **f32)
    odir = torch.empty((n, 3) if want_rays else (0, 3), **f32)
    oint = torch.empty(n if want_rays else 0, **f32)
    seq = torch.empty((n, nbounces), dtype=torch.uint8, device=dev)
    nh = torch.empty(n, dtype=torch.uint8, device=dev)
    K = max(1, int(record_depth))
    rec_on = bool(want_record and ns)
    records = torch.zeros((ns, K, n, 4), **f32) if rec_on else torch.empty((0, K, n, 4), **f32)
    counts = torch.zeros((ns if rec_on else 0, n), dtype=torch.uint8, device=dev)
    images = torch.zeros(_image_numel(sensor_cfg), **f32)
    sens, cnt = _sensor_reqs(sensor_cfg, n, records if rec_on else None, images, counts if rec_on else None, K)
    req = _table_req(table_f, table_i, lut, lut_w)
    with torch.cuda.device(dev):
        lib.call("rtt_trace_nonseq_fwd", _ptr(pos), _ptr(dir), _ptr(intensity), _ptr(wavelength),
                 ct.byref(src) if src is not None else None,
                 _ptr(opos), _ptr(odir), _ptr(oint), seq.data_ptr(), nh.data_ptr(),
                 ct.byref(req), sens, cnt, nbounces, n, _with_tune(mode, _tune_nonseq), _stream(table_f))
    return [opos, odir, oint, seq, nh, records, images, counts]


def _nonseq_bwd_body(dev, n, pos, dir, intensity, wavelength, src, hit_seq, g_pos, g_dir, g_int, g_records,
                     table_f, table_i, lut, lut_w, need_rays, need_table, mode):
    lib = _cabi.load()
    S = table_f.shape[0]
    f32 = dict(dtype=torch.float32, device=dev)
    # every unrequested output is its own empty tensor: custom-op returns must not alias each other
    gp = torch.empty((n, 3) if need_rays else 0, **f32)
    gd = torch.empty((n, 3) if need_rays else 0, **f32)
    gi = torch.empty(n if need_rays else 0, **f32)
    gt = torch.zeros((S, C.ROW_G), **f32) if need_table else torch.empty(0, **f32)
    has_lut = lut is not None and lut.numel() > 0
    gl = torch.zeros_like(lut) if (need_table and has_lut) else torch.empty(0, **f32)
    ns = 0 if g_records is None else g_records.shape[0]          # g_records: [ns, K, N, 4]
    rec_arr = (ct.c_void_p * ns)(*[g_records[s].data_ptr() for s in range(ns)]) if ns else None
    depth = (ct.c_int32 * ns)(*([g_records.shape[1]] * ns)) if ns else None
    req = _table_req(table_f, table_i, lut, lut_w)
    with torch.cuda.device(dev):
        lib.call("rtt_trace_nonseq_bwd", _ptr(pos), _ptr(dir), _ptr(intensity), _ptr(wavelength),
                 ct.byref(src) if src is not None else None,
                 hit_seq.data_ptr(), hit_seq.shape[1], _ptr(g_pos), _ptr(g_dir), _ptr(g_int), rec_arr, depth,
                 _ptr(gp), _ptr(gd), _ptr(gi), _ptr(gt), _ptr(gl), ct.byref(req), ns, n, mode, _stream(table_f))
    return [gp, gd, gi, gt, gl]


def _fake_seq_fwd(like, n, sensor_cfg, want_record, want_rays):
    ns = len(sensor_cfg) // SENSOR_CFG
    m = n if want_rays else 0
    return [like.new_empty((m, 3)), like.new_empty((m, 3)), like.new_empty(m), like.new_empty(n, dtype=torch.int64),
            like.new_empty(((ns if want_record else 0), n, 4)), like.new_empty(_image_numel(sensor_cfg))]


def _fake_bwd(like, n, table_f, lut, need_rays, need_table):
    return [like.new_empty((n, 3) if need_rays else 0), like.new_empty((n, 3) if need_rays else 0),
            like.new_empty(n if need_rays else 0),
            like.new_empty((table_f.shape[0], C.ROW_G) if need_table else 0),
            torch.empty_like(lut) if (need_table and lut is not None) else like.new_empty(0)]


def _fake_nonseq_fwd(like, n, sensor_cfg, want_record, want_rays, nbounces, record_depth):
    ns = len(sensor_cfg) // SENSOR_CFG
    K = max(1, int(record_depth))
    m = n if want_rays else 0
    return [like.new_empty((m, 3)), like.new_empty((m, 3)), like.new_empty(m),
            like.new_empty((n, nbounces), dtype=torch.uint8), like.new_empty(n, dtype=torch.uint8),
            like.new_empty(((ns if want_record else 0), K, n, 4)), like.new_empty(_image_numel(sensor_cfg)),
            like.new_empty(((ns if want_record else 0), n), dtype=torch.uint8)]


@torch.library.custom_op("rtt_b200::trace_seq_fwd", mutates_args=())
def _trace_seq_fwd(pos: torch.Tensor, dir: torch.Tensor, intensity: torch.Tensor,
                   wavelength: Optional[torch.Tensor], table_f: torch.Tensor, table_i: torch.Tensor,
                   lut: Optional[torch.Tensor], lut_w: Optional[torch.Tensor],
                   sensor_cfg: List[float], want_record: bool, mode: int) -> List[torch.Tensor]:
    """-> [out_pos, out_dir, out_intensity, hitmask(int64), records [ns,N,4], images (flat)]"""
    _need_cuda(pos, dir, intensity, table_f, table_i)
    want_rays = not (mode & MODE_NO_FINAL_RAYS)         # op-level flag (not part of the C ABI): empty final-ray outputs
    return _seq_fwd_body(pos.device, pos.shape[0], pos, dir, intensity, wavelength, None, table_f, table_i, lut, lut_w,
                         sensor_cfg, want_record, want_rays, mode & ~MODE_NO_FINAL_RAYS)


@_trace_seq_fwd.register_fake
def _(pos, dir, intensity, wavelength, table_f, table_i, lut, lut_w, sensor_cfg, want_record, mode):
    return _fake_seq_fwd(pos, pos.shape[0], sensor_cfg, want_record, not (mode & MODE_NO_FINAL_RAYS))


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
    return _seq_bwd_body(pos.device, pos.shape[0], pos, dir, intensity, wavelength, None, hitmask, g_pos, g_dir, g_int,
                         g_records, table_f, table_i, lut, lut_w, need_rays, need_table, mode)


@_trace_seq_bwd.register_fake
def _(pos, dir, intensity, wavelength, hitmask, g_pos, g_dir, g_int, g_records, table_f, table_i, lut, lut_w,
      need_rays, need_table, mode):
    return _fake_bwd(pos, pos.shape[0], table_f, lut, need_rays, need_table)


@torch.library.custom_op("rtt_b200::trace_seq_src_fwd", mutates_args=())
def _trace_seq_src_fwd(src_cfg: List[float], pose: torch.Tensor, state: torch.Tensor, n: int,
                       table_f: torch.Tensor, table_i: torch.Tensor,
                       lut: Optional[torch.Tensor], lut_w: Optional[torch.Tensor],
                       sensor_cfg: List[float], want_record: bool, want_rays: bool, mode: int) -> List[torch.Tensor]:
    """Sequential trace of rays GENERATED in the kernel from a ray source (no ray input in HBM).
    want_rays=False also skips the final-ray outputs (goals only read the sensor records)."""
    _need_cuda(pose, state, table_f, table_i)
    return _seq_fwd_body(pose.device, n, None, None, None, None, _source_req(src_cfg, pose, state), table_f, table_i,
                         lut, lut_w, sensor_cfg, want_record, want_rays, mode)


@_trace_seq_src_fwd.register_fake
def _(src_cfg, pose, state, n, table_f, table_i, lut, lut_w, sensor_cfg, want_record, want_rays, mode):
    return _fake_seq_fwd(pose, n, sensor_cfg, want_record, want_rays)


@torch.library.custom_op("rtt_b200::trace_seq_src_bwd", mutates_args=())
def _trace_seq_src_bwd(src_cfg: List[float], pose: torch.Tensor, state: torch.Tensor, n: int, hitmask: torch.Tensor,
                       g_pos: Optional[torch.Tensor], g_dir: Optional[torch.Tensor], g_int: Optional[torch.Tensor],
                       g_records: Optional[torch.Tensor],
                       table_f: torch.Tensor, table_i: torch.Tensor,
                       lut: Optional[torch.Tensor], lut_w: Optional[torch.Tensor],
                       need_table: bool, mode: int) -> List[torch.Tensor]:
    """Adjoint of trace_seq_src_fwd: the rays are regenerated from the same {key, counter}."""
    _need_cuda(pose, state, table_f, table_i, hitmask)
    return _seq_bwd_body(pose.device, n, None, None, None, None, _source_req(src_cfg, pose, state), hitmask,
                         g_pos, g_dir, g_int, g_records, table_f, table_i, lut, lut_w, False, need_table, mode)


@_trace_seq_src_bwd.register_fake
def _(src_cfg, pose, state, n, hitmask, g_pos, g_dir, g_int, g_records, table_f, table_i, lut, lut_w, need_table, mode):
    return _fake_bwd(pose, n, table_f, lut, False, need_table)


@torch.library.custom_op("rtt_b200::trace_nonseq_fwd", mutates_args=())
def _trace_nonseq_fwd(pos: torch.Tensor, dir: torch.Tensor, intensity: torch.Tensor,
                      wavelength: Optional[torch.Tensor], table_f: torch.Tensor, table_i: torch.Tensor,
                      lut: Optional[torch.Tensor], lut_w: Optional[torch.Tensor],
                      sensor_cfg: List[float], want_record: bool, nbounces: int, mode: int,
                      record_depth: int = 1) -> List[torch.Tensor]:
    """-> [out_pos, out_dir, out_intensity, hit_seq uint8 [N,B], n_hits uint8 [N], records [ns,K,N,4],
    images (flat), counts uint8 [ns,N]]; K = record_depth = sensor interactions kept per ray."""
    _need_cuda(pos, dir, intensity, table_f, table_i)
    return _nonseq_fwd_body(pos.device, pos.shape[0], pos, dir, intensity, wavelength, None, table_f, table_i, lut,
                            lut_w, sensor_cfg, want_record, True, nbounces, mode, record_depth)


@_trace_nonseq_fwd.register_fake
def _(pos, dir, intensity, wavelength, table_f, table_i, lut, lut_w, sensor_cfg, want_record, nbounces, mode,
      record_depth=1):
    return _fake_nonseq_fwd(pos, pos.shape[0], sensor_cfg, want_record, True, nbounces, record_depth)


@torch.library.custom_op("rtt_b200::trace_nonseq_bwd", mutates_args=())
def _trace_nonseq_bwd(pos: torch.Tensor, dir: torch.Tensor, intensity: torch.Tensor,
                      wavelength: Optional[torch.Tensor], hit_seq: torch.Tensor,
                      g_pos: Optional[torch.Tensor], g_dir: Optional[torch.Tensor], g_int: Optional[torch.Tensor],
                      g_records: Optional[torch.Tensor],
                      table_f: torch.Tensor, table_i: torch.Tensor,
                      lut: Optional[torch.Tensor], lut_w: Optional[torch.Tensor],
                      need_rays: bool, need_table: bool, mode: int) -> List[torch.Tensor]:
    _need_cuda(pos, dir, intensity, table_f, table_i, hit_seq)
    return _nonseq_bwd_body(pos.device, pos.shape[0], pos, dir, intensity, wavelength, None, hit_seq, g_pos, g_dir,
                            g_int, g_records, table_f, table_i, lut, lut_w, need_rays, need_table, mode)


@_trace_nonseq_bwd.register_fake
def _(pos, dir, intensity, wavelength, hit_seq, g_pos, g_dir, g_int, g_records, table_f, table_i, lut, lut_w,
      need_rays, need_table, mode):
    return _fake_bwd(pos, pos.shape[0], table_f, lut, need_rays, need_table)


@torch.library.custom_op("rtt_b200::trace_nonseq_src_fwd", mutates_args=())
def _trace_nonseq_src_fwd(src_cfg: List[float], pose: torch.Tensor, state: torch.Tensor, n: int,
                          table_f: torch.Tensor, table_i: torch.Tensor,
                          lut: Optional[torch.Tensor], lut_w: Optional[torch.Tensor],
                          sensor_cfg: List[float], want_record: bool, want_rays: bool, nbounces: int, mode: int,
                          record_depth: int) -> List[torch.Tensor]:
    _need_cuda(pose, state, table_f, table_i)
    return _nonseq_fwd_body(pose.device, n, None, None, None, None, _source_req(src_cfg, pose, state), table_f, table_i,
                            lut, lut_w, sensor_cfg, want_record, want_rays, nbounces, mode, record_depth)


@_trace_nonseq_src_fwd.register_fake
def _(src_cfg, pose, state, n, table_f, table_i, lut, lut_w, sensor_cfg, want_record, want_rays, nbounces, mode,
      record_depth):
    return _fake_nonseq_fwd(pose, n, sensor_cfg, want_record, want_rays, nbounces, record_depth)


@torch.library.custom_op("rtt_b200::trace_nonseq_src_bwd", mutates_args=())
def _trace_nonseq_src_bwd(src_cfg: List[float], pose: torch.Tensor, state: torch.Tensor, n: int, hit_seq: torch.Tensor,
                          g_pos: Optional[torch.Tensor], g_dir: Optional[torch.Tensor], g_int: Optional[torch.Tensor],
                          g_records: Optional[torch.Tensor],
                          table_f: torch.Tensor, table_i: torch.Tensor,
                          lut: Optional[torch.Tensor], lut_w: Optional[torch.Tensor],
                          need_table: bool, mode: int) -> List[torch.Tensor]:
    _need_cuda(pose, state, table_f, table_i, hit_seq)
    return _nonseq_bwd_body(pose.device, n, None, None, None, None, _source_req(src_cfg, pose, state), hit_seq,
                            g_pos, g_dir, g_int, g_records, table_f, table_i, lut, lut_w, False, need_table, mode)


@_trace_nonseq_src_bwd.register_fake
def _(src_cfg, pose, state, n, hit_seq, g_pos, g_dir, g_int, g_records, table_f, table_i, lut, lut_w, need_table, mode):
    return _fake_bwd(pose, n, table_f, lut, False, need_table)


@torch.library.custom_op("rtt_b200::sample_bundle", mutates_args=())
def _sample_bundle(src_cfg: List[float], pose: torch.Tensor, state: torch.Tensor, n: int, mode: int) -> List[torch.Tensor]:
    """Bundle.sample on the device (rays/bundle.py:30-37) -> [pos, dir, intensity, wavelength]."""
    _need_cuda(pose, state)
    lib = _cabi.load()
    f32 = dict(dtype=torch.float32, device=pose.device)
    pos, dir_ = torch.empty((n, 3), **f32), torch.empty((n, 3), **f32)
    inten, wav = torch.empty(n, **f32), torch.empty(n, **f32)
    src = _source_req(src_cfg, pose, state)
    with torch.cuda.device(pose.device):
        lib.call("rtt_sample_bundle", ct.byref(src), _ptr(pos), _ptr(dir_), _ptr(inten), _ptr(wav), n, mode,
                 _stream(pose))
    return [pos, dir_, inten, wav]


@_sample_bundle.register_fake
def _(src_cfg, pose, state, n, mode):
    return [pose.new_empty((n, 3)), pose.new_empty((n, 3)), pose.new_empty(n), pose.new_empty(n)]


# ---- goal reductions (rtt_goals.cu) ----------------------------------------------------------------
_SPOT_WORK = {}


def _spot_work(dev) -> torch.Tensor:
    """Per-device, per-stream scratch of the two-stage reductions (zeroed once; the kernels leave it zeroed)."""
    key = (dev, torch.cuda.current_stream(dev).cuda_stream)
    w = _SPOT_WORK.get(key)
    if w is None:
        w = _SPOT_WORK[key] = torch.zeros(_cabi.SPOT_WORK, dtype=torch.float32, device=dev)
    return w


@torch.library.custom_op("rtt_b200::spot_moments", mutates_args=())
def _spot_moments(rec: torch.Tensor, active_only: bool) -> torch.Tensor:
    """rec [M,4] -> [sum w, sum w x, sum w y, #(w > 0)]"""
    _need_cuda(rec)
    out = torch.empty(4, dtype=torch.float32, device=rec.device)
    with torch.cuda.device(rec.device):
        _cabi.load().call("rtt_spot_moments", _ptr(rec), rec.shape[0], int(active_only), out.data_ptr(),
                          _spot_work(rec.device).data_ptr(), _stream(rec))
    return out


@_spot_moments.register_fake
def _(rec, active_only):
    return rec.new_empty(4)


@torch.library.custom_op("rtt_b200::spot_moments_bwd", mutates_args=())
def _spot_moments_bwd(rec: torch.Tensor, active_only: bool, g3: torch.Tensor) -> torch.Tensor:
    _need_cuda(rec, g3)
    g = torch.empty_like(rec)
    with torch.cuda.device(rec.device):
        _cabi.load().call("rtt_spot_moments_bwd", _ptr(rec), rec.shape[0], int(active_only), g3.data_ptr(), _ptr(g),
                          _stream(rec))
    return g


@_spot_moments_bwd.register_fake
def _(rec, active_only, g3):
    return torch.empty_like(rec)


@torch.library.custom_op("rtt_b200::spot_size_fwd", mutates_args=())
def _spot_size_fwd(rec: torch.Tensor, mom4: torch.Tensor, target: Optional[torch.Tensor]) -> torch.Tensor:
    """-> [sum_i sqrt(q_i), d/d cx, d/d cy]"""
    _need_cuda(rec, mom4)
    out = torch.empty(3, dtype=torch.float32, device=rec.device)
    with torch.cuda.device(rec.device):
        _cabi.load().call("rtt_spot_size_fwd", _ptr(rec), rec.shape[0], mom4.data_ptr(), _ptr(target), out.data_ptr(),
                          _spot_work(rec.device).data_ptr(), _stream(rec))
    return out


@_spot_size_fwd.register_fake
def _(rec, mom4, target):
    return rec.new_empty(3)


@torch.library.custom_op("rtt_b200::spot_size_bwd", mutates_args=())
def _spot_size_bwd(rec: torch.Tensor, mom4: torch.Tensor, target: Optional[torch.Tensor], out3: torch.Tensor,
                   g_loss: torch.Tensor) -> torch.Tensor:
    _need_cuda(rec, mom4, out3, g_loss)
    g = torch.empty_like(rec)
    with torch.cuda.device(rec.device):
        _cabi.load().call("rtt_spot_size_bwd", _ptr(rec), rec.shape[0], mom4.data_ptr(), _ptr(target), out3.data_ptr(),
                          g_loss.data_ptr(), _ptr(g), _stream(rec))
    return g


@_spot_size_bwd.register_fake
def _(rec, mom4, target, out3, g_loss):
    return torch.empty_like(rec)


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
    def forward(ctx, pos, dir_, intensity, wavelength, table_f, table_i, lut, lut_w, sensor_cfg, want_record, mode,
                bwd_hint=0):
        outs = torch.ops.rtt_b200.trace_seq_fwd(pos, dir_, intensity, wavelength, table_f, table_i, lut, lut_w,
                                                sensor_cfg, want_record, mode)
        opos, odir, oint, hitmask, records, images = outs
        ctx.save_for_backward(pos, dir_, intensity, wavelength, hitmask, table_f, table_i, lut, lut_w)
        ctx.mode = (mode & ~MODE_NO_FINAL_RAYS) | bwd_hint
        ctx.mark_non_differentiable(hitmask, images)
        ctx.set_materialize_grads(False)     # unused outputs hand None to backward, not [N,3] zero tensors
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
                gl if ctx.needs_input_grad[6] else None, None, None, None, None, None)


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
        ctx.set_materialize_grads(False)     # unused outputs hand None to backward, not [N,3] zero tensors
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


class _TraceSeqSrc(torch.autograd.Function):
    """Sequential trace of in-kernel generated rays; only the table (and LUT) receive gradients."""

    @staticmethod
    def forward(ctx, src_cfg, pose, state, n, table_f, table_i, lut, lut_w, sensor_cfg, want_record, want_rays, mode,
                bwd_hint=0):
        outs = torch.ops.rtt_b200.trace_seq_src_fwd(src_cfg, pose, state, n, table_f, table_i, lut, lut_w,
                                                    sensor_cfg, want_record, want_rays, mode)
        opos, odir, oint, hitmask, records, images = outs
        ctx.save_for_backward(pose, state, hitmask, table_f, table_i, lut, lut_w)
        ctx.src_cfg, ctx.n, ctx.mode, ctx.want_rays = list(src_cfg), n, mode | bwd_hint, want_rays
        ctx.mark_non_differentiable(hitmask, images)
        ctx.set_materialize_grads(False)     # unused outputs hand None to backward, not [N,3] zero tensors
        return opos, odir, oint, hitmask, records, images

    @staticmethod
    def backward(ctx, g_pos, g_dir, g_int, _g_mask, g_records, _g_images):
        pose, state, hitmask, table_f, table_i, lut, lut_w = ctx.saved_tensors
        if g_records is not None and g_records.numel() == 0:
            g_records = None
        if not ctx.want_rays:
            g_pos = g_dir = g_int = None
        _gp, _gd, _gi, gt, gl = torch.ops.rtt_b200.trace_seq_src_bwd(
            ctx.src_cfg, pose, state, ctx.n, hitmask, _cg(g_pos), _cg(g_dir), _cg(g_int), _cg(g_records),
            table_f, table_i, lut, lut_w, True, ctx.mode)
        return (None, None, None, None, _pad_table_grad(gt, table_f) if ctx.needs_input_grad[4] else None, None,
                gl if ctx.needs_input_grad[6] else None, None, None, None, None, None, None)


class _TraceNonseqSrc(torch.autograd.Function):
    @staticmethod
    def forward(ctx, src_cfg, pose, state, n, table_f, table_i, lut, lut_w, sensor_cfg, want_record, want_rays,
                nbounces, mode, record_depth):
        outs = torch.ops.rtt_b200.trace_nonseq_src_fwd(src_cfg, pose, state, n, table_f, table_i, lut, lut_w,
                                                       sensor_cfg, want_record, want_rays, nbounces, mode, record_depth)
        opos, odir, oint, seq, nh, records, images, counts = outs
        ctx.save_for_backward(pose, state, seq, table_f, table_i, lut, lut_w)
        ctx.src_cfg, ctx.n, ctx.mode, ctx.want_rays = list(src_cfg), n, mode, want_rays
        ctx.mark_non_differentiable(seq, nh, images, counts)
        ctx.set_materialize_grads(False)     # unused outputs hand None to backward, not [N,3] zero tensors
        return opos, odir, oint, seq, nh, records, images, counts

    @staticmethod
    def backward(ctx, g_pos, g_dir, g_int, _g_seq, _g_nh, g_records, _g_images, _g_counts):
        pose, state, seq, table_f, table_i, lut, lut_w = ctx.saved_tensors
        if g_records is not None and g_records.numel() == 0:
            g_records = None
        if not ctx.want_rays:
            g_pos = g_dir = g_int = None
        _gp, _gd, _gi, gt, gl = torch.ops.rtt_b200.trace_nonseq_src_bwd(
            ctx.src_cfg, pose, state, ctx.n, seq, _cg(g_pos), _cg(g_dir), _cg(g_int), _cg(g_records),
            table_f, table_i, lut, lut_w, True, ctx.mode)
        return (None, None, None, None, _pad_table_grad(gt, table_f) if ctx.needs_input_grad[4] else None, None,
                gl if ctx.needs_input_grad[6] else None, None, None, None, None, None, None, None)


def _all_reduce_sum(t: torch.Tensor) -> torch.Tensor:
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t


class _SpotMoments(torch.autograd.Function):
    """[W, sum w x, sum w y, count] of sensor records (summed over ranks when torch.distributed is up)."""

    @staticmethod
    def forward(ctx, rec, active_only):
        mom = _all_reduce_sum(torch.ops.rtt_b200.spot_moments(rec, active_only))
        ctx.save_for_backward(rec)
        ctx.active_only = active_only
        return mom

    @staticmethod
    def backward(ctx, g):
        (rec,) = ctx.saved_tensors
        return torch.ops.rtt_b200.spot_moments_bwd(rec, ctx.active_only, g[:3].contiguous()), None


class _SpotSize(torch.autograd.Function):
    """sum_i sqrt(|xy_i - c|^2 w_i / W) over the active records (optim/goals.py:165-183), c = weighted centroid
    or a fixed target.  Two reduction launches forward, one elementwise launch backward; sums are global over
    ranks, so every rank holds the same loss and its records get the full derivative."""

    @staticmethod
    def forward(ctx, rec, target):
        mom = _all_reduce_sum(torch.ops.rtt_b200.spot_moments(rec, True))
        out3 = _all_reduce_sum(torch.ops.rtt_b200.spot_size_fwd(rec, mom, target))
        ctx.save_for_backward(rec, mom, out3, target)
        active = (mom[3] > 0).to(torch.float32)          # 1 if any ray of the bundle reached the sensor with w > 0
        ctx.mark_non_differentiable(active)
        return out3[0].clone(), active

    @staticmethod
    def backward(ctx, g, _g_active):
        rec, mom, out3, target = ctx.saved_tensors
        return torch.ops.rtt_b200.spot_size_bwd(rec, mom, target, out3, g.reshape(1).contiguous()), None


def spot_moments(rec: torch.Tensor, active_only: bool = False) -> torch.Tensor:
    """Differentiable [sum w, sum w x, sum w y, #(w>0)] of sensor records rec [..., 4] (summed over ranks).  An empty
    local shard still takes part in the collective (a rank must never skip an all-reduce its peers run)."""
    rec = _f32c(rec).reshape(-1, 4)
    if rec.shape[0] == 0:
        return _all_reduce_sum(torch.zeros(4, dtype=torch.float32, device=rec.device))
    return _SpotMoments.apply(rec, bool(active_only))


def spot_size_active(rec: torch.Tensor, target_xy: Optional[torch.Tensor] = None):
    """(SpotSizeLoss term of ONE bundle, active flag) from its sensor records rec [..., 4]: the flag is a device
    scalar, 1.0 iff some ray reached the sensor with positive weight — the reference skips bundles without active
    hits (optim/goals.py:165-167), which the caller reproduces by weighting, without a host synchronisation."""
    tgt = None if target_xy is None else _f32c(target_xy.to(rec.device)).reshape(2)
    return _SpotSize.apply(_f32c(rec).reshape(-1, 4), tgt)


def spot_size(rec: torch.Tensor, target_xy: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Differentiable SpotSizeLoss term of ONE bundle from its sensor records rec [..., 4]."""
    return spot_size_active(rec, target_xy)[0]

@torch.library.custom_op("rtt_b200::render_shade", mutates_args=())
def _render_shade(pos: Optional[torch.Tensor], dir: Optional[torch.Tensor], src_cfg: List[float],
                  pose: Optional[torch.Tensor], state: Optional[torch.Tensor], n: int,
                  table_f: torch.Tensor, table_i: torch.Tensor, base_rgb: torch.Tensor,
                  light: List[float], background: List[float]) -> List[torch.Tensor]:
    """Renderer.render_3d in one launch (rtt_render_shade) -> [rgb [n,3] f32, winning row [n] u8 (255 = background)];
    rays from (pos, dir) or — pose/state given — generated in the kernel from the camera source ``src_cfg``."""
    _need_cuda(table_f, table_i, base_rgb)
    dev = table_f.device
    rgb = torch.empty((n, 3), dtype=torch.float32, device=dev)
    win = torch.empty(n, dtype=torch.uint8, device=dev)
    src = _source_req(src_cfg, pose, state) if pose is not None else None
    req = _table_req(table_f, table_i, None, None)
    la, bg = (ct.c_float * 3)(*light), (ct.c_float * 3)(*background)
    with torch.cuda.device(dev):
        _cabi.load().call("rtt_render_shade", _ptr(pos), _ptr(dir), ct.byref(src) if src is not None else None,
                          ct.byref(req), base_rgb.data_ptr(), la, bg, rgb.data_ptr(), win.data_ptr(), n, MODE_EXACT,
                          _stream(table_f))
    return [rgb, win]


@_render_shade.register_fake
def _(pos, dir, src_cfg, pose, state, n, table_f, table_i, base_rgb, light, background):
    return [table_f.new_empty((n, 3)), table_f.new_empty(n, dtype=torch.uint8)]


# ---- per-id sensor moments (rtt_goals.cu k_spot_id_*) -------------------------------------------------------------
_SPOT_ID_WORK = {}


def _spot_id_work(dev) -> torch.Tensor:
    key = (dev, torch.cuda.current_stream(dev).cuda_stream)
    w = _SPOT_ID_WORK.get(key)
    if w is None:
        w = _SPOT_ID_WORK[key] = torch.zeros(_cabi.SPOT_ID_WORK, dtype=torch.float32, device=dev)
    return w


@torch.library.custom_op("rtt_b200::spot_id_moments", mutates_args=())
def _spot_id_moments(rec: torch.Tensor, ids: torch.Tensor, group_of: torch.Tensor, n_groups: int) -> torch.Tensor:
    """rec [M,4], ids int8 [M], group_of int32 [256] -> [K,4] = (sum w, sum w x, sum w y, #(w > 0)) per queried id"""
    _need_cuda(rec, ids, group_of)
    out = torch.empty((n_groups, 4), dtype=torch.float32, device=rec.device)
    with torch.cuda.device(rec.device):
        _cabi.load().call("rtt_spot_id_moments", _ptr(rec), _ptr(ids), rec.shape[0], group_of.data_ptr(), n_groups,
                          out.data_ptr(), _spot_id_work(rec.device).data_ptr(), _stream(rec))
    return out


@_spot_id_moments.register_fake
def _(rec, ids, group_of, n_groups):
    return rec.new_empty((n_groups, 4))


@torch.library.custom_op("rtt_b200::spot_id_size", mutates_args=())
def _spot_id_size(rec: torch.Tensor, ids: torch.Tensor, group_of: torch.Tensor, n_groups: int, centres: torch.Tensor,
                  norm_ord: float) -> torch.Tensor:
    """-> [K,4] = (sum w (|dx|^p + |dy|^p), sum w p|dx|^(p-1) sgn dx, sum w p|dy|^(p-1) sgn dy, 0)"""
    _need_cuda(rec, ids, group_of, centres)
    out = torch.empty((n_groups, 4), dtype=torch.float32, device=rec.device)
    with torch.cuda.device(rec.device):
        _cabi.load().call("rtt_spot_id_size", _ptr(rec), _ptr(ids), rec.shape[0], group_of.data_ptr(), n_groups,
                          centres.data_ptr(), float(norm_ord), out.data_ptr(), _spot_id_work(rec.device).data_ptr(),
                          _stream(rec))
    return out


@_spot_id_size.register_fake
def _(rec, ids, group_of, n_groups, centres, norm_ord):
    return rec.new_empty((n_groups, 4))


@torch.library.custom_op("rtt_b200::spot_id_size_bwd", mutates_args=())
def _spot_id_size_bwd(rec: torch.Tensor, ids: torch.Tensor, group_of: torch.Tensor, n_groups: int, coef: torch.Tensor,
                      norm_ord: float) -> torch.Tensor:
    _need_cuda(rec, ids, group_of, coef)
    g = torch.empty_like(rec)
    with torch.cuda.device(rec.device):
        _cabi.load().call("rtt_spot_id_size_bwd", _ptr(rec), _ptr(ids), rec.shape[0], group_of.data_ptr(), n_groups,
                          coef.data_ptr(), float(norm_ord), _ptr(g), _stream(rec))
    return g


@_spot_id_size_bwd.register_fake
def _(rec, ids, group_of, n_groups, coef, norm_ord):
    return torch.empty_like(rec)


class _SpotSizePerId(torch.autograd.Function):
    """Sensor.getSpotSizeParallel_xy (elements/sensor.py:87-176) on dense records: per queried ray id,
    sum_i w_i (|x_i - cx|^p + |y_i - cy|^p) / (2 W), W = sum_i w_i (1 where no hit), c = the id's intensity centroid
    or its target.  Two reduction launches forward, one elementwise launch backward; sums are global over ranks."""

    @staticmethod
    def forward(ctx, rec, ids, group_of, n_groups, targets, norm_ord):
        mom = _all_reduce_sum(torch.ops.rtt_b200.spot_id_moments(rec, ids, group_of, n_groups))
        W = mom[:, 0]
        safe = torch.where(W == 0, torch.ones_like(W), W)                # safe_denom (sensor.py:124-125)
        centres = (mom[:, 1:3] / safe[:, None]) if targets is None else targets
        centres = centres.contiguous()
        s4 = _all_reduce_sum(torch.ops.rtt_b200.spot_id_size(rec, ids, group_of, n_groups, centres, norm_ord))
        ctx.save_for_backward(rec, ids, group_of, centres, s4, safe, W)
        ctx.meta = (n_groups, float(norm_ord), targets is None)
        ctx.mark_non_differentiable(W)
        return s4[:, 0] / (2.0 * safe), W

    @staticmethod
    def backward(ctx, g_out, _g_w):
        rec, ids, group_of, centres, s4, safe, W = ctx.saved_tensors
        K, p, free_centre = ctx.meta
        a = g_out / (2.0 * safe)
        hit = (W != 0).to(a.dtype)
        bx = s4[:, 1] / safe * hit if free_centre else torch.zeros_like(a)
        by = s4[:, 2] / safe * hit if free_centre else torch.zeros_like(a)
        sW = s4[:, 0] / safe * hit                                       # d/dW of S / (2W); no W-dependence where W was replaced by 1
        z = torch.zeros_like(a)
        coef = torch.stack([centres[:, 0], centres[:, 1], a, bx, by, sW, z, z], 1).contiguous()
        return torch.ops.rtt_b200.spot_id_size_bwd(rec, ids, group_of, K, coef, p), None, None, None, None, None


def _group_table(query_ids, device) -> torch.Tensor:
    """int32 [256]: group_of[id + 128] = position of `id` in query_ids, -1 elsewhere (int8 ids, rays/ray.py:17)."""
    q = torch.as_tensor(query_ids).to(torch.int64).reshape(-1).cpu()
    if q.numel() < 1 or q.numel() > 256 or int(q.min()) < -128 or int(q.max()) > 127 or q.unique().numel() != q.numel():
        raise ValueError("query_ids must be 1..256 distinct int8 ray ids")
    lut = torch.full((256,), -1, dtype=torch.int32)
    lut[q + 128] = torch.arange(q.numel(), dtype=torch.int32)
    return lut.to(device)


def spot_size_per_id(rec: torch.Tensor, ids: torch.Tensor, query_ids, target_xy: Optional[torch.Tensor] = None,
                     norm_ord: float = 2):
    """Differentiable per-id spot sizes of dense sensor records rec [M,4] with ray ids [M] int8:
    (spot_size [K] in the order of ``query_ids``, intensity_sum [K] in the same order)."""
    rec = _f32c(rec).reshape(-1, 4)
    ids = ids.reshape(-1).to(torch.int8).contiguous()
    lut = _group_table(query_ids, rec.device)
    K = int((lut >= 0).sum())
    tgt = None if target_xy is None else _f32c(torch.as_tensor(target_xy).to(rec.device)).reshape(K, 2)
    return _SpotSizePerId.apply(rec, ids, lut, K, tgt, float(norm_ord))


def sample_source(rays, mode: Optional[int] = None):
    """Materialise SourceRays: (pos, dir, intensity, wavelength) from rtt_sample_bundle."""
    mode = _default_mode if mode is None else mode
    return tuple(torch.ops.rtt_b200.sample_bundle(source_cfg_of(rays), rays.pose, rays.state, rays.n, mode))


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


def adjoint_hint(table: SurfaceTable) -> int:
    """``MODE_SCALAR_GRADS`` when no row of ``table`` requests pose gradients (read from the host copy of the int
    block), else 0: OR-ed into the mode of the sequential adjoint, which then runs its build without pose-gradient code."""
    pose = C.FLAG_GRAD_POSE_E | C.FLAG_GRAD_POSE_S
    return 0 if any(m[C.I_FLAGS] & pose for m in table.i_host) else MODE_SCALAR_GRADS


def trace_sequential(table: SurfaceTable, pos=None, dir_=None, intensity=None, wavelength=None, *, want_record=True,
                     sensor_cfg: Optional[List[float]] = None, mode: Optional[int] = None, source=None,
                     want_rays: bool = True):
    """Fused SequentialScene.simulate (scene/sequential.py:12-36).

    Rays come from (pos, dir_, intensity, wavelength) or — ``source=SourceRays`` — are generated in the kernel;
    ``want_rays=False`` skips the final-ray outputs (empty tensors; what the goals ask for: they read sensor records).
    Returns dict(pos, dir, intensity, hitmask [N] int64 (bit r = interacted with row r),
    records [n_sensors,N,4] (hit_local xyz, weight-before), images [per sensor: [C,H,W] or None])."""
    cfg = sensor_cfg_of(table) if sensor_cfg is None else list(sensor_cfg)
    mode = _default_mode if mode is None else mode
    hint = adjoint_hint(table)
    if not _tune_fwd and not (mode & MODE_TUNE_MASK) and (mode & 0xFF) == MODE_FAST:
        # the build depends on the TABLE only: rays generated in the kernel and their materialised twin run the same
        # arithmetic and stay bit-identical (tests/test_goals.py)
        mode |= _fwd_build_hint(table.i_host) << MODE_TUNE_SHIFT
    if source is not None:
        _need_cuda(table.f)
        opos, odir, oint, hitmask, records, images = _TraceSeqSrc.apply(
            source_cfg_of(source), source.pose, source.state, source.n, table.f, table.i, table.lut,
            table.lut_wavelengths, cfg, bool(want_record), bool(want_rays), mode, hint)
        return dict(pos=opos, dir=odir, intensity=oint, hitmask=hitmask, records=records,
                    images=split_images(images, cfg))
    pos, dir_, intensity, wav = _prep_rays(pos, dir_, intensity, wavelength, table)
    opos, odir, oint, hitmask, records, images = _TraceSeq.apply(
        pos, dir_, intensity, wav, table.f, table.i, table.lut, table.lut_wavelengths, cfg, bool(want_record),
        mode | (0 if want_rays else MODE_NO_FINAL_RAYS), hint)
    return dict(pos=opos, dir=odir, intensity=oint, hitmask=hitmask, records=records,
                images=split_images(images, cfg))


_copy_streams = {}


def _copy_stream(dev: torch.device) -> "torch.cuda.Stream":
    key = (dev.type, dev.index)
    if key not in _copy_streams:
        _copy_streams[key] = torch.cuda.Stream(device=dev)
    return _copy_streams[key]


def trace_sequential_host(table: SurfaceTable, pos, dir_, intensity, wavelength=None, *, want_record=False,
                          sensor_cfg: Optional[List[float]] = None, mode: Optional[int] = None,
                          chunk_rays: int = 1 << 23, ids=None):
    """``trace_sequential`` for a bundle that lives in HOST memory (pinned for full copy speed).

    The bundle is cut into chunks; chunk k+1 is copied host->device on a side stream while chunk k is traced
    (one ``rtt_trace_seq_fwd`` launch per chunk, all accumulating into the same sensor images), so the step costs
    about max(copy, trace) instead of their sum.  Forward only: no autograd graph is recorded (move the rays to
    the device to differentiate).  Returns the same dict as ``trace_sequential`` plus the device copies of the
    inputs (``in_pos, in_dir, in_intensity, in_wavelength``), everything on the table's device."""
    _need_cuda(table.f)
    for t in (pos, dir_, intensity):
        if t.is_cuda:
            raise ValueError("trace_sequential_host takes host tensors; use trace_sequential for device rays")
    dev = table.f.device
    lib = _cabi.load()
    cfg = sensor_cfg_of(table) if sensor_cfg is None else list(sensor_cfg)
    mode = _default_mode if mode is None else mode
    if not _tune_fwd and not (mode & MODE_TUNE_MASK) and (mode & 0xFF) == MODE_FAST:
        mode |= _fwd_build_hint(table.i_host) << MODE_TUNE_SHIFT
    f32 = dict(dtype=torch.float32, device=dev)
    pos, dir_, intensity = (_f32c(t.detach()) for t in (pos, dir_, intensity))
    use_wav = table.lut is not None and wavelength is not None       # the kernel reads it only with an index table
    wav = _f32c(wavelength.detach()) if wavelength is not None else None   # but it is part of the Rays state: copied
    n = pos.shape[0]
    ns = len(cfg) // SENSOR_CFG
    tf = table.f.detach()
    with torch.cuda.device(dev):
        cur, cp = torch.cuda.current_stream(dev), _copy_stream(dev)
        d_pos, d_dir, d_int = torch.empty((n, 3), **f32), torch.empty((n, 3), **f32), torch.empty(n, **f32)
        d_wav = torch.empty(n, **f32) if wav is not None else None
        opos, odir, oint = torch.empty((n, 3), **f32), torch.empty((n, 3), **f32), torch.empty(n, **f32)
        hitmask = torch.empty(n, dtype=torch.int64, device=dev)
        rec_on = bool(want_record and ns)
        records = torch.zeros((ns, n, 4), **f32) if rec_on else torch.empty((0, n, 4), **f32)
        images = torch.zeros(_image_numel(cfg), **f32)
        req = _table_req(tf, table.i, table.lut, table.lut_wavelengths)
        cp.wait_stream(cur)                       # the fresh buffers may reuse blocks still in flight on `cur`
        chunk = max(1, int(chunk_rays))
        ready = []
        with torch.cuda.stream(cp):
            for start in range(0, n, chunk):
                sl = slice(start, min(n, start + chunk))
                d_pos[sl].copy_(pos[sl], non_blocking=True)
                d_dir[sl].copy_(dir_[sl], non_blocking=True)
                d_int[sl].copy_(intensity[sl], non_blocking=True)
                if wav is not None:
                    d_wav[sl].copy_(wav[sl], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(cp)
                ready.append((sl, ev))
            d_id = None
            if ids is not None:                   # ray ids ride along behind the last chunk
                d_id = torch.empty(ids.shape, dtype=ids.dtype, device=dev)
                d_id.copy_(ids, non_blocking=True)
                d_id.record_stream(cur)
            tail = torch.cuda.Event()
            tail.record(cp)
        for sl, ev in ready:
            cur.wait_event(ev)
            m = sl.stop - sl.start
            sens, cnt = _sensor_reqs(cfg, m, [records[s_, sl] for s_ in range(ns)] if rec_on else None, images)
            lib.call("rtt_trace_seq_fwd", d_pos[sl].data_ptr(), d_dir[sl].data_ptr(), d_int[sl].data_ptr(),
                     d_wav[sl].data_ptr() if use_wav else 0, None,
                     opos[sl].data_ptr(), odir[sl].data_ptr(), oint[sl].data_ptr(), hitmask[sl].data_ptr(),
                     ct.byref(req), sens, cnt, m, _with_tune(mode, _tune_fwd), ct.c_void_p(cur.cuda_stream))
        cur.wait_event(tail)
    return dict(pos=opos, dir=odir, intensity=oint, hitmask=hitmask, records=records,
                images=split_images(images, cfg), in_pos=d_pos, in_dir=d_dir, in_intensity=d_int,
                in_wavelength=d_wav, in_id=d_id)


def trace_nonsequential(table: SurfaceTable, pos, dir_, intensity, nbounces: int, wavelength=None, *,
                        want_record=True, sensor_cfg: Optional[List[float]] = None, mode: Optional[int] = None,
                        record_depth: int = 1, source=None, want_rays: bool = True):
    """Fused Scene.simulate bounce loop (scene/base.py:129-235).

    Arithmetic: EXACT (the reference's rounding; default).  ``mode=MODE_FAST | MODE_NONSEQ_FAST`` opts in to the FAST
    arithmetic: hit sequences of rays that do not depend on the t > 1e-6 threshold are unchanged, the others — which
    are decided by fp32 rounding in the reference itself — may take another path.

    Returns dict(pos, dir, intensity, hit_seq [N,B] uint8 (255 = none), n_hits [N] uint8,
    records [n_sensors, K, N, 4] (the k-th interaction of ray i with the sensor, K = record_depth),
    sensor_counts [n_sensors, N] uint8 (interactions per ray, may exceed K), images)."""
    if not 0 <= nbounces <= C.MAX_BOUNCES:
        raise ValueError(f"nbounces must be in [0, {C.MAX_BOUNCES}]")
    cfg = sensor_cfg_of(table) if sensor_cfg is None else list(sensor_cfg)
    mode = _default_mode_nonseq if mode is None else mode
    if source is not None:
        _need_cuda(table.f)
        opos, odir, oint, seq, nh, records, images, counts = _TraceNonseqSrc.apply(
            source_cfg_of(source), source.pose, source.state, source.n, table.f, table.i, table.lut,
            table.lut_wavelengths, cfg, bool(want_record), bool(want_rays), int(nbounces), mode,
            max(1, int(record_depth)))
        return dict(pos=opos, dir=odir, intensity=oint, hit_seq=seq, n_hits=nh, records=records,
                    sensor_counts=counts, images=split_images(images, cfg))
    pos, dir_, intensity, wav = _prep_rays(pos, dir_, intensity, wavelength, table)
    opos, odir, oint, seq, nh, records, images, counts = _TraceNonseq.apply(
        pos, dir_, intensity, wav, table.f, table.i, table.lut, table.lut_wavelengths, cfg, bool(want_record),
        int(nbounces), mode, max(1, int(record_depth)))
    return dict(pos=opos, dir=odir, intensity=oint, hit_seq=seq, n_hits=nh, records=records,
                sensor_counts=counts, images=split_images(images, cfg))


def trace_nonsequential_host(table: SurfaceTable, pos, dir_, intensity, nbounces: int, wavelength=None, *,
                             want_record=False, sensor_cfg: Optional[List[float]] = None, mode: Optional[int] = None,
                             record_depth: int = 1, chunk_rays: int = 1 << 23, ids=None):
    """``trace_nonsequential`` for a bundle that lives in HOST memory (pinned for full copy speed): chunk k+1 is copied
    host->device on a side stream while chunk k runs its bounce loop (one ``rtt_trace_nonseq_fwd`` launch per chunk,
    all accumulating into the same sensor images).  Forward only, like ``trace_sequential_host``; same dict as
    ``trace_nonsequential`` plus the device copies of the inputs."""
    if not 0 <= nbounces <= C.MAX_BOUNCES:
        raise ValueError(f"nbounces must be in [0, {C.MAX_BOUNCES}]")
    _need_cuda(table.f)
    for t in (pos, dir_, intensity):
        if t.is_cuda:
            raise ValueError("trace_nonsequential_host takes host tensors; use trace_nonsequential for device rays")
    dev = table.f.device
    lib = _cabi.load()
    cfg = sensor_cfg_of(table) if sensor_cfg is None else list(sensor_cfg)
    mode = _default_mode_nonseq if mode is None else mode
    f32 = dict(dtype=torch.float32, device=dev)
    pos, dir_, intensity = (_f32c(t.detach()) for t in (pos, dir_, intensity))
    use_wav = table.lut is not None and wavelength is not None
    wav = _f32c(wavelength.detach()) if wavelength is not None else None
    n, nb = pos.shape[0], int(nbounces)
    ns = len(cfg) // SENSOR_CFG
    K = max(1, int(record_depth))
    tf = table.f.detach()
    with torch.cuda.device(dev):
        cur, cp = torch.cuda.current_stream(dev), _copy_stream(dev)
        d_pos, d_dir, d_int = torch.empty((n, 3), **f32), torch.empty((n, 3), **f32), torch.empty(n, **f32)
        d_wav = torch.empty(n, **f32) if wav is not None else None
        opos, odir, oint = torch.empty((n, 3), **f32), torch.empty((n, 3), **f32), torch.empty(n, **f32)
        seq = torch.empty((n, nb), dtype=torch.uint8, device=dev)
        nh = torch.empty(n, dtype=torch.uint8, device=dev)
        rec_on = bool(want_record and ns)
        counts = torch.zeros((ns if rec_on else 0, n), dtype=torch.uint8, device=dev)
        images = torch.zeros(_image_numel(cfg), **f32)
        req = _table_req(tf, table.i, table.lut, table.lut_wavelengths)
        cp.wait_stream(cur)
        chunk = max(1, int(chunk_rays))
        ready, rec_chunks = [], []
        with torch.cuda.stream(cp):
            for start in range(0, n, chunk):
                sl = slice(start, min(n, start + chunk))
                d_pos[sl].copy_(pos[sl], non_blocking=True)
                d_dir[sl].copy_(dir_[sl], non_blocking=True)
                d_int[sl].copy_(intensity[sl], non_blocking=True)
                if wav is not None:
                    d_wav[sl].copy_(wav[sl], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(cp)
                ready.append((sl, ev))
            d_id = None
            if ids is not None:
                d_id = torch.empty(ids.shape, dtype=ids.dtype, device=dev)
                d_id.copy_(ids, non_blocking=True)
                d_id.record_stream(cur)
            tail = torch.cuda.Event()
            tail.record(cp)
        for sl, ev in ready:
            cur.wait_event(ev)
            m = sl.stop - sl.start
            # the kernel addresses records as [K, m, 4] of ITS launch: one buffer per chunk, joined below
            rec_c = torch.zeros((ns, K, m, 4), **f32) if rec_on else None
            if rec_on:
                rec_chunks.append(rec_c)
            sens, cnt = _sensor_reqs(cfg, m, rec_c, images, [counts[s_, sl] for s_ in range(ns)] if rec_on else None, K)
            lib.call("rtt_trace_nonseq_fwd", d_pos[sl].data_ptr(), d_dir[sl].data_ptr(), d_int[sl].data_ptr(),
                     d_wav[sl].data_ptr() if use_wav else 0, None,
                     opos[sl].data_ptr(), odir[sl].data_ptr(), oint[sl].data_ptr(), seq[sl].data_ptr(),
                     nh[sl].data_ptr(), ct.byref(req), sens, cnt, nb, m, mode, ct.c_void_p(cur.cuda_stream))
        cur.wait_event(tail)
        records = torch.cat(rec_chunks, dim=2) if rec_chunks else torch.empty((0, K, n, 4), **f32)
    return dict(pos=opos, dir=odir, intensity=oint, hit_seq=seq, n_hits=nh, records=records, sensor_counts=counts,
                images=split_images(images, cfg), in_pos=d_pos, in_dir=d_dir, in_intensity=d_int,
                in_wavelength=d_wav, in_id=d_id)


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
