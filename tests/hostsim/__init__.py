"""TEST INFRASTRUCTURE: host build of the kernels' per-ray arithmetic (see hostsim.cpp).

``build(variant)`` compiles ``hostsim.cpp`` (which includes the product's
``csrc/rtt_core.cuh``) with g++ into ``tests/hostsim/_build/`` and returns a numpy-driven
runner with the same entry points as the CUDA library.  Variants: ``exact``
(``-ffp-contract=off``: separately rounded mul/add, like ``nvcc -fmad=false``) and ``fast``
(``-ffp-contract=fast -mfma``: contraction like nvcc's default).  Used by the CPU tests to
validate forward and adjoint arithmetic against the torch oracle without a GPU.
"""
from __future__ import annotations

import ctypes as ct
import os
import subprocess

import numpy as np

from raytracetorch_b200 import _cabi, codes as C
from simcommon import SourceGoalMixin

_HERE = os.path.dirname(os.path.abspath(__file__))
_BUILD = os.path.join(_HERE, "_build")
_FLAGS = {"exact": ["-ffp-contract=off", "-mfma"],
          "fast": ["-ffp-contract=fast", "-mfma", "-DRTT_HOST_TILE", "-DRTT_TILE_LEAN"],
          # mirror of k_trace_seq_fwd_pair (csrc/rtt_pair.cuh): packed two-ray arithmetic, lanes = fmaf on the host
          "pair": ["-ffp-contract=fast", "-mfma", "-DRTT_HOST_TILE", "-DRTT_TILE_LEAN", "-DRTT_HOST_PAIR"]}
_cache = {}


def build(variant: str = "exact") -> "HostSim":
    if variant in _cache:
        return _cache[variant]
    os.makedirs(_BUILD, exist_ok=True)
    out = os.path.join(_BUILD, f"librtt_hostsim_{variant}.so")
    src = os.path.join(_HERE, "hostsim.cpp")
    core = os.path.join(_HERE, "..", "..", "raytracetorch_b200", "csrc", "rtt_core.cuh")
    tile = os.path.join(_HERE, "..", "..", "raytracetorch_b200", "csrc", "rtt_tile.cuh")
    pair = os.path.join(_HERE, "..", "..", "raytracetorch_b200", "csrc", "rtt_pair.cuh")
    lean = os.path.join(_HERE, "..", "..", "raytracetorch_b200", "csrc", "rtt_lean.cuh")
    hdr = os.path.join(_HERE, "..", "..", "include", "rtt_b200.h")
    newest = max(os.path.getmtime(p) for p in (src, core, tile, pair, lean, hdr, __file__))
    if not os.path.exists(out) or os.path.getmtime(out) < newest:
        cmd = ["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-fno-fast-math", *_FLAGS[variant], src, "-o", out]
        subprocess.run(cmd, check=True)
    _cache[variant] = HostSim(_cabi.RttLib(out))
    return _cache[variant]


def _p(a):
    return 0 if a is None else a.ctypes.data


def _f32(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float32)


class HostSim(SourceGoalMixin):
    def __init__(self, lib: _cabi.RttLib):
        self.lib = lib
        self._mode = 0

    # hooks of SourceGoalMixin
    _a = staticmethod(lambda x: np.ascontiguousarray(x))
    _z = staticmethod(lambda shape, dtype: np.zeros(shape, dtype))
    _pp = staticmethod(lambda x: 0 if x is None else x.ctypes.data)
    _host = staticmethod(lambda x: x)
    _st = staticmethod(lambda: None)
    _sync = staticmethod(lambda: None)

    @staticmethod
    def _table(tf, ti, lut, lut_w):
        tf, ti = _f32(tf), np.ascontiguousarray(ti, dtype=np.int32)
        lut, lut_w = _f32(lut), _f32(lut_w)
        n_lut = 0 if lut is None else lut.shape[0]
        req = _cabi.make_table(_p(tf), _p(ti), tf.shape[0], _p(lut), _p(lut_w), n_lut)
        return req, (tf, ti, lut, lut_w)

    @staticmethod
    def _sensors(n, specs, want_record=True, K=1, want_count=False):
        """specs: list of (H, W, x0, x1, y0, y1, channels) or None per slot.
        Returns per slot (record [n,4] (K == 1) or [K,n,4], image, count [n] or None)."""
        reqs, keep = [], []
        for sp in specs or []:
            rec = np.zeros((K, n, 4) if K > 1 else (n, 4), np.float32) if want_record else None
            cnt_a = np.zeros(n, np.uint8) if want_count else None
            img = None
            r = dict(record=_p(rec), record_hits=K, count=_p(cnt_a))
            if sp is not None:
                H, W, x0, x1, y0, y1, ch = sp
                img = np.zeros((ch, H, W), np.float32)
                r.update(image=_p(img), height=H, width=W, channels=ch, x0=x0, y0=y0,
                         sx=float(np.float32(W / (x1 - x0))), sy=float(np.float32(H / (y1 - y0))))
            reqs.append(r)
            keep.append((rec, img, cnt_a))
        arr, cnt = _cabi.make_sensors(reqs)
        return arr, cnt, keep

    def trace_seq(self, tf, ti, pos, dir_, inten, wav=None, lut=None, lut_w=None, sensor_specs=None):
        pos, dir_, inten, wav = _f32(pos), _f32(dir_), _f32(inten), _f32(wav)
        n = pos.shape[0]
        req, hold = self._table(tf, ti, lut, lut_w)
        sens, ns, keep = self._sensors(n, sensor_specs)
        op, od, oi = np.empty_like(pos), np.empty_like(dir_), np.empty_like(inten)
        mask = np.zeros(n, np.uint64)
        self.lib.call("rtt_trace_seq_fwd", _p(pos), _p(dir_), _p(inten), _p(wav), None, _p(op), _p(od), _p(oi), _p(mask),
                      ct.byref(req), sens, ns, n, 0, None)
        return dict(pos=op, dir=od, intensity=oi, hitmask=mask, sensors=keep)

    def trace_seq_bwd(self, tf, ti, pos, dir_, inten, mask, g_pos, g_dir, g_int, wav=None, lut=None, lut_w=None,
                      g_records=None, hint=0, need_rays=True):
        pos, dir_, inten, wav = _f32(pos), _f32(dir_), _f32(inten), _f32(wav)
        g_pos, g_dir, g_int = _f32(g_pos), _f32(g_dir), _f32(g_int)
        n = pos.shape[0]
        req, hold = self._table(tf, ti, lut, lut_w)
        gp, gd, gi = np.zeros_like(pos), np.zeros_like(dir_), np.zeros_like(inten)
        gt = np.zeros((req.n_rows, C.ROW_G), np.float32)
        gl = None if lut is None else np.zeros_like(hold[2])
        g_records = [_f32(g) for g in (g_records or [])]
        ns = len(g_records)
        rec_arr = (ct.c_void_p * max(ns, 1))(*[_p(g) or None for g in g_records]) if ns else None
        self.lib.call("rtt_trace_seq_bwd", _p(pos), _p(dir_), _p(inten), _p(wav), None, _p(mask),
                      _p(g_pos), _p(g_dir), _p(g_int), rec_arr,
                      _p(gp) if need_rays else 0, _p(gd) if need_rays else 0, _p(gi) if need_rays else 0, _p(gt), _p(gl),
                      ct.byref(req), ns, n, hint, None)
        return dict(g_pos=gp, g_dir=gd, g_intensity=gi, g_table=gt, g_lut=gl)

    def trace_nonseq(self, tf, ti, pos, dir_, inten, nbounces, wav=None, lut=None, lut_w=None, sensor_specs=None,
                     record_hits=1):
        pos, dir_, inten, wav = _f32(pos), _f32(dir_), _f32(inten), _f32(wav)
        n = pos.shape[0]
        req, hold = self._table(tf, ti, lut, lut_w)
        sens, ns, keep = self._sensors(n, sensor_specs, K=record_hits, want_count=True)
        op, od, oi = np.empty_like(pos), np.empty_like(dir_), np.empty_like(inten)
        seq = np.zeros((n, nbounces), np.uint8)
        nh = np.zeros(n, np.uint8)
        self.lib.call("rtt_trace_nonseq_fwd", _p(pos), _p(dir_), _p(inten), _p(wav), None, _p(op), _p(od), _p(oi),
                      _p(seq), _p(nh), ct.byref(req), sens, ns, nbounces, n, 0, None)
        return dict(pos=op, dir=od, intensity=oi, seq=seq, nb=nh, sensors=keep)

    def trace_nonseq_bwd(self, tf, ti, pos, dir_, inten, seq, g_pos, g_dir, g_int, wav=None, lut=None, lut_w=None,
                         g_records=None, record_hits=1):
        pos, dir_, inten, wav = _f32(pos), _f32(dir_), _f32(inten), _f32(wav)
        g_pos, g_dir, g_int = _f32(g_pos), _f32(g_dir), _f32(g_int)
        n = pos.shape[0]
        req, hold = self._table(tf, ti, lut, lut_w)
        gp, gd, gi = np.zeros_like(pos), np.zeros_like(dir_), np.zeros_like(inten)
        gt = np.zeros((req.n_rows, C.ROW_G), np.float32)
        gl = None if lut is None else np.zeros_like(hold[2])
        seq = np.ascontiguousarray(seq, np.uint8)
        g_records = [_f32(g) for g in (g_records or [])]
        ns = len(g_records)
        rec_arr = (ct.c_void_p * max(ns, 1))(*[_p(g) or None for g in g_records]) if ns else None
        hits = (ct.c_int32 * max(ns, 1))(*([record_hits] * max(ns, 1)))
        self.lib.call("rtt_trace_nonseq_bwd", _p(pos), _p(dir_), _p(inten), _p(wav), None, _p(seq), seq.shape[1],
                      _p(g_pos), _p(g_dir), _p(g_int), rec_arr, hits, _p(gp), _p(gd), _p(gi), _p(gt), _p(gl),
                      ct.byref(req), ns, n, 0, None)
        return dict(g_pos=gp, g_dir=gd, g_intensity=gi, g_table=gt, g_lut=gl)

    def intersect_test(self, tf, ti, pos, dir_, row0, k):
        pos, dir_ = _f32(pos), _f32(dir_)
        n = pos.shape[0]
        req, hold = self._table(tf, ti, None, None)
        t = np.empty((n, k), np.float32)
        self.lib.call("rtt_intersect_test", _p(pos), _p(dir_), _p(t), ct.byref(req), row0, k, n, 0, None)
        return t

    def surface_step(self, tf, ti, pos, dir_, row, wav=None, lut=None, lut_w=None):
        pos, dir_, wav = _f32(pos), _f32(dir_), _f32(wav)
        n = pos.shape[0]
        req, hold = self._table(tf, ti, lut, lut_w)
        npos, ndir, hl, nrm = (np.empty_like(pos) for _ in range(4))
        mod, t = np.empty(n, np.float32), np.empty(n, np.float32)
        self.lib.call("rtt_surface_step_fwd", _p(pos), _p(dir_), _p(wav), _p(npos), _p(ndir), _p(mod), _p(hl), _p(t),
                      _p(nrm), ct.byref(req), row, n, 0, None)
        return dict(pos=npos, dir=ndir, mod=mod, hit_local=hl, t=t, normal=nrm)

    def surface_step_bwd(self, tf, ti, pos, dir_, row, g_npos=None, g_ndir=None, g_hl=None, g_t=None, g_n=None,
                         wav=None, lut=None, lut_w=None):
        pos, dir_, wav = _f32(pos), _f32(dir_), _f32(wav)
        g_npos, g_ndir, g_hl, g_t, g_n = (_f32(x) for x in (g_npos, g_ndir, g_hl, g_t, g_n))
        n = pos.shape[0]
        req, hold = self._table(tf, ti, lut, lut_w)
        gp, gd = np.zeros_like(pos), np.zeros_like(dir_)
        gt = np.zeros((req.n_rows, C.ROW_G), np.float32)
        gl = None if lut is None else np.zeros_like(hold[2])
        self.lib.call("rtt_surface_step_bwd", _p(pos), _p(dir_), _p(wav), _p(g_npos), _p(g_ndir), _p(g_hl), _p(g_t),
                      _p(g_n), _p(gp), _p(gd), _p(gt), _p(gl), ct.byref(req), row, n, 0, None)
        return dict(g_pos=gp, g_dir=gd, g_table=gt, g_lut=gl)
