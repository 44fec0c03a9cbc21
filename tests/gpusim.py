"""TEST INFRASTRUCTURE: drive the CUDA library through its C ABI with numpy in / numpy out.

Same method names and return layout as tests/hostsim.HostSim, so CPU and GPU parity tests share
their comparison code.  Buffers are torch CUDA tensors (device memory + stream plumbing only);
every compute call goes through ctypes into raytracetorch_b200/librtt_b200.so.
"""
from __future__ import annotations

import ctypes as ct

import numpy as np
import torch

from raytracetorch_b200 import _cabi, codes as C
from simcommon import SourceGoalMixin


def _dev(a, dtype=torch.float32):
    if a is None:
        return None
    return torch.as_tensor(np.ascontiguousarray(a)).to(device="cuda", dtype=dtype).contiguous()


def _p(t):
    return 0 if t is None or t.numel() == 0 else t.data_ptr()


_TORCH_DT = {np.dtype(np.float32): torch.float32, np.dtype(np.int64): torch.int64, np.dtype(np.uint8): torch.uint8,
             np.dtype(np.int32): torch.int32}


class GpuSim(SourceGoalMixin):
    def __init__(self, mode: int):
        self.lib = _cabi.load()
        self.mode = self._mode = mode

    # hooks of SourceGoalMixin
    @staticmethod
    def _a(x):
        x = np.ascontiguousarray(x)
        return torch.as_tensor(x).to("cuda").contiguous()

    @staticmethod
    def _z(shape, dtype):
        return torch.zeros(shape, dtype=_TORCH_DT[np.dtype(dtype)], device="cuda")

    _pp = staticmethod(lambda t: 0 if t is None or t.numel() == 0 else t.data_ptr())
    _host = staticmethod(lambda t: None if t is None else t.cpu().numpy())

    def _st(self):
        return self._stream()

    _sync = staticmethod(lambda: torch.cuda.synchronize())

    def _stream(self):
        return ct.c_void_p(torch.cuda.current_stream().cuda_stream)

    def _table(self, tf, ti, lut, lut_w):
        tf, ti = _dev(tf), _dev(ti, torch.int32)
        lut, lut_w = _dev(lut), _dev(lut_w)
        n_lut = 0 if lut is None else lut.shape[0]
        return _cabi.make_table(_p(tf), _p(ti), tf.shape[0], _p(lut), _p(lut_w), n_lut), (tf, ti, lut, lut_w)

    def _sensors(self, n, specs, want_record=True, K=1, want_count=False):
        reqs, keep = [], []
        for sp in specs or []:
            rec = torch.zeros((K, n, 4) if K > 1 else (n, 4), device="cuda") if want_record else None
            cnt_a = torch.zeros(n, dtype=torch.uint8, device="cuda") if want_count else None
            img = None
            r = dict(record=_p(rec), record_hits=K, count=_p(cnt_a))
            if sp is not None:
                H, W, x0, x1, y0, y1, ch = sp
                img = torch.zeros((ch, H, W), device="cuda")
                r.update(image=_p(img), height=H, width=W, channels=ch, x0=x0, y0=y0,
                         sx=float(np.float32(W / (x1 - x0))), sy=float(np.float32(H / (y1 - y0))))
            reqs.append(r)
            keep.append((rec, img, cnt_a))
        arr, cnt = _cabi.make_sensors(reqs)
        return arr, cnt, keep

    @staticmethod
    def _np(t):
        return None if t is None else t.cpu().numpy()

    def trace_seq(self, tf, ti, pos, dir_, inten, wav=None, lut=None, lut_w=None, sensor_specs=None):
        pos, dir_, inten, wav = _dev(pos), _dev(dir_), _dev(inten), _dev(wav)
        n = pos.shape[0]
        req, hold = self._table(tf, ti, lut, lut_w)
        sens, ns, keep = self._sensors(n, sensor_specs)
        op, od, oi = torch.empty_like(pos), torch.empty_like(dir_), torch.empty_like(inten)
        mask = torch.zeros(n, dtype=torch.int64, device="cuda")
        self.lib.call("rtt_trace_seq_fwd", _p(pos), _p(dir_), _p(inten), _p(wav), None, _p(op), _p(od), _p(oi), _p(mask),
                      ct.byref(req), sens, ns, n, self.mode, self._stream())
        torch.cuda.synchronize()
        return dict(pos=self._np(op), dir=self._np(od), intensity=self._np(oi),
                    hitmask=self._np(mask).view(np.uint64),
                    sensors=[(self._np(r), self._np(i), self._np(c)) for r, i, c in keep])

    def trace_seq_bwd(self, tf, ti, pos, dir_, inten, mask, g_pos, g_dir, g_int, wav=None, lut=None, lut_w=None,
                      g_records=None, hint=0, need_rays=True):
        pos, dir_, inten, wav = _dev(pos), _dev(dir_), _dev(inten), _dev(wav)
        g_pos, g_dir, g_int = _dev(g_pos), _dev(g_dir), _dev(g_int)
        mask = _dev(np.asarray(mask).view(np.int64), torch.int64)
        n = pos.shape[0]
        req, hold = self._table(tf, ti, lut, lut_w)
        gp, gd, gi = torch.zeros_like(pos), torch.zeros_like(dir_), torch.zeros_like(inten)
        gt = torch.zeros((req.n_rows, C.ROW_G), device="cuda")
        gl = None if lut is None else torch.zeros_like(hold[2])
        g_records = [_dev(g) for g in (g_records or [])]
        ns = len(g_records)
        rec_arr = (ct.c_void_p * ns)(*[_p(g) or None for g in g_records]) if ns else None
        self.lib.call("rtt_trace_seq_bwd", _p(pos), _p(dir_), _p(inten), _p(wav), None, _p(mask),
                      _p(g_pos), _p(g_dir), _p(g_int), rec_arr,
                      _p(gp) if need_rays else 0, _p(gd) if need_rays else 0, _p(gi) if need_rays else 0, _p(gt), _p(gl),
                      ct.byref(req), ns, n, self.mode | hint, self._stream())
        torch.cuda.synchronize()
        return dict(g_pos=self._np(gp), g_dir=self._np(gd), g_intensity=self._np(gi), g_table=self._np(gt),
                    g_lut=self._np(gl))

    def trace_nonseq(self, tf, ti, pos, dir_, inten, nbounces, wav=None, lut=None, lut_w=None, sensor_specs=None,
                     record_hits=1):
        pos, dir_, inten, wav = _dev(pos), _dev(dir_), _dev(inten), _dev(wav)
        n = pos.shape[0]
        req, hold = self._table(tf, ti, lut, lut_w)
        sens, ns, keep = self._sensors(n, sensor_specs, K=record_hits, want_count=True)
        op, od, oi = torch.empty_like(pos), torch.empty_like(dir_), torch.empty_like(inten)
        seq = torch.zeros((n, nbounces), dtype=torch.uint8, device="cuda")
        nh = torch.zeros(n, dtype=torch.uint8, device="cuda")
        self.lib.call("rtt_trace_nonseq_fwd", _p(pos), _p(dir_), _p(inten), _p(wav), None, _p(op), _p(od), _p(oi),
                      _p(seq), _p(nh), ct.byref(req), sens, ns, nbounces, n, self.mode, self._stream())
        torch.cuda.synchronize()
        return dict(pos=self._np(op), dir=self._np(od), intensity=self._np(oi), seq=self._np(seq), nb=self._np(nh),
                    sensors=[(self._np(r), self._np(i), self._np(c)) for r, i, c in keep])

    def trace_nonseq_bwd(self, tf, ti, pos, dir_, inten, seq, g_pos, g_dir, g_int, wav=None, lut=None, lut_w=None,
                         g_records=None, record_hits=1):
        pos, dir_, inten, wav = _dev(pos), _dev(dir_), _dev(inten), _dev(wav)
        g_pos, g_dir, g_int = _dev(g_pos), _dev(g_dir), _dev(g_int)
        seq = _dev(seq, torch.uint8)
        n = pos.shape[0]
        req, hold = self._table(tf, ti, lut, lut_w)
        gp, gd, gi = torch.zeros_like(pos), torch.zeros_like(dir_), torch.zeros_like(inten)
        gt = torch.zeros((req.n_rows, C.ROW_G), device="cuda")
        gl = None if lut is None else torch.zeros_like(hold[2])
        g_records = [_dev(g) for g in (g_records or [])]
        ns = len(g_records)
        rec_arr = (ct.c_void_p * ns)(*[_p(g) or None for g in g_records]) if ns else None
        hits = (ct.c_int32 * max(ns, 1))(*([record_hits] * max(ns, 1)))
        self.lib.call("rtt_trace_nonseq_bwd", _p(pos), _p(dir_), _p(inten), _p(wav), None, _p(seq), seq.shape[1],
                      _p(g_pos), _p(g_dir), _p(g_int), rec_arr, hits, _p(gp), _p(gd), _p(gi), _p(gt), _p(gl),
                      ct.byref(req), ns, n, self.mode, self._stream())
        torch.cuda.synchronize()
        return dict(g_pos=self._np(gp), g_dir=self._np(gd), g_intensity=self._np(gi), g_table=self._np(gt),
                    g_lut=self._np(gl))

    def intersect_test(self, tf, ti, pos, dir_, row0, k):
        pos, dir_ = _dev(pos), _dev(dir_)
        n = pos.shape[0]
        req, hold = self._table(tf, ti, None, None)
        t = torch.empty((n, k), device="cuda")
        self.lib.call("rtt_intersect_test", _p(pos), _p(dir_), _p(t), ct.byref(req), row0, k, n, self.mode,
                      self._stream())
        torch.cuda.synchronize()
        return self._np(t)

    def surface_step(self, tf, ti, pos, dir_, row, wav=None, lut=None, lut_w=None):
        pos, dir_, wav = _dev(pos), _dev(dir_), _dev(wav)
        n = pos.shape[0]
        req, hold = self._table(tf, ti, lut, lut_w)
        npos, ndir, hl, nrm = (torch.empty_like(pos) for _ in range(4))
        mod, t = torch.empty(n, device="cuda"), torch.empty(n, device="cuda")
        self.lib.call("rtt_surface_step_fwd", _p(pos), _p(dir_), _p(wav), _p(npos), _p(ndir), _p(mod), _p(hl), _p(t),
                      _p(nrm), ct.byref(req), row, n, self.mode, self._stream())
        torch.cuda.synchronize()
        return dict(pos=self._np(npos), dir=self._np(ndir), mod=self._np(mod), hit_local=self._np(hl), t=self._np(t),
                    normal=self._np(nrm))

    def surface_step_bwd(self, tf, ti, pos, dir_, row, g_npos=None, g_ndir=None, g_hl=None, g_t=None, g_n=None,
                         wav=None, lut=None, lut_w=None):
        pos, dir_, wav = _dev(pos), _dev(dir_), _dev(wav)
        g_npos, g_ndir, g_hl, g_t, g_n = (_dev(x) for x in (g_npos, g_ndir, g_hl, g_t, g_n))
        n = pos.shape[0]
        req, hold = self._table(tf, ti, lut, lut_w)
        gp, gd = torch.zeros_like(pos), torch.zeros_like(dir_)
        gt = torch.zeros((req.n_rows, C.ROW_G), device="cuda")
        gl = None if lut is None else torch.zeros_like(hold[2])
        self.lib.call("rtt_surface_step_bwd", _p(pos), _p(dir_), _p(wav), _p(g_npos), _p(g_ndir), _p(g_hl), _p(g_t),
                      _p(g_n), _p(gp), _p(gd), _p(gt), _p(gl), ct.byref(req), row, n, self.mode, self._stream())
        torch.cuda.synchronize()
        return dict(g_pos=self._np(gp), g_dir=self._np(gd), g_table=self._np(gt), g_lut=self._np(gl))
