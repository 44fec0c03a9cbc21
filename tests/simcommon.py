"""Back-end independent drivers of the ray-source and goal-reduction entry points (shared by
tests/hostsim and tests/gpusim; the back-end supplies array allocation and pointer access)."""
import ctypes as ct

import numpy as np

from raytracetorch_b200 import _cabi, codes as C


class SourceGoalMixin:
    # hooks: _a(np) -> backend array, _z(shape, dtype) -> zeros, _pp(x) -> address, _host(x) -> numpy,
    #        _mode, _st() -> stream, _sync()

    def _source(self, src):
        pose = self._a(np.asarray(src["pose"], np.float32))
        state = None
        if src.get("state") is not None:
            state = self._a(np.asarray(src["state"], np.int64))
        req = _cabi.make_source(src["kind"], src["a"], self._pp(pose), seed=src.get("seed", 0),
                                first=src.get("first", 0), state_ptr=self._pp(state) if state is not None else 0,
                                width=src.get("width", 0), height=src.get("height", 0),
                                intensity=src.get("intensity", 1.0), wavelength=src.get("wavelength", 0.0))
        return req, (pose, state)

    def sample(self, src, n):
        req, hold = self._source(src)
        pos, dir_ = self._z((n, 3), np.float32), self._z((n, 3), np.float32)
        inten, wav = self._z(n, np.float32), self._z(n, np.float32)
        self.lib.call("rtt_sample_bundle", ct.byref(req), self._pp(pos), self._pp(dir_), self._pp(inten), self._pp(wav),
                      n, self._mode, self._st())
        self._sync()
        return dict(pos=self._host(pos), dir=self._host(dir_), intensity=self._host(inten), wavelength=self._host(wav))

    def trace_seq_src(self, tf, ti, src, n, sensor_specs=None, want_rays=True, lut=None, lut_w=None):
        req, hold = self._table(tf, ti, lut, lut_w)
        sreq, shold = self._source(src)
        sens, ns, keep = self._sensors(n, sensor_specs)
        op = self._z((n, 3), np.float32) if want_rays else None
        od = self._z((n, 3), np.float32) if want_rays else None
        oi = self._z(n, np.float32) if want_rays else None
        mask = self._z(n, np.int64)
        self.lib.call("rtt_trace_seq_fwd", None, None, None, None, ct.byref(sreq), self._pp(op), self._pp(od),
                      self._pp(oi), self._pp(mask), ct.byref(req), sens, ns, n, self._mode, self._st())
        self._sync()
        return dict(pos=self._host(op), dir=self._host(od), intensity=self._host(oi),
                    hitmask=self._host(mask).view(np.uint64),
                    sensors=[tuple(self._host(x) for x in k) for k in keep])

    def trace_seq_src_bwd(self, tf, ti, src, n, mask, g_records, lut=None, lut_w=None):
        req, hold = self._table(tf, ti, lut, lut_w)
        sreq, shold = self._source(src)
        mask = self._a(np.asarray(mask).view(np.int64))
        gt = self._z((req.n_rows, C.ROW_G), np.float32)
        g_records = [self._a(np.asarray(g, np.float32)) for g in g_records]
        ns = len(g_records)
        rec_arr = (ct.c_void_p * ns)(*[self._pp(g) or None for g in g_records]) if ns else None
        self.lib.call("rtt_trace_seq_bwd", None, None, None, None, ct.byref(sreq), self._pp(mask), None, None, None,
                      rec_arr, None, None, None, self._pp(gt), None, ct.byref(req), ns, n, self._mode, self._st())
        self._sync()
        return dict(g_table=self._host(gt))

    # ---- goal reductions ----
    def spot_moments(self, rec, active_only):
        rec = self._a(np.asarray(rec, np.float32))
        out, work = self._z(4, np.float32), self._z(C.SPOT_WORK, np.float32)
        self.lib.call("rtt_spot_moments", self._pp(rec), rec.shape[0], int(active_only), self._pp(out), self._pp(work),
                      self._st())
        self._sync()
        return self._host(out)

    def spot_moments_bwd(self, rec, active_only, g3):
        rec, g3 = self._a(np.asarray(rec, np.float32)), self._a(np.asarray(g3, np.float32))
        g = self._z(rec.shape, np.float32)
        self.lib.call("rtt_spot_moments_bwd", self._pp(rec), rec.shape[0], int(active_only), self._pp(g3), self._pp(g),
                      self._st())
        self._sync()
        return self._host(g)

    def spot_size(self, rec, mom4, target=None, g_loss=1.0, repeat=1):
        rec, mom4 = self._a(np.asarray(rec, np.float32)), self._a(np.asarray(mom4, np.float32))
        tgt = None if target is None else self._a(np.asarray(target, np.float32))
        out, work = self._z(3, np.float32), self._z(C.SPOT_WORK, np.float32)
        for _ in range(repeat):          # the kernels must leave `work` reusable
            self.lib.call("rtt_spot_size_fwd", self._pp(rec), rec.shape[0], self._pp(mom4), self._pp(tgt), self._pp(out),
                          self._pp(work), self._st())
        gl = self._a(np.asarray([g_loss], np.float32))
        g = self._z(rec.shape, np.float32)
        self.lib.call("rtt_spot_size_bwd", self._pp(rec), rec.shape[0], self._pp(mom4), self._pp(tgt), self._pp(out),
                      self._pp(gl), self._pp(g), self._st())
        self._sync()
        return self._host(out), self._host(g)

    # ---- per-id sensor moments ----
    def spot_id(self, rec, ids, query, targets=None, p=2.0, g_out=None, repeat=1):
        """(moments [K,4], sums [K,4], result [K], g_rec [M,4]) through rtt_spot_id_moments / _size / _size_bwd, with
        the same host-side glue as raytracetorch_b200.ops._SpotSizePerId."""
        rec_np = np.asarray(rec, np.float32)
        K = len(query)
        lut = np.full(256, -1, np.int32)
        lut[np.asarray(query, np.int64) + 128] = np.arange(K, dtype=np.int32)
        rec_d, ids_d, lut_d = self._a(rec_np), self._a(np.asarray(ids, np.int8)), self._a(lut)
        mom, work = self._z((K, 4), np.float32), self._z(_cabi.SPOT_ID_WORK, np.float32)
        for _ in range(repeat):          # the kernels must leave `work` reusable
            self.lib.call("rtt_spot_id_moments", self._pp(rec_d), self._pp(ids_d), rec_np.shape[0], self._pp(lut_d), K,
                          self._pp(mom), self._pp(work), self._st())
        self._sync()
        mom_h = self._host(mom)
        W = mom_h[:, 0]
        safe = np.where(W == 0, 1.0, W).astype(np.float32)
        centres = (mom_h[:, 1:3] / safe[:, None]) if targets is None else np.asarray(targets, np.float32)
        cen_d = self._a(np.ascontiguousarray(centres, np.float32))
        s4 = self._z((K, 4), np.float32)
        self.lib.call("rtt_spot_id_size", self._pp(rec_d), self._pp(ids_d), rec_np.shape[0], self._pp(lut_d), K,
                      self._pp(cen_d), float(p), self._pp(s4), self._pp(work), self._st())
        self._sync()
        s4_h = self._host(s4)
        result = s4_h[:, 0] / (2.0 * safe)
        g_rec = None
        if g_out is not None:
            a = np.asarray(g_out, np.float32) / (2.0 * safe)
            hit = (W != 0).astype(np.float32)
            free = targets is None
            bx = s4_h[:, 1] / safe * hit if free else np.zeros(K, np.float32)
            by = s4_h[:, 2] / safe * hit if free else np.zeros(K, np.float32)
            coef = np.stack([centres[:, 0], centres[:, 1], a, bx, by, s4_h[:, 0] / safe * hit, np.zeros(K), np.zeros(K)],
                            1).astype(np.float32)
            coef_d = self._a(np.ascontiguousarray(coef))
            g = self._z(rec_np.shape, np.float32)
            self.lib.call("rtt_spot_id_size_bwd", self._pp(rec_d), self._pp(ids_d), rec_np.shape[0], self._pp(lut_d), K,
                          self._pp(coef_d), float(p), self._pp(g), self._st())
            self._sync()
            g_rec = self._host(g)
        return mom_h, s4_h, result, g_rec

    # ---- Renderer.render_3d in one call ----
    def render_shade(self, tf, ti, pos, dir_, base_rgb, light, background):
        req, hold = self._table(tf, ti, None, None)
        n = pos.shape[0]
        pos_d, dir_d = self._a(np.asarray(pos, np.float32)), self._a(np.asarray(dir_, np.float32))
        base_d = self._a(np.ascontiguousarray(base_rgb, np.float32))
        rgb, row = self._z((n, 3), np.float32), self._z(n, np.uint8)
        la, bg = (ct.c_float * 3)(*map(float, light)), (ct.c_float * 3)(*map(float, background))
        self.lib.call("rtt_render_shade", self._pp(pos_d), self._pp(dir_d), None, ct.byref(req), self._pp(base_d), la, bg,
                      self._pp(rgb), self._pp(row), n, 1, self._st())
        self._sync()
        return self._host(rgb), self._host(row)
