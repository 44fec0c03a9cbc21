// hostsim.cpp — TEST INFRASTRUCTURE: runs the kernels' per-ray arithmetic on the host.
//
// The CUDA kernels cannot execute in the build container (no GPU).  This file compiles the
// very same per-ray source (raytracetorch_b200/csrc/rtt_core.cuh) with g++ and drives it
// with plain loops that mirror the kernel bodies of rtt_kernels.inl, so the CPU test-suite
// can check the kernel arithmetic (forward AND hand-written adjoint) against the torch
// oracle before any GPU time is spent.  It exports the same C signatures as
// include/rtt_b200.h but with HOST pointers.  It is built into tests/hostsim/ and is never
// loaded by the product package (raytracetorch_b200/_cabi.py only opens librtt_b200.so).
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../raytracetorch_b200/csrc/rtt_core.cuh"
#include "../../raytracetorch_b200/csrc/rtt_tile.cuh"
#include "../../raytracetorch_b200/csrc/rtt_pair.cuh"
#include "../../raytracetorch_b200/csrc/rtt_lean.cuh"

using namespace rtt;

namespace {

struct HostTable {
    std::vector<RowDev> rows;
    int S = 0, L = 0;
    std::vector<float> ni, no, mu_enter, mu_exit, lut_w;
};

HostTable stage(const rtt_table_t* t) {
    HostTable T;
    T.S = t->n_rows; T.L = t->n_lut;
    T.rows.resize(T.S);
    for (int r = 0; r < T.S; ++r) {
        std::memset(&T.rows[r], 0, sizeof(RowDev));
        std::memcpy(T.rows[r].f, t->f + (size_t)r * RTT_ROW_F, sizeof(float) * RTT_ROW_F);
        std::memcpy(T.rows[r].i, t->i + (size_t)r * RTT_ROW_I, sizeof(int32_t) * RTT_ROW_I);
        prepare_row(T.rows[r]);
    }
    const int LS = T.L * T.S;
    T.ni.resize(LS); T.no.resize(LS); T.mu_enter.resize(LS); T.mu_exit.resize(LS);
    for (int idx = 0; idx < LS; ++idx) {
        T.ni[idx] = t->lut[2 * idx]; T.no[idx] = t->lut[2 * idx + 1];
        T.mu_enter[idx] = T.no[idx] / T.ni[idx]; T.mu_exit[idx] = T.ni[idx] / T.no[idx];
    }
    T.lut_w.assign(t->lut_w, t->lut_w + T.L);
    return T;
}

int lam_index(const HostTable& T, float w) {
    int best = 0;
    float bd = std::fabs(w - T.lut_w[0]);
    for (int l = 1; l < T.L; ++l) {
        const float dd = std::fabs(w - T.lut_w[l]);
        if (dd < bd) { bd = dd; best = l; }
    }
    return best;
}

struct Ior { float ni, no, mu_enter, mu_exit; };

Ior row_ior(const HostTable& T, int r, int lam) {
    Ior q;
    if (T.L > 0) {
        const int idx = lam * T.S + r;
        q.ni = T.ni[idx]; q.no = T.no[idx]; q.mu_enter = T.mu_enter[idx]; q.mu_exit = T.mu_exit[idx];
    } else {
        const RowDev& R = T.rows[r];
        q.ni = R.f[RTT_F_IOR_IN]; q.no = R.f[RTT_F_IOR_OUT]; q.mu_enter = R.f[D_MU_ENTER]; q.mu_exit = R.f[D_MU_EXIT];
    }
    return q;
}

V3 load3(const float* a, int64_t i) { return v3(a[3 * i], a[3 * i + 1], a[3 * i + 2]); }
void store3(float* a, int64_t i, V3 v) { a[3 * i] = v.x; a[3 * i + 1] = v.y; a[3 * i + 2] = v.z; }

void deposit(const rtt_sensor_t& sd, int64_t i, V3 hl, float w, int lam, int ordinal = 0, int64_t n = 0) {
    const int K = sd.record_hits < 1 ? 1 : sd.record_hits;
    if (sd.record && ordinal < K) {
        float* r = sd.record + 4 * ((int64_t)ordinal * n + i);
        r[0] = hl.x; r[1] = hl.y; r[2] = hl.z; r[3] = w;
    }
    if (sd.image && w != 0.0f) {
        int ix, iy;
        if (sensor_bin(hl.x, hl.y, sd.x0, sd.y0, sd.sx, sd.sy, sd.width, sd.height, &ix, &iy)) {
            const int ch = (sd.channels > 1) ? (lam < sd.channels - 1 ? lam : sd.channels - 1) : 0;
            sd.image[((size_t)ch * sd.height + iy) * sd.width + ix] += w;
        }
    }
}

void add_row_grad(const RowGrad& G, int flags, float* acc_row) {
    auto span = [&](int lo, int hi) { for (int e = lo; e < hi; ++e) acc_row[e] += G.g[e]; };
    if (flags & RTT_FLAG_GRAD_POSE_E) span(RTT_F_RE, RTT_F_TE + 3);
    if (flags & RTT_FLAG_GRAD_POSE_S) span(RTT_F_RS, RTT_F_TS + 3);
    if (flags & RTT_FLAG_GRAD_CK) span(RTT_F_C, RTT_F_K + 1);
    if (flags & RTT_FLAG_GRAD_RADIUS) span(RTT_F_RADIUS, RTT_F_RADIUS + 1);
    if (flags & RTT_FLAG_GRAD_IOR) span(RTT_F_IOR_IN, RTT_F_IOR_OUT + 1);
}

struct Ck { V3 p, d; };

// shared reverse step used by both adjoint drivers
void reverse_row(const HostTable& T, int r, int lam, const Ck& ck, V3& gp, V3& gd, float& gI,
                 V3 g_hl, float g_w, float* g_table, float* g_lut, int64_t i, int b, int flag_mask = ~0) {
    const RowDev& R = T.rows[r];
    const int flags = R.i[RTT_I_FLAGS] & flag_mask;
    const Ior io = row_ior(T, r, lam);
    RowGrad G;
    zero(G);
    V3 ngp, ngd; float mod;
    interact_adjoint(R, ck.p, ck.d, io.ni, io.no, io.mu_enter, io.mu_exit, gp, gd, g_hl, v3(0, 0, 0), 0.0f,
                     ngp, ngd, mod, G, flags, make_aux(T.rows.data(), R, io.ni, io.no, i, r, b).u);
    gp = ngp; gd = ngd; gI = gI * mod + g_w;
    if (g_table && flags) {
        int fl = flags;
        if (T.L > 0 && (flags & RTT_FLAG_GRAD_IOR)) {
            if (g_lut) {
                g_lut[((size_t)lam * T.S + r) * 2] += G.g[RTT_F_IOR_IN];
                g_lut[((size_t)lam * T.S + r) * 2 + 1] += G.g[RTT_F_IOR_OUT];
            }
            fl &= ~RTT_FLAG_GRAD_IOR;
        }
        add_row_grad(G, fl, g_table + (size_t)r * RTT_ROW_G);
    }
}

struct HostRay { V3 p, d; float I, wav; };
HostRay fetch_ray(const rtt_source_t* src, const float* pos, const float* dir, const float* inten, const float* wav,
                  bool want_wav, int64_t i) {
    HostRay r;
    if (src) {
        source_ray(*src, source_key(*src), (long long)i, r.p, r.d);
        r.I = src->intensity; r.wav = src->wavelength;
    } else {
        r.p = load3(pos, i); r.d = load3(dir, i);
        r.I = inten ? inten[i] : 1.0f;
        r.wav = want_wav ? wav[i] : 0.0f;
    }
    return r;
}

}  // namespace

extern "C" {

int rtt_trace_seq_fwd(const float* in_pos, const float* in_dir, const float* in_intensity,
                      const float* in_wavelength, const rtt_source_t* source,
                      float* out_pos, float* out_dir, float* out_intensity, uint64_t* hitmask,
                      const rtt_table_t* table, const rtt_sensor_t* sensors, int32_t n_sensors,
                      int64_t n, int32_t, void*) {
    const HostTable T = stage(table);
#ifdef RTT_HOST_TILE
    // mirror of k_trace_seq_fwd_tile (FAST variant): frame-resident walk of rtt_tile.cuh
    std::vector<Xf> xf(T.S + 1);
    for (int r = 0; r <= T.S; ++r)
        xf[r] = make_xf(r > 0 ? &T.rows[r - 1] : nullptr, r < T.S ? &T.rows[r] : nullptr);
    for (int r = 0; r < T.S; ++r) edge_run_at(T.rows.data(), T.S, xf.data(), r);
#endif
#ifdef RTT_HOST_PAIR
    // mirror of k_trace_seq_fwd_pair: rays (i, i+1) are the two lanes of the packed arithmetic of rtt_pair.cuh; a ray
    // the kernel would walk in the reference's order (non-finite / un-normalised direction) is handled by the scalar
    // loop below, with its lane switched off in the pair
    std::vector<char> done(n, 0);
    struct HostDep {
        const rtt_sensor_t* sensors; int n_sensors; int64_t i0; int lam[2];
        void operator()(int lane, int slot, V3 hl, float w) const {
            if (slot >= 0 && slot < n_sensors) deposit(sensors[slot], i0 + lane, hl, w, lam[lane]);
        }
    };
    for (int64_t i0 = 0; i0 < n; i0 += 2) {
        HostRay ray[2];
        bool act[2] = {false, false};
        int lam[2] = {0, 0};
        for (int j = 0; j < 2; ++j) {
            ray[j].p = v3(0, 0, 0); ray[j].d = v3(0, 0, 0); ray[j].I = 0.0f; ray[j].wav = 0.0f;
            if (i0 + j >= n) continue;
            ray[j] = fetch_ray(source, in_pos, in_dir, in_intensity, in_wavelength, T.L > 0, i0 + j);
            lam[j] = T.L > 0 ? lam_index(T, ray[j].wav) : 0;
            act[j] = finite_ray(ray[j].p, ray[j].d) && regular_dir(ray[j].d);
            if (!act[j]) { ray[j].p = v3(0, 0, 0); ray[j].d = v3(0, 0, 0); ray[j].I = 0.0f; }
        }
        if (!act[0] && !act[1]) continue;
        P3 P = pack3(ray[0].p, ray[1].p), D = pack3(ray[0].d, ray[1].d);
        F2 I = f2(ray[0].I, ray[1].I);
        const unsigned actb = (act[0] ? 1u : 0u) | (act[1] ? 2u : 0u);
        uint64_t mask[2] = {0, 0};
        HostDep dep{sensors, n_sensors, i0, {lam[0], lam[1]}};
        for (int r = 0; r < T.S; ++r) {
            pair_apply_xf(xf[r], P, D);
            if (xf[r].run > 0 && pair_edge_culled(xf[r], P, D, act[0], act[1])) { r += xf[r].run - 1; continue; }
            const RowDev& R = T.rows[r];
            const Ior ia = row_ior(T, r, lam[0]), ib = row_ior(T, r, lam[1]);
            const F2 me = f2(ia.mu_enter, ib.mu_enter), mx = f2(ia.mu_exit, ib.mu_exit);
            unsigned hit;
            switch (tile_opcode(R)) {
                case 1: hit = pair_conic_face<true, RTT_SHAPE_SPHERIC_FACE>(T.rows.data(), r, P, D, I, actb, me, mx, dep); break;
                case 4: hit = pair_conic_face<false, RTT_SHAPE_CYL_FACE>(T.rows.data(), r, P, D, I, actb, me, mx, dep); break;
                case 7: hit = pair_plane<RTT_BOUND_DISK, RTT_PHYS_APERTURE, false>(T.rows.data(), r, P, D, I, actb, dep); break;
                case 8: hit = pair_plane<RTT_BOUND_DISK, RTT_PHYS_TRANSMIT, true>(T.rows.data(), r, P, D, I, actb, dep); break;
                case 9: hit = pair_plane<RTT_BOUND_RECT, RTT_PHYS_TRANSMIT, true>(T.rows.data(), r, P, D, I, actb, dep); break;
                default:
                    hit = pair_row_scalar<KDyn>(T.rows.data(), r, P, D, I, actb, me, mx,
                                                make_aux(T.rows.data(), R, ia.ni, ia.no, i0, r, 0),
                                                make_aux(T.rows.data(), R, ib.ni, ib.no, i0 + 1, r, 0), dep);
                    break;
            }
            if (hit & 1u) mask[0] |= 1ull << r;
            if (hit & 2u) mask[1] |= 1ull << r;
        }
        pair_apply_xf(xf[T.S], P, D);
        const V3 po[2] = {lane_a(P), lane_b(P)}, dout[2] = {lane_a(D), lane_b(D)};
        const float Io[2] = {I.x, I.y};
        for (int j = 0; j < 2; ++j) {
            if (!act[j]) continue;
            const int64_t i = i0 + j;
            if (out_pos) { store3(out_pos, i, po[j]); store3(out_dir, i, dout[j]); out_intensity[i] = Io[j]; }
            if (hitmask) hitmask[i] = mask[j];
            done[i] = 1;
        }
    }
#endif
    for (int64_t i = 0; i < n; ++i) {
#ifdef RTT_HOST_PAIR
        if (done[i]) continue;
#endif
        const HostRay ray = fetch_ray(source, in_pos, in_dir, in_intensity, in_wavelength, T.L > 0, i);
        V3 p = ray.p, d = ray.d;
        float I = ray.I;
        const int lam = T.L > 0 ? lam_index(T, ray.wav) : 0;
        uint64_t mask = 0;
        if (!finite_ray(p, d)) {                                 // NaN / inf rays hit nothing, like the kernels
            if (out_pos) { store3(out_pos, i, p); store3(out_dir, i, d); out_intensity[i] = I; }
            if (hitmask) hitmask[i] = 0;
            continue;
        }
#ifdef RTT_HOST_TILE
        if (regular_dir(d)) {
            for (int r = 0; r < T.S; ++r) {
                apply_xf(xf[r], p, d);
                if (xf[r].run > 0 && edge_culled(xf[r], p, d)) { r += xf[r].run - 1; continue; }   // lens-edge culling
                float t;
                if (!tile_test(T.rows.data(), r, p, d, t)) continue;
                const RowDev& R = T.rows[r];
                const Ior io = row_ior(T, r, lam);
                V3 np, nd, hl; float mod;
                tile_interact(R, p, d, t, io.mu_enter, io.mu_exit, np, nd, mod, hl,
                              make_aux(T.rows.data(), R, io.ni, io.no, i, r, 0));
                const int slot = R.i[RTT_I_SENSOR];
                if (slot >= 0 && slot < n_sensors) deposit(sensors[slot], i, hl, I, lam);
                p = np; d = nd; I = I * mod;
                mask |= 1ull << r;
            }
            apply_xf(xf[T.S], p, d);
            if (out_pos) { store3(out_pos, i, p); store3(out_dir, i, d); out_intensity[i] = I; }
            if (hitmask) hitmask[i] = mask;
            continue;
        }
#endif
        for (int r = 0; r < T.S; ++r) {
            Frames F; Roots q; float t; int which;
            if (!intersect<true>(T.rows.data(), r, p, d, F, q, t, which)) continue;
            const RowDev& R = T.rows[r];
            const Ior io = row_ior(T, r, lam);
            const Step s = interact(R, F, t, p, d, io.mu_enter, io.mu_exit, make_aux(T.rows.data(), R, io.ni, io.no, i, r, 0));
            const int slot = R.i[RTT_I_SENSOR];
            if (slot >= 0 && slot < n_sensors) deposit(sensors[slot], i, s.hit_local, I, lam);
            p = s.hit_global; d = s.new_dir; I = I * s.mod;
            mask |= 1ull << r;
        }
        if (out_pos) { store3(out_pos, i, p); store3(out_dir, i, d); out_intensity[i] = I; }
        if (hitmask) hitmask[i] = mask;
    }
    return 0;
}

int rtt_trace_seq_bwd(const float* in_pos, const float* in_dir, const float*,
                      const float* in_wavelength, const rtt_source_t* source, const uint64_t* hitmask,
                      const float* g_out_pos, const float* g_out_dir, const float* g_out_intensity,
                      const float* const* g_record,
                      float* g_in_pos, float* g_in_dir, float* g_in_intensity,
                      float* g_table, float* g_lut,
                      const rtt_table_t* table, int32_t n_sensors,
                      int64_t n, int32_t mode, void*) {
    const HostTable T = stage(table);
    // RTT_MODE_SCALAR_GRADS: pose-gradient requests are ignored (include/rtt_b200.h)
    const int flag_mask = (mode & RTT_MODE_SCALAR_GRADS) ? ~(RTT_FLAG_GRAD_POSE_E | RTT_FLAG_GRAD_POSE_S) : ~0;
    std::vector<Ck> ck(RTT_MAX_ROWS);
#ifdef RTT_HOST_TILE
    // mirror of the LEAN path of k_trace_seq_bwd (FAST build, scalar gradients, no input-ray gradients; mode tune bit 8
    // switches it off): rays whose every interaction is a lean row take the frame-resident steps of rtt_lean.cuh
    const bool lean_on = (mode & RTT_MODE_SCALAR_GRADS) && !g_in_pos && !g_in_dir && !g_in_intensity &&
                         !(((mode & RTT_MODE_TUNE_MASK) >> RTT_MODE_TUNE_SHIFT) & 8);
    std::vector<Xf> xf(T.S + 1);
    std::vector<int> tile_op(T.S, 0);
    uint64_t lean_rows = 0;
    if (lean_on) {
        for (int r = 0; r <= T.S; ++r)
            xf[r] = make_xf(r > 0 ? &T.rows[r - 1] : nullptr, r < T.S ? &T.rows[r] : nullptr);
        int used = 0;                                                    // private gradient slots, as in the kernel
        for (int r = 0; r < T.S; ++r) {
            tile_op[r] = tile_opcode(T.rows[r]);
            const int fl_all = T.rows[r].i[RTT_I_FLAGS];
            const bool scalar_only = fl_all != 0 && !(fl_all & (RTT_FLAG_GRAD_POSE_E | RTT_FLAG_GRAD_POSE_S)) &&
                                     !(T.L > 0 && (fl_all & RTT_FLAG_GRAD_IOR));
            const bool slot = g_table && scalar_only && used < 12;
            if (slot) ++used;
            const int fl = g_table ? (fl_all & flag_mask) : 0;
            if (lean_tile_op(tile_op[r]) && (fl == 0 || slot)) lean_rows |= 1ull << r;
        }
    }
    std::vector<float> ckw(RTT_MAX_ROWS * 6);
#endif
    for (int64_t i = 0; i < n; ++i) {
        const HostRay ray = fetch_ray(source, in_pos, in_dir, nullptr, in_wavelength, T.L > 0, i);
        V3 p = ray.p, d = ray.d;
        const int lam = T.L > 0 ? lam_index(T, ray.wav) : 0;
        const uint64_t mask = hitmask[i];
        int nh = 0;
#ifdef RTT_HOST_TILE
        const int cap = (T.S <= 24 ? 24 : RTT_MAX_ROWS) * 6 / kLeanCkWords;
        if (lean_on && mask != 0 && (mask & ~lean_rows) == 0 && __builtin_popcountll(mask) <= cap && finite_ray(p, d) &&
            regular_dir(d)) {
            V3 lp = p, ld = d;
            for (int r = 0; r < T.S; ++r) {
                apply_xf(xf[r], lp, ld);
                if (!((mask >> r) & 1ull)) continue;
                const RowDev& R = T.rows[r];
                const Ior io = row_ior(T, r, lam);
                LeanCk c;
                if (tile_op[r] == 1) lean_face_replay<true>(R, io.mu_enter, io.mu_exit, lp, ld, c);
                else if (tile_op[r] == 4) lean_face_replay<false>(R, io.mu_enter, io.mu_exit, lp, ld, c);
                else lean_plane_replay(R, tile_op[r], lp, ld, c);
                lean_ck_store(ckw.data() + kLeanCkWords * nh, c);
                ++nh;
            }
            V3 gp = g_out_pos ? load3(g_out_pos, i) : v3(0, 0, 0);
            V3 gd = g_out_dir ? load3(g_out_dir, i) : v3(0, 0, 0);
            lean_xf_transpose(xf[T.S], gp, gd);
            for (int r = T.S - 1; r >= 0; --r) {
                if ((mask >> r) & 1ull) {
                    --nh;
                    const LeanCk c = lean_ck_load(ckw.data() + kLeanCkWords * nh);
                    const RowDev& R = T.rows[r];
                    if (tile_op[r] == 1 || tile_op[r] == 4) {
                        const int flags = g_table ? R.i[RTT_I_FLAGS] : 0;
                        const Ior io = row_ior(T, r, lam);
                        float g5[5] = {0, 0, 0, 0, 0};
                        if (tile_op[r] == 1) lean_face_reverse<true>(R, io.mu_enter, io.mu_exit, io.ni, io.no, c, gp, gd, flags, g5);
                        else lean_face_reverse<false>(R, io.mu_enter, io.mu_exit, io.ni, io.no, c, gp, gd, flags, g5);
                        if (g_table) {
                            float* row = g_table + (size_t)r * RTT_ROW_G;
                            if (flags & RTT_FLAG_GRAD_CK) { row[RTT_F_C] += g5[0]; row[RTT_F_K] += g5[1]; }
                            if (flags & RTT_FLAG_GRAD_IOR) { row[RTT_F_IOR_IN] += g5[3]; row[RTT_F_IOR_OUT] += g5[4]; }
                        }
                    } else {
                        V3 g_hl = v3(0, 0, 0);
                        const int slot = R.i[RTT_I_SENSOR];
                        if (tile_op[r] != 7 && slot >= 0 && slot < n_sensors && g_record && g_record[slot]) {
                            const float* gr = g_record[slot] + 4 * i;
                            g_hl = v3(gr[0], gr[1], gr[2]);
                        }
                        lean_plane_reverse(R, tile_op[r], c, g_hl, gp, gd);
                    }
                }
                lean_xf_transpose(xf[r], gp, gd);
            }
            continue;
        }
#endif
        for (int r = 0; r < T.S; ++r) {
            if (!((mask >> r) & 1ull)) continue;
            ck[nh].p = p; ck[nh].d = d; ++nh;
            const RowDev& R = T.rows[r];
            const Frames F = to_frames(R, p, d);
            const Roots q = solve_roots(R, F.o, F.dd);
            int which;
            const float t = select_root(R, q, F.o, F.dd, &which);
            const Ior io = row_ior(T, r, lam);
            const Step s = interact(R, F, t, p, d, io.mu_enter, io.mu_exit, make_aux(T.rows.data(), R, io.ni, io.no, i, r, 0));
            p = s.hit_global; d = s.new_dir;
        }
        V3 gp = g_out_pos ? load3(g_out_pos, i) : v3(0, 0, 0);
        V3 gd = g_out_dir ? load3(g_out_dir, i) : v3(0, 0, 0);
        float gI = g_out_intensity ? g_out_intensity[i] : 0.0f;
        for (int r = T.S - 1; r >= 0; --r) {
            if (!((mask >> r) & 1ull)) continue;
            --nh;
            V3 g_hl = v3(0, 0, 0); float g_w = 0.0f;
            const int slot = T.rows[r].i[RTT_I_SENSOR];
            if (slot >= 0 && slot < n_sensors && g_record && g_record[slot]) {
                const float* gr = g_record[slot] + 4 * i;
                g_hl = v3(gr[0], gr[1], gr[2]); g_w = gr[3];
            }
            reverse_row(T, r, lam, ck[nh], gp, gd, gI, g_hl, g_w, g_table, g_lut, i, 0, flag_mask);
        }
        if (g_in_pos) store3(g_in_pos, i, gp);
        if (g_in_dir) store3(g_in_dir, i, gd);
        if (g_in_intensity) g_in_intensity[i] = gI;
    }
    return 0;
}

int rtt_trace_nonseq_fwd(const float* in_pos, const float* in_dir, const float* in_intensity,
                         const float* in_wavelength, const rtt_source_t* source,
                         float* out_pos, float* out_dir, float* out_intensity,
                         uint8_t* hit_seq, uint8_t* n_hits,
                         const rtt_table_t* table, const rtt_sensor_t* sensors, int32_t n_sensors,
                         int32_t nbounces, int64_t n, int32_t, void*) {
    const HostTable T = stage(table);
    std::vector<NsCull> cull(T.S);
    for (int r = 0; r < T.S; ++r) cull[r] = box_cull_info(T.rows.data(), T.S, r);
    for (int64_t i = 0; i < n; ++i) {
        const HostRay ray = fetch_ray(source, in_pos, in_dir, in_intensity, in_wavelength, T.L > 0, i);
        V3 p = ray.p, d = ray.d;
        float I = ray.I;
        const int lam = T.L > 0 ? lam_index(T, ray.wav) : 0;
        int cnt[RTT_MAX_SENSORS] = {0, 0, 0, 0};
        int nb = 0;
        for (; nb < nbounces; ++nb) {
            if (!(I > 0.0f) || !finite_ray(p, d)) break;
            float best = rtt_inf();
            int win = -1;
            bool poisoned = false;
            for (int r = 0; r < T.S; ++r) {
                const NsCull& bc = cull[r];                                   // box culling, as in the kernel
                if (bc.run > 0 && sphere_missed(bc, p, d)) { r += bc.run - 1; continue; }
                Frames F; Roots q; float t; int which;
                const bool valid = intersect<true>(T.rows.data(), r, p, d, F, q, t, which);
                if (T.rows[r].i[RTT_I_SHAPE] == RTT_SHAPE_NONE && is_nan(t)) poisoned = true;
                if (valid && t < best) { best = t; win = r; }
            }
            if (poisoned || win < 0) break;
            Frames F; Roots q; float t; int which;
            intersect<false>(T.rows.data(), win, p, d, F, q, t, which);
            const RowDev& R = T.rows[win];
            const Ior io = row_ior(T, win, lam);
            const Step s = interact(R, F, t, p, d, io.mu_enter, io.mu_exit, make_aux(T.rows.data(), R, io.ni, io.no, i, win, nb));
            const int slot = R.i[RTT_I_SENSOR];
            if (slot >= 0 && slot < n_sensors) {
                deposit(sensors[slot], i, s.hit_local, I, lam, cnt[slot], n);
                if (cnt[slot] < 255) ++cnt[slot];
            }
            p = s.hit_global; d = s.new_dir; I = I * s.mod;
            if (hit_seq) hit_seq[i * nbounces + nb] = (uint8_t)win;
        }
        if (hit_seq) for (int b = nb; b < nbounces; ++b) hit_seq[i * nbounces + b] = 255;
        if (n_hits) n_hits[i] = (uint8_t)nb;
        for (int s = 0; s < n_sensors; ++s) if (sensors[s].count) sensors[s].count[i] = (uint8_t)cnt[s];
        if (out_pos) { store3(out_pos, i, p); store3(out_dir, i, d); out_intensity[i] = I; }
    }
    return 0;
}

int rtt_trace_nonseq_bwd(const float* in_pos, const float* in_dir, const float*,
                         const float* in_wavelength, const rtt_source_t* source,
                         const uint8_t* hit_seq, int32_t nbounces,
                         const float* g_out_pos, const float* g_out_dir, const float* g_out_intensity,
                         const float* const* g_record, const int32_t* record_hits,
                         float* g_in_pos, float* g_in_dir, float* g_in_intensity,
                         float* g_table, float* g_lut,
                         const rtt_table_t* table, int32_t n_sensors, int64_t n, int32_t, void*) {
    const HostTable T = stage(table);
    std::vector<Ck> ck(nbounces + 1);
    std::vector<int> rows_hit(nbounces + 1);
    for (int64_t i = 0; i < n; ++i) {
        const HostRay ray = fetch_ray(source, in_pos, in_dir, nullptr, in_wavelength, T.L > 0, i);
        V3 p = ray.p, d = ray.d;
        const int lam = T.L > 0 ? lam_index(T, ray.wav) : 0;
        int cnt[RTT_MAX_SENSORS] = {0, 0, 0, 0};
        int nh = 0;
        for (int b = 0; b < nbounces; ++b) {                     // every interaction (the CUDA kernel replays in windows)
            const int r = hit_seq[i * nbounces + b];
            if (r == 255) break;
            ck[nh].p = p; ck[nh].d = d; rows_hit[nh] = r; ++nh;
            const RowDev& R = T.rows[r];
            if (R.i[RTT_I_SENSOR] >= 0 && R.i[RTT_I_SENSOR] < n_sensors && cnt[R.i[RTT_I_SENSOR]] < 255) ++cnt[R.i[RTT_I_SENSOR]];
            const Frames F = to_frames(R, p, d);
            const Roots q = solve_roots(R, F.o, F.dd);
            int which;
            const float t = select_root(R, q, F.o, F.dd, &which);
            const Ior io = row_ior(T, r, lam);
            const Step s = interact(R, F, t, p, d, io.mu_enter, io.mu_exit, make_aux(T.rows.data(), R, io.ni, io.no, i, r, b));
            p = s.hit_global; d = s.new_dir;
        }
        V3 gp = g_out_pos ? load3(g_out_pos, i) : v3(0, 0, 0);
        V3 gd = g_out_dir ? load3(g_out_dir, i) : v3(0, 0, 0);
        float gI = g_out_intensity ? g_out_intensity[i] : 0.0f;
        while (nh > 0) {
            --nh;
            V3 g_hl = v3(0, 0, 0); float g_w = 0.0f;
            const int slot = T.rows[rows_hit[nh]].i[RTT_I_SENSOR];
            if (slot >= 0 && slot < n_sensors) {
                const int ord = --cnt[slot];
                const int K = (record_hits && record_hits[slot] > 1) ? record_hits[slot] : 1;
                if (g_record && g_record[slot] && ord < K) {
                    const float* gr = g_record[slot] + 4 * ((int64_t)ord * n + i);
                    g_hl = v3(gr[0], gr[1], gr[2]); g_w = gr[3];
                }
            }
            reverse_row(T, rows_hit[nh], lam, ck[nh], gp, gd, gI, g_hl, g_w, g_table, g_lut, i, nh);
        }
        if (g_in_pos) store3(g_in_pos, i, gp);
        if (g_in_dir) store3(g_in_dir, i, gd);
        if (g_in_intensity) g_in_intensity[i] = gI;
    }
    return 0;
}

int rtt_intersect_test(const float* in_pos, const float* in_dir, float* t_out,
                       const rtt_table_t* table, int32_t row0, int32_t k,
                       int64_t n, int32_t, void*) {
    rtt_table_t tb = *table; tb.n_lut = 0;
    const HostTable T = stage(&tb);
    for (int64_t i = 0; i < n; ++i) {
        const V3 p = load3(in_pos, i), d = load3(in_dir, i);
        for (int j = 0; j < k; ++j) {
            Frames F; Roots q; float t; int which;
            const bool valid = intersect<true>(T.rows.data(), row0 + j, p, d, F, q, t, which);
            const bool bare = T.rows[row0 + j].i[RTT_I_SHAPE] == RTT_SHAPE_NONE;
            t_out[i * k + j] = bare ? t : (valid ? t : rtt_inf());
        }
    }
    return 0;
}

int rtt_surface_step_fwd(const float* in_pos, const float* in_dir, const float* in_wavelength,
                         float* new_pos, float* new_dir, float* mod,
                         float* hit_local, float* t_out, float* normal,
                         const rtt_table_t* table, int32_t row, int64_t n, int32_t, void*) {
    const HostTable T = stage(table);
    const RowDev& R = T.rows[row];
    for (int64_t i = 0; i < n; ++i) {
        const V3 p = load3(in_pos, i), d = load3(in_dir, i);
        const int lam = T.L > 0 ? lam_index(T, in_wavelength[i]) : 0;
        const Frames F = to_frames(R, p, d);
        const Roots q = solve_roots(R, F.o, F.dd);
        int which;
        const float t = select_root(R, q, F.o, F.dd, &which);
        const Ior io = row_ior(T, row, lam);
        const Step s = interact(R, F, t, p, d, io.mu_enter, io.mu_exit, make_aux(T.rows.data(), R, io.ni, io.no, i, row, 0));
        store3(new_pos, i, s.hit_global); store3(new_dir, i, s.new_dir);
        mod[i] = s.mod;
        if (hit_local) store3(hit_local, i, s.hit_local);
        if (t_out) t_out[i] = t;
        if (normal) store3(normal, i, s.normal);
    }
    return 0;
}

int rtt_surface_step_bwd(const float* in_pos, const float* in_dir, const float* in_wavelength,
                         const float* g_new_pos, const float* g_new_dir, const float* g_hit_local,
                         const float* g_t, const float* g_normal,
                         float* g_in_pos, float* g_in_dir, float* g_table, float* g_lut,
                         const rtt_table_t* table, int32_t row, int64_t n, int32_t, void*) {
    const HostTable T = stage(table);
    const RowDev& R = T.rows[row];
    const int flags = R.i[RTT_I_FLAGS];
    for (int64_t i = 0; i < n; ++i) {
        const V3 p = load3(in_pos, i), d = load3(in_dir, i);
        const int lam = T.L > 0 ? lam_index(T, in_wavelength[i]) : 0;
        const Ior io = row_ior(T, row, lam);
        RowGrad G;
        zero(G);
        V3 gp, gd; float mod;
        interact_adjoint(R, p, d, io.ni, io.no, io.mu_enter, io.mu_exit,
                         g_new_pos ? load3(g_new_pos, i) : v3(0, 0, 0),
                         g_new_dir ? load3(g_new_dir, i) : v3(0, 0, 0),
                         g_hit_local ? load3(g_hit_local, i) : v3(0, 0, 0),
                         g_normal ? load3(g_normal, i) : v3(0, 0, 0), g_t ? g_t[i] : 0.0f,
                         gp, gd, mod, G, flags, make_aux(T.rows.data(), R, io.ni, io.no, i, row, 0).u);
        if (g_in_pos) store3(g_in_pos, i, gp);
        if (g_in_dir) store3(g_in_dir, i, gd);
        if (g_table && flags) {
            int fl = flags;
            if (T.L > 0 && (flags & RTT_FLAG_GRAD_IOR)) {
                if (g_lut) {
                    g_lut[((size_t)lam * T.S + row) * 2] += G.g[RTT_F_IOR_IN];
                    g_lut[((size_t)lam * T.S + row) * 2 + 1] += G.g[RTT_F_IOR_OUT];
                }
                fl &= ~RTT_FLAG_GRAD_IOR;
            }
            add_row_grad(G, fl, g_table + (size_t)row * RTT_ROW_G);
        }
    }
    return 0;
}

int rtt_sample_bundle(const rtt_source_t* source, float* pos, float* dir, float* intensity, float* wavelength,
                      int64_t n, int32_t, void*) {
    if (!source || !source->pose) return RTT_E_ARG;
    const SourceKey k = source_key(*source);
    for (int64_t i = 0; i < n; ++i) {
        V3 p, d;
        source_ray(*source, k, (long long)i, p, d);
        store3(pos, i, p); store3(dir, i, d);
        intensity[i] = source->intensity;
        if (wavelength) wavelength[i] = source->wavelength;
    }
    return 0;
}

// ---- goal reductions (rtt_goals.cu), restated with plain double-precision loops -----------------------
namespace {
struct HCentre { double W, cx, cy; bool clamped; };
HCentre hcentre(const float* mom4, const float* target_xy) {
    HCentre c;
    c.clamped = !(mom4[0] >= 1e-12f);
    c.W = c.clamped ? 1e-12 : (double)mom4[0];
    c.cx = target_xy ? (double)target_xy[0] : (double)mom4[1] / c.W;
    c.cy = target_xy ? (double)target_xy[1] : (double)mom4[2] / c.W;
    return c;
}
}  // namespace

int rtt_spot_moments(const float* rec, int64_t m, int32_t active_only, float* out4, float*, void*) {
    double a[4] = {0, 0, 0, 0};
    for (int64_t i = 0; i < m; ++i) {
        const float x = rec[4 * i], y = rec[4 * i + 1], w = rec[4 * i + 3];
        if (!active_only || w > 0.0f) { a[0] += w; a[1] += (double)x * w; a[2] += (double)y * w; }
        if (w > 0.0f) a[3] += 1.0;
    }
    for (int k = 0; k < 4; ++k) out4[k] = (float)a[k];
    return 0;
}

int rtt_spot_moments_bwd(const float* rec, int64_t m, int32_t active_only, const float* g3, float* g_rec, void*) {
    for (int64_t i = 0; i < m; ++i) {
        const float x = rec[4 * i], y = rec[4 * i + 1], w = rec[4 * i + 3];
        const bool on = !active_only || w > 0.0f;
        g_rec[4 * i] = on ? g3[1] * w : 0.0f; g_rec[4 * i + 1] = on ? g3[2] * w : 0.0f; g_rec[4 * i + 2] = 0.0f;
        g_rec[4 * i + 3] = on ? g3[0] + g3[1] * x + g3[2] * y : 0.0f;
    }
    return 0;
}

int rtt_spot_size_fwd(const float* rec, int64_t m, const float* mom4, const float* target_xy, float* out3,
                      float*, void*) {
    const HCentre c = hcentre(mom4, target_xy);
    double L = 0, gx = 0, gy = 0;
    for (int64_t i = 0; i < m; ++i) {
        const double x = rec[4 * i], y = rec[4 * i + 1], w = rec[4 * i + 3];
        if (!(w > 0.0)) continue;
        const double dx = x - c.cx, dy = y - c.cy, wn = w / c.W;
        const double q = (dx * dx + dy * dy) * wn, rms = std::sqrt(q);
        L += rms;
        const double a = q > 0.0 ? 0.5 / rms : 0.0;
        gx += a * (-2.0 * dx * wn); gy += a * (-2.0 * dy * wn);
    }
    out3[0] = (float)L; out3[1] = (float)gx; out3[2] = (float)gy;
    return 0;
}

int rtt_spot_size_bwd(const float* rec, int64_t m, const float* mom4, const float* target_xy, const float* out3,
                      const float* g_loss, float* g_rec, void*) {
    const HCentre c = hcentre(mom4, target_xy);
    const double gL = g_loss[0];
    const double Gcx = target_xy ? 0.0 : out3[1], Gcy = target_xy ? 0.0 : out3[2];
    const double gW = c.clamped ? 0.0 : (-0.5 * out3[0] / c.W - (Gcx * c.cx + Gcy * c.cy) / c.W);
    for (int64_t i = 0; i < m; ++i) {
        const double x = rec[4 * i], y = rec[4 * i + 1], w = rec[4 * i + 3];
        double g0 = 0, g1 = 0, g3 = 0;
        if (w > 0.0) {
            const double dx = x - c.cx, dy = y - c.cy, wn = w / c.W, r2 = dx * dx + dy * dy, q = r2 * wn;
            const double a = q > 0.0 ? 0.5 / std::sqrt(q) : 0.0;
            g0 = gL * (a * 2.0 * dx * wn + Gcx * wn);
            g1 = gL * (a * 2.0 * dy * wn + Gcy * wn);
            g3 = gL * (a * r2 / c.W + (Gcx * x + Gcy * y) / c.W + gW);
        }
        g_rec[4 * i] = (float)g0; g_rec[4 * i + 1] = (float)g1; g_rec[4 * i + 2] = 0.0f; g_rec[4 * i + 3] = (float)g3;
    }
    return 0;
}


// ---- per-id sensor moments (host mirror of csrc/rtt_goals.cu k_spot_id_*; same formulas, double accumulation) ----
int rtt_spot_id_moments(const float* rec, const int8_t* ids, int64_t m, const int32_t* group_of, int32_t K,
                        float* out, float*, void*) {
    std::vector<double> acc(4 * (size_t)K, 0.0);
    for (int64_t i = 0; i < m; ++i) {
        const int g = group_of[(int)ids[i] + 128];
        const float* r = rec + 4 * i;
        if (g < 0 || r[3] == 0.0f) continue;
        acc[4 * g] += r[3]; acc[4 * g + 1] += (double)(r[3] * r[0]); acc[4 * g + 2] += (double)(r[3] * r[1]);
        acc[4 * g + 3] += r[3] > 0.0f ? 1.0 : 0.0;
    }
    for (size_t k = 0; k < acc.size(); ++k) out[k] = (float)acc[k];
    return 0;
}

static inline void spot_id_terms(float dx, float dy, float p, float& pw, float& sx, float& sy) {
    if (p == 2.0f) { sx = 2.0f * dx; sy = 2.0f * dy; pw = dx * dx + dy * dy; return; }
    const float ax = std::fabs(dx), ay = std::fabs(dy);
    const float px = std::pow(ax, p - 1.0f), py = std::pow(ay, p - 1.0f);
    sx = p * px * (dx > 0.0f ? 1.0f : (dx < 0.0f ? -1.0f : 0.0f));
    sy = p * py * (dy > 0.0f ? 1.0f : (dy < 0.0f ? -1.0f : 0.0f));
    pw = px * ax + py * ay;
}

int rtt_spot_id_size(const float* rec, const int8_t* ids, int64_t m, const int32_t* group_of, int32_t K,
                     const float* centres, float p, float* out, float*, void*) {
    std::vector<double> acc(4 * (size_t)K, 0.0);
    for (int64_t i = 0; i < m; ++i) {
        const int g = group_of[(int)ids[i] + 128];
        const float* r = rec + 4 * i;
        if (g < 0 || r[3] == 0.0f) continue;
        float pw, sx, sy;
        spot_id_terms(r[0] - centres[2 * g], r[1] - centres[2 * g + 1], p, pw, sx, sy);
        acc[4 * g] += (double)(r[3] * pw); acc[4 * g + 1] += (double)(r[3] * sx); acc[4 * g + 2] += (double)(r[3] * sy);
    }
    for (size_t k = 0; k < acc.size(); ++k) out[k] = (float)acc[k];
    return 0;
}

int rtt_spot_id_size_bwd(const float* rec, const int8_t* ids, int64_t m, const int32_t* group_of, int32_t,
                         const float* coef, float p, float* g_rec, void*) {
    for (int64_t i = 0; i < m; ++i) {
        const int g = group_of[(int)ids[i] + 128];
        float* o = g_rec + 4 * i;
        o[0] = o[1] = o[2] = o[3] = 0.0f;
        if (g < 0) continue;
        const float* r = rec + 4 * i;
        const float* c = coef + 8 * g;
        const float dx = r[0] - c[0], dy = r[1] - c[1];
        float pw, sx, sy;
        spot_id_terms(dx, dy, p, pw, sx, sy);
        o[0] = c[2] * r[3] * (sx - c[3]);
        o[1] = c[2] * r[3] * (sy - c[4]);
        o[3] = c[2] * (pw - c[3] * dx - c[4] * dy - c[5]);
    }
    return 0;
}

// Renderer.render_3d: host mirror of k_render_shade (nearest hit, winner's normal, Lambert shading)
int rtt_render_shade(const float* in_pos, const float* in_dir, const rtt_source_t* source, const rtt_table_t* table,
                     const float* base_rgb, const float* light_dir, const float* background,
                     float* out_rgb, uint8_t* out_row, int64_t n, int32_t, void*) {
    rtt_table_t tb = *table; tb.n_lut = 0;
    const HostTable T = stage(&tb);
    const V3 light = v3(light_dir[0], light_dir[1], light_dir[2]);
    auto clamp01 = [](float v) { return std::fmin(std::fmax(v, 0.0f), 1.0f); };
    for (int64_t i = 0; i < n; ++i) {
        const HostRay ray = fetch_ray(source, in_pos, in_dir, nullptr, nullptr, false, i);
        float best = rtt_inf();
        int win = -1;
        bool poisoned = !finite_ray(ray.p, ray.d);
        for (int r = 0; r < T.S && !poisoned; ++r) {
            Frames F; Roots q; float t; int which;
            const bool valid = intersect<true>(T.rows.data(), r, ray.p, ray.d, F, q, t, which);
            if (T.rows[r].i[RTT_I_SHAPE] == RTT_SHAPE_NONE && is_nan(t)) poisoned = true;
            if (valid && t < best) { best = t; win = r; }
        }
        V3 rgb = v3(background[0], background[1], background[2]);
        if (!poisoned && win >= 0) {
            Frames F; Roots q; float t; int which;
            intersect<false>(T.rows.data(), win, ray.p, ray.d, F, q, t, which);
            const RowDev& R = T.rows[win];
            float nlen;
            const V3 nn = normal_global(R, normal_local(R, along(F.o, t, F.dd), &nlen));
            const float shade = 0.3f + 0.7f * std::fabs(dot(nn, light));
            rgb = v3(base_rgb[3 * win] * shade, base_rgb[3 * win + 1] * shade, base_rgb[3 * win + 2] * shade);
        } else {
            win = 255;
        }
        out_rgb[3 * i] = clamp01(rgb.x); out_rgb[3 * i + 1] = clamp01(rgb.y); out_rgb[3 * i + 2] = clamp01(rgb.z);
        if (out_row) out_row[i] = (uint8_t)win;
    }
    return 0;
}

}  // extern "C"
