// rtt_lean.cuh — frame-resident replay and reverse steps of the FAST sequential adjoint ("lean" path).
//
// The general adjoint (rtt_core.cuh interact_adjoint) differentiates every row kind in the reference's operation order:
// element pose, renormalisation, surface pose, root selection, normal, physics, each with its pose-gradient outer
// products.  The optimisation loops of the reference (tests/test_optimize_singlet.py:66-116, optim/goals.py:144-187)
// differentiate lens prescriptions — curvatures, conic constants, indices — through lens faces, stops and sensors, and
// for those row kinds most of that work is the identity:
//   * the ray state lives in the ELEMENT frame of the row (rtt_tile.cuh): poses are frame changes between elements
//     (Xf), applied once per element on the way down and transposed once on the way back;
//   * a unit direction stays unit through Snell / reflection / transmission, so the per-row F.normalize of
//     geom/shape.py:38 is the identity and its Jacobian (I - d d^T) acts as the identity on every perturbation that
//     reaches it (the perturbations of a unit vector are tangent to the sphere);
//   * the distance t is the root of F(o + t d) = 0, so dt = -(grad F . (do + t dd) + F_c dc + F_k dk) / (grad F . d):
//     the reverse step needs the surface gradient at the hit point — which the normal needs anyway — and never solves
//     the quadratic again.  The replay keeps (hit point, incoming direction, t) per interaction: 7 words.
// Same selections as the forward pass (lower valid root, A ~ 0 fallback, entering / exiting, total internal reflection):
// masks are non-differentiable, as in the reference's autograd.  Formulas: geom/primitives.py:280-343, 356-395 (conic),
// :124-136 (plane), phys/std.py:123-145 (Snell), phys/filter.py:24-33 (stop), elements/sensor.py:22-39 (records).
//
// A ray takes this path when every row it interacted with is one of the kinds below, its input direction is regular
// (rtt_tile.cuh) and only scalar parameter gradients are requested; any other ray runs the general adjoint in the same
// launch.  RTT_HD: tests/hostsim drives the same functions on the CPU against oracle autograd.
#pragma once
#include "rtt_tile.cuh"

namespace rtt {

// tile opcodes (RTT_TILE_SPECS) with a lean adjoint
RTT_HD bool lean_tile_op(int op) { return op == 1 || op == 4 || op == 7 || op == 8 || op == 9; }

struct LeanCk { V3 h, d; float t; };       // surface-frame hit point, incoming direction (row frame), distance
constexpr int kLeanCkWords = 7;

// 1 / sqrt(x): MUFU.RSQ in the FAST device build (<= 2 ulp), IEEE on the host checker
RTT_HD float lean_rsqrt(float x) {
#if defined(RTT_APPROX) && defined(__CUDA_ARCH__)
    return rsqrt_(x);
#else
    return 1.0f / sqrtf(x);
#endif
}
RTT_HD void lean_ck_store(float* w, const LeanCk& c) {
    w[0] = c.h.x; w[1] = c.h.y; w[2] = c.h.z; w[3] = c.d.x; w[4] = c.d.y; w[5] = c.d.z; w[6] = c.t;
}
RTT_HD LeanCk lean_ck_load(const float* w) {
    LeanCk c; c.h = v3(w[0], w[1], w[2]); c.d = v3(w[3], w[4], w[5]); c.t = w[6]; return c;
}

// transpose of a frame change, for gradients: p' = p @ M + c  =>  g_p = g_p' @ M^T
RTT_HD void lean_xf_transpose(const Xf& x, V3& gp, V3& gd) {
    if (x.kind != 2) return;
    gp = mul_RT(gp, x.M);
    gd = mul_RT(gd, x.M);
}

// gradient of the implicit form c (x^2 + y^2) + c (1 + k) z^2 - 2 z at h (QUADRIC_ZY: x dropped)
template <bool FULL>
RTT_HD V3 lean_conic_grad(float tc, float tc1k, V3 h) {
    return v3(FULL ? tc * h.x : 0.0f, tc * h.y, fmaf(tc1k, h.z, -2.0f));
}

// ---- lens face (tile ops 1, 4): conic, HALF bound (not inverted), Rs == I, Snell -------------------------------
template <bool FULL>
RTT_HD void lean_face_replay(const RowDev& R, float mu_enter, float mu_exit, V3& p, V3& d, LeanCk& ck) {
    const float c = R.f[RTT_F_C], c1k = R.f[D_C1K];
    const float tc = 2.0f * c, tc1k = 2.0f * c1k;
    const V3 o = p - ld3(R.f + RTT_F_TS);
    float A, B, Cq;
    if (FULL) {
        A = c * (d.x * d.x + d.y * d.y) + c1k * (d.z * d.z);
        B = (tc * (o.x * d.x + o.y * d.y) + (tc1k * o.z) * d.z) - 2.0f * d.z;
        Cq = (c * (o.x * o.x + o.y * o.y) + c1k * (o.z * o.z)) - 2.0f * o.z;
    } else {
        A = c * (d.y * d.y) + c1k * (d.z * d.z);
        B = (tc * (o.y * d.y) + (tc1k * o.z) * d.z) - 2.0f * d.z;
        Cq = (c * (o.y * o.y) + c1k * (o.z * o.z)) - 2.0f * o.z;
    }
    float t;
    if (fabsf(A) < 1e-6f) {                                             // geom/primitives.py:305-313
        const float Bs = (fabsf(B) < 1e-6f) ? 1e-6f : B;
        t = div_(-Cq, Bs);
    } else {                                                            // the lower root that passes its own tests, else the
        const float disc = B * B - (4.0f * A) * Cq;                     // upper (the hit mask says one of them did)
        const float sq = sqrt_(fabsf(disc));
        const float inv = rcp_(2.0f * A);
        const float r1 = (-B - sq) * inv, r2 = (-B + sq) * inv;
        const float lo = fminf(r1, r2), hi = fmaxf(r1, r2);
        const bool oklo = (lo > 1e-6f) & (fabsf(fmaf(lo, d.z, o.z) * c) < 1.000001f);
        t = oklo ? lo : hi;
    }
    const V3 h = along(o, t, d);
    const V3 g = lean_conic_grad<FULL>(tc, tc1k, h);
    const float inv = lean_rsqrt(fmaxf(fma3(g.x, g.x, g.y, g.y, g.z, g.z), 1e-30f));
    const V3 n = v3(-g.x * inv, -g.y * inv, -g.z * inv);
    const float dt = FULL ? fma3(d.x, n.x, d.y, n.y, d.z, n.z) : fmaf(d.z, n.z, d.y * n.y);
    const bool entering = dt < 0.0f;
    const float c1 = fabsf(dt);
    const float mu = entering ? mu_enter : mu_exit;
    const float term = 1.0f - (mu * mu) * (1.0f - c1 * c1);
    ck.h = h; ck.d = d; ck.t = t;
    p = along(p, t, d);
    if (term < 0.0f) {                                                  // total internal reflection
        const float tw = 2.0f * dt;
        d = v3(d.x - tw * n.x, d.y - tw * n.y, d.z - tw * n.z);
    } else {
        const float qf = mu * c1 - sqrt_(term);
        const float qs = entering ? qf : -qf;
        d = v3(fmaf(qs, n.x, mu * d.x), fmaf(qs, n.y, mu * d.y), fmaf(qs, n.z, mu * d.z));
    }
}

// (gp, gd): d/d (new position, new direction) in, d/d (incoming position, direction) out.  g5 = this row's private
// (c, k, radius, ior_in, ior_out) gradient slots; (ni, no) the indices behind (mu_enter, mu_exit).
template <bool FULL>
RTT_HD void lean_face_reverse(const RowDev& R, float mu_enter, float mu_exit, float ni, float no, const LeanCk& ck,
                              V3& gp, V3& gd, int want, float* g5) {
    const float c = R.f[RTT_F_C], k = R.f[RTT_F_K], c1k = R.f[D_C1K];
    const float tc = 2.0f * c, tc1k = 2.0f * c1k;
    const V3 h = ck.h, d = ck.d;
    const V3 g = lean_conic_grad<FULL>(tc, tc1k, h);
    const float inv = lean_rsqrt(fmaxf(fma3(g.x, g.x, g.y, g.y, g.z, g.z), 1e-30f));
    const V3 n = v3(-g.x * inv, -g.y * inv, -g.z * inv);
    const float dt = FULL ? fma3(d.x, n.x, d.y, n.y, d.z, n.z) : fmaf(d.z, n.z, d.y * n.y);
    const bool entering = dt < 0.0f;
    const float c1 = fabsf(dt);
    const float mu = entering ? mu_enter : mu_exit;
    const float one_m = 1.0f - c1 * c1;
    const float term = 1.0f - (mu * mu) * one_m;
    const V3 gdir = gd;
    const float gn = dot(gdir, n);
    V3 g_d, g_n;
    if (term < 0.0f) {                                                  // new_dir = d - 2 (d . n) n
        g_d = v3(gdir.x - 2.0f * gn * n.x, gdir.y - 2.0f * gn * n.y, gdir.z - 2.0f * gn * n.z);
        g_n = v3(-2.0f * (gn * d.x + dt * gdir.x), -2.0f * (gn * d.y + dt * gdir.y), -2.0f * (gn * d.z + dt * gdir.z));
    } else {                                                            // new_dir = mu d + sgn (mu |d . n| - c2) n
        const float c2 = sqrt_(term);
        const float sgn = entering ? 1.0f : -1.0f;
        const float qs = (mu * c1 - c2) * sgn;
        const float g_q = sgn * gn;
        const float g_term = (term > 0.0f) ? -0.5f * g_q * rcp_(c2) : 0.0f;
        const float g_c1 = mu * g_q + g_term * (2.0f * mu * mu * c1);
        const float g_dt = entering ? -g_c1 : ((dt > 0.0f) ? g_c1 : 0.0f);
        g_d = v3(fmaf(g_dt, n.x, mu * gdir.x), fmaf(g_dt, n.y, mu * gdir.y), fmaf(g_dt, n.z, mu * gdir.z));
        g_n = v3(fmaf(g_dt, d.x, qs * gdir.x), fmaf(g_dt, d.y, qs * gdir.y), fmaf(g_dt, d.z, qs * gdir.z));
        if (want & RTT_FLAG_GRAD_IOR) {
            const float g_mu = dot(gdir, d) + c1 * g_q + g_term * (-2.0f * mu * one_m);
            if (entering) {                                             // mu = no / ni
                const float s = g_mu * rcp_(ni);
                g5[4] += s; g5[3] -= s * mu;
            } else {                                                    // mu = ni / no
                const float s = g_mu * rcp_(no);
                g5[3] += s; g5[4] -= s * mu;
            }
        }
    }
    // n = -g / |g|:  d/d g = -(g_n - n (n . g_n)) / |g|
    const float pr = dot(n, g_n);
    const V3 g_g = v3((pr * n.x - g_n.x) * inv, (pr * n.y - g_n.y) * inv, (pr * n.z - g_n.z) * inv);
    // hit point: new position = h + Ts, and g = g(h)
    const V3 g_h = v3(FULL ? fmaf(tc, g_g.x, gp.x) : gp.x, fmaf(tc, g_g.y, gp.y), fmaf(tc1k, g_g.z, gp.z));
    // t = root of F(o + t d) = 0:  dF/dt = g . d,  dF/do = g,  dF/dd = t g,  dF/dc = x^2 + y^2 + (1 + k) z^2,  dF/dk = c z^2
    const float gdd = dot(g, d);                                        // = 2 A t + B
    const float dd2 = FULL ? fmaf(d.x, d.x, d.y * d.y) : d.y * d.y, dzz = d.z * d.z;
    const float A = fmaf(c, dd2, c1k * dzz);
    const bool lin = fabsf(A) < 1e-6f;
    float den = gdd;
    if (lin) {                                                          // the forward pass took t = -C / B (geom/primitives.py:305-313)
        const float B = gdd - 2.0f * A * ck.t;
        den = (fabsf(B) < 1e-6f) ? 1e-6f : B;
    }
    const float gC = -dot(g_h, d) * rcp_(den);
    const V3 g_o = v3(fmaf(gC, g.x, g_h.x), fmaf(gC, g.y, g_h.y), fmaf(gC, g.z, g_h.z));
    gp = g_o;
    gd = v3(fmaf(ck.t, g_o.x, g_d.x), fmaf(ck.t, g_o.y, g_d.y), fmaf(ck.t, g_o.z, g_d.z));
    float rr = FULL ? fmaf(h.x, h.x, h.y * h.y) : h.y * h.y, zz = h.z * h.z;
    if (lin) {
        // ... in which A does not appear: the reference's autograd sends nothing through A (a flat face, c = 0, has
        // dA/dc = dx^2 + dy^2 + (1 + k) dz^2 ~ 1), so the t^2 dA terms of the implicit form are taken out again
        const float gA = gC * ck.t * ck.t;
        if (FULL) gd.x -= gA * tc * d.x;
        gd.y -= gA * tc * d.y;
        gd.z -= gA * tc1k * d.z;
        rr -= ck.t * ck.t * dd2; zz -= ck.t * ck.t * dzz;
    }
    if (want & RTT_FLAG_GRAD_CK) {
        const float g_tc = FULL ? fmaf(g_g.x, h.x, g_g.y * h.y) : g_g.y * h.y, g_tc1k = g_g.z * h.z;
        g5[0] += fmaf(gC, fmaf(1.0f + k, zz, rr), 2.0f * fmaf(1.0f + k, g_tc1k, g_tc));
        g5[1] += c * fmaf(gC, zz, 2.0f * g_tc1k);
    }
}

// ---- bounded plane with Rs == I (tile ops 7: circular stop, 8 / 9: disk / rectangle sensor) ------------------
RTT_HD void lean_plane_replay(const RowDev& R, int op, V3& p, V3& d, LeanCk& ck) {
    const V3 o = p - ld3(R.f + RTT_F_TS);
    const float safe = (fabsf(d.z) < 1e-6f) ? 1e-8f : d.z;
    const float t = -o.z * rcp_(safe);
    ck.h = along(o, t, d); ck.d = d; ck.t = t;
    p = along(p, t, d);
    if (op == 7) {                                                      // phys/filter.py:31-33: mask WITHOUT the invert flag
        const float m = R.i[RTT_I_INVERT] ? 0.0f : 1.0f;
        d = v3(d.x * m, d.y * m, d.z * m);
    }
}

// g_hl: d/d hit_local of the sensor record (zero for a stop)
RTT_HD void lean_plane_reverse(const RowDev& R, int op, const LeanCk& ck, V3 g_hl, V3& gp, V3& gd) {
    const V3 d = ck.d;
    const float m = (op == 7 && R.i[RTT_I_INVERT]) ? 0.0f : 1.0f;
    const V3 g_h = gp + g_hl;
    const float safe = (fabsf(d.z) < 1e-6f) ? 1e-8f : d.z;
    const float g_ts = dot(g_h, d) * rcp_(safe);                        // t = -o.z / d.z
    V3 g_d = v3(fmaf(ck.t, g_h.x, m * gd.x), fmaf(ck.t, g_h.y, m * gd.y), fmaf(ck.t, g_h.z, m * gd.z));
    if (!(fabsf(d.z) < 1e-6f)) g_d.z -= g_ts * ck.t;
    gp = v3(g_h.x, g_h.y, g_h.z - g_ts);
    gd = g_d;
}

}  // namespace rtt
