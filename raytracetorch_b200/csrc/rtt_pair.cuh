// rtt_pair.cuh — packed two-ray arithmetic of the FAST sequential forward kernel (sm_100a FFMA2 / FMUL2 / FADD2).
//
// Blackwell's FP32 pipe executes `fma.rn.f32x2` / `mul.f32x2` / `add.f32x2`: one instruction, one issue slot, two
// independent IEEE fp32 results in an aligned register pair.  The forward trace is bound by instruction issue
// (profiles/r1_c2_seq_v3.md: 69 % issue slots, 26 % FMA pipe), and every thread of the tile kernel already carries two
// rays through the same row at the same time — so here the two rays of a thread ARE the two lanes of the packed
// operands: ray A in .x, ray B in .y.  Row constants enter as broadcast scalars (the SASS operand form `R.F32`, no
// packing instruction), compares / selects / MUFU stay scalar per lane and read the halves of the pair in place.
//
// Same algorithm as rtt_tile.cuh (frame-resident walk of scene/sequential.py:12-36; the reference lines are cited at
// the scalar twins in rtt_core.cuh / rtt_tile.cuh): the packed forms below restate
//   conic lens face  = conic_half_hit + shape_in_bounds + normal_local + physics<SNELL>   (rtt_tile.cuh:309, rtt_core.cuh)
//   bounded plane    = solve_roots<PLANE> + surface_in_bounds + physics<APERTURE|TRANSMIT>
//   frame change     = apply_xf, lens-edge cull = edge_culled
// with explicit FMA contraction.  Anything rare (total internal reflection, a degenerate normal, one lane of a pair on
// the A ~ 0 fallback and the other not) and every row kind without a packed form runs the SCALAR twin per lane, so the
// packed code never has to reproduce a corner case by itself.  RTT_HD: tests/hostsim compiles this file for the CPU
// (the lanes become two fmaf calls — bit-identical to the device's packed IEEE results) and checks it against the oracle.
#pragma once
#include "rtt_tile.cuh"

namespace rtt {

#if defined(__CUDACC__)
typedef float2 F2;
#else
struct F2 { float x, y; };
#endif

// Tile row kinds (RTT_TILE_SPECS) WITHOUT a packed form: they run the scalar twin per lane with their specialised policy.
#define RTT_PAIR_SCALAR_SPECS(X)                                                                              \
    X(2, RTT_SURF_CYLINDER, RTT_BOUND_NONE, RTT_SHAPE_SPHERIC_EDGE, RTT_PHYS_BLOCK, 1, 0)    /* inked edge     */ \
    X(3, RTT_SURF_CYLINDER, RTT_BOUND_NONE, RTT_SHAPE_SPHERIC_EDGE, RTT_PHYS_SNELL, 1, 0)    /* clear edge     */ \
    X(5, RTT_SURF_PLANE, RTT_BOUND_NONE, RTT_SHAPE_CYL_EDGE, RTT_PHYS_SNELL, 0, 0)           /* cyl. lens side */ \
    X(6, RTT_SURF_PLANE, RTT_BOUND_NONE, RTT_SHAPE_CYL_EDGE, RTT_PHYS_BLOCK, 0, 0)                                \
    X(10, RTT_SURF_QUADRIC, RTT_BOUND_HALF_DISK, RTT_SHAPE_NONE, RTT_PHYS_REFLECT, 1, 0)     /* sph. mirror    */ \
    X(11, RTT_SURF_PLANE, RTT_BOUND_NONE, RTT_SHAPE_POLY, RTT_PHYS_BLOCK, 0, 0)              /* box face       */

RTT_HD F2 f2(float a, float b) { F2 r; r.x = a; r.y = b; return r; }
RTT_HD F2 bc(float s) { return f2(s, s); }

#if defined(__CUDA_ARCH__)
RTT_HD F2 fma2(F2 a, F2 b, F2 c) { return __ffma2_rn(a, b, c); }
RTT_HD F2 mul2(F2 a, F2 b) { return __fmul2_rn(a, b); }
RTT_HD F2 add2(F2 a, F2 b) { return __fadd2_rn(a, b); }
#else
RTT_HD F2 fma2(F2 a, F2 b, F2 c) { return f2(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y)); }
RTT_HD F2 mul2(F2 a, F2 b) { return f2(a.x * b.x, a.y * b.y); }
RTT_HD F2 add2(F2 a, F2 b) { return f2(a.x + b.x, a.y + b.y); }
#endif
// a - b: b * (-1) is exact, so the fused form rounds a - b once, like a subtraction
RTT_HD F2 sub2(F2 a, F2 b) { return fma2(b, bc(-1.0f), a); }
RTT_HD F2 fma2s(F2 a, float s, F2 c) { return fma2(a, bc(s), c); }
RTT_HD F2 mul2s(F2 a, float s) { return mul2(a, bc(s)); }
RTT_HD F2 add2s(F2 a, float s) { return add2(a, bc(s)); }

// 1 / sqrt(x): MUFU.RSQ in the FAST device build (<= 2 ulp), IEEE on the host checker
RTT_HD float rsqrt_pk(float x) {
#if defined(RTT_APPROX) && defined(__CUDA_ARCH__)
    return rsqrt_(x);
#else
    return 1.0f / sqrtf(x);
#endif
}

struct P3 { F2 x, y, z; };
RTT_HD V3 lane_a(const P3& p) { return v3(p.x.x, p.y.x, p.z.x); }
RTT_HD V3 lane_b(const P3& p) { return v3(p.x.y, p.y.y, p.z.y); }
RTT_HD P3 pack3(V3 a, V3 b) { P3 p; p.x = f2(a.x, b.x); p.y = f2(a.y, b.y); p.z = f2(a.z, b.z); return p; }
// p + t * d
RTT_HD P3 along2(const P3& p, F2 t, const P3& d) {
    P3 r; r.x = fma2(t, d.x, p.x); r.y = fma2(t, d.y, p.y); r.z = fma2(t, d.z, p.z); return r;
}
// a . b as the chain fma(a.z, b.z, fma(a.y, b.y, a.x * b.x))
RTT_HD F2 dot2(const P3& a, const P3& b) { return fma2(a.z, b.z, fma2(a.y, b.y, mul2(a.x, b.x))); }

// ---- frame change (apply_xf) ------------------------------------------------------------------------
RTT_HD void pair_apply_xf(const Xf& x, P3& p, P3& d) {
    if (x.kind == 0) return;
    if (x.kind == 1) { p.x = add2s(p.x, x.c[0]); p.y = add2s(p.y, x.c[1]); p.z = add2s(p.z, x.c[2]); return; }
    const float* M = x.M;
    P3 q, e;                                             // a @ M: fma(a.z, M[6+j], fma(a.y, M[3+j], a.x * M[j]))  (mul_R)
    q.x = add2s(fma2s(p.z, M[6], fma2s(p.y, M[3], mul2s(p.x, M[0]))), x.c[0]);
    q.y = add2s(fma2s(p.z, M[7], fma2s(p.y, M[4], mul2s(p.x, M[1]))), x.c[1]);
    q.z = add2s(fma2s(p.z, M[8], fma2s(p.y, M[5], mul2s(p.x, M[2]))), x.c[2]);
    e.x = fma2s(d.z, M[6], fma2s(d.y, M[3], mul2s(d.x, M[0])));
    e.y = fma2s(d.z, M[7], fma2s(d.y, M[4], mul2s(d.x, M[1])));
    e.z = fma2s(d.z, M[8], fma2s(d.y, M[5], mul2s(d.x, M[2])));
    p = q; d = e;
}

// ---- lens-edge cull (edge_culled), both lanes; a lane that holds no live ray counts as culled ---------
RTT_HD bool pair_edge_culled(const Xf& x, const P3& p, const P3& d, bool act_a, bool act_b) {
    const bool za = (p.z.x > x.zhi && d.z.x >= 0.0f) || (p.z.x < x.zlo && d.z.x <= 0.0f);
    const bool zb = (p.z.y > x.zhi && d.z.y >= 0.0f) || (p.z.y < x.zlo && d.z.y <= 0.0f);
    const bool da = za || !act_a, db = zb || !act_b;
    if (da && db) return true;
    if (x.ctype < 2) return false;
    if ((!da && d.z.x == 0.0f) || (!db && d.z.y == 0.0f)) return false;
    // time at which each ray leaves the z enclosure, stretched a little (a longer segment is conservative)
    const F2 zt = f2(d.z.x > 0.0f ? x.zhi : x.zlo, d.z.y > 0.0f ? x.zhi : x.zlo);
    const F2 rz = f2(rcp_(d.z.x), rcp_(d.z.y));
    const F2 te = fma2(mul2(sub2(zt, p.z), rz), bc(1.0001f), bc(1e-4f));
    bool ia, ib;
    if (x.ctype == 2) {
        // both ends of the segment strictly inside the rectangle: max(|p - c|, |q - c|) < half extent, per axis (the
        // rectangle is already shrunk by 2e-3, far above the rounding of its centre / half-extent form)
        const F2 ux = add2s(p.x, -x.q[0]), uy = add2s(p.y, -x.q[2]);
        const F2 vx = fma2(te, d.x, ux), vy = fma2(te, d.y, uy);
        ia = fmaxf(fabsf(ux.x), fabsf(vx.x)) < x.q[1] && fmaxf(fabsf(uy.x), fabsf(vy.x)) < x.q[3];
        ib = fmaxf(fabsf(ux.y), fabsf(vx.y)) < x.q[1] && fmaxf(fabsf(uy.y), fabsf(vy.y)) < x.q[3];
    } else {
        const F2 qx = fma2(te, d.x, p.x), qy = fma2(te, d.y, p.y);
        const F2 rp = fma2(p.x, p.x, mul2(p.y, p.y)), rq = fma2(qx, qx, mul2(qy, qy));
        ia = rp.x < x.b[0] && rq.x < x.b[0];
        ib = rp.y < x.b[0] && rq.y < x.b[0];
    }
    return (da || ia) && (db || ib);
}

// ---- scalar twin of one row (rare cases and row kinds without a packed form): tile_test / tile_interact per lane ----
// DEP: dep(lane, slot, hit_local, weight).
template <class K, class DEP>
RTT_HD unsigned pair_row_scalar(const RowDev* rows, int r, P3& P, P3& D, F2& I, unsigned act, F2 mu_enter, F2 mu_exit,
                                PhysAux aux_a, PhysAux aux_b, DEP& dep) {
    V3 pa = lane_a(P), pb = lane_b(P), da = lane_a(D), db = lane_b(D);
    // both tests first: two independent dependency chains for the scheduler, as in the tile kernel
    float ta, tb;
    const bool ha = tile_test<K>(rows, r, pa, da, ta) && (act & 1u);
    const bool hb = tile_test<K>(rows, r, pb, db, tb) && (act & 2u);
    if (!(ha || hb)) return 0u;
    const RowDev& R = rows[r];
    float ia = I.x, ib = I.y;
    if (ha) {
        V3 np, nd, hl; float mod;
        tile_interact<K>(R, pa, da, ta, mu_enter.x, mu_exit.x, np, nd, mod, hl, aux_a);
        if (K::sensor(R)) dep(0, R.i[RTT_I_SENSOR], hl, ia);
        pa = np; da = nd; ia = ia * mod;
    }
    if (hb) {
        V3 np, nd, hl; float mod;
        tile_interact<K>(R, pb, db, tb, mu_enter.y, mu_exit.y, np, nd, mod, hl, aux_b);
        if (K::sensor(R)) dep(1, R.i[RTT_I_SENSOR], hl, ib);
        pb = np; db = nd; ib = ib * mod;
    }
    P = pack3(pa, pb); D = pack3(da, db); I = f2(ia, ib);
    return (ha ? 1u : 0u) | (hb ? 2u : 0u);
}

// ---- conic lens face, packed -----------------------------------------------------------------------------
// FULL: QUADRIC (x, y, z) / QUADRIC_ZY (y, z); SHAPE: RTT_SHAPE_SPHERIC_FACE / RTT_SHAPE_CYL_FACE; physics SNELL;
// Rs == I; HALF bound, not inverted (tile_opcode guarantees all of it).
template <bool FULL, int SHAPE, class DEP>
RTT_HD unsigned pair_conic_face(const RowDev* rows, int r, P3& P, P3& D, F2& I, unsigned act, F2 mu_enter, F2 mu_exit,
                                DEP& dep) {
    typedef KTile<FULL ? RTT_SURF_QUADRIC : RTT_SURF_QUADRIC_ZY, RTT_BOUND_HALF, SHAPE, RTT_PHYS_SNELL, 1, 0> K;
    const RowDev& R = rows[r];
    const float c = R.f[RTT_F_C], c1k = R.f[D_C1K];
    const float tc = 2.0f * c, tc1k = 2.0f * c1k;
    P3 o;                                                                // surface frame: o = pe - Ts, dd = de
    o.x = add2s(P.x, -R.f[RTT_F_TS]); o.y = add2s(P.y, -R.f[RTT_F_TS + 1]); o.z = add2s(P.z, -R.f[RTT_F_TS + 2]);
    const P3& d = D;
    // A, B, C of c (x^2 + y^2) + c (1 + k) z^2 - 2 z = 0 along the ray (geom/primitives.py:280-286, 356-376)
    F2 dd2, od, oo;
    if (FULL) {
        dd2 = fma2(d.x, d.x, mul2(d.y, d.y)); od = fma2(o.x, d.x, mul2(o.y, d.y)); oo = fma2(o.x, o.x, mul2(o.y, o.y));
    } else {
        dd2 = mul2(d.y, d.y); od = mul2(o.y, d.y); oo = mul2(o.y, o.y);
    }
    const F2 A = fma2s(dd2, c, mul2s(mul2(d.z, d.z), c1k));
    const F2 B = fma2s(d.z, -2.0f, fma2s(od, tc, mul2(mul2s(o.z, tc1k), d.z)));
    const F2 Cq = fma2s(o.z, -2.0f, fma2s(oo, c, mul2s(mul2(o.z, o.z), c1k)));
    const bool lin_a = fabsf(A.x) < 1e-6f, lin_b = fabsf(A.y) < 1e-6f;   // flat face / ray along a generator
    F2 t;
    bool va, vb;
    if (lin_a != lin_b)
        return pair_row_scalar<K>(rows, r, P, D, I, act, mu_enter, mu_exit, no_aux(), no_aux(), dep);
    if (lin_a) {                                                         // geom/primitives.py:305-313
        const F2 Bs = f2(fabsf(B.x) < 1e-6f ? 1e-6f : B.x, fabsf(B.y) < 1e-6f ? 1e-6f : B.y);
        t = mul2(mul2s(Cq, -1.0f), f2(rcp_(Bs.x), rcp_(Bs.y)));
        const F2 zc = mul2s(fma2(t, d.z, o.z), c);
        va = (t.x > 1e-6f) && (fabsf(zc.x) < 1.000001f);
        vb = (t.y > 1e-6f) && (fabsf(zc.y) < 1.000001f);
    } else {
        const F2 disc = fma2(B, B, mul2(mul2s(A, -4.0f), Cq));
        const F2 sq = f2(sqrt_(fabsf(disc.x)), sqrt_(fabsf(disc.y)));
        const F2 m2a = mul2s(A, -2.0f);
        const F2 ninv = f2(rcp_(m2a.x), rcp_(m2a.y));                    // -1 / (2A)
        const F2 r1 = mul2(add2(B, sq), ninv), r2 = mul2(sub2(B, sq), ninv);   // (-B - sq) / 2A, (-B + sq) / 2A
        const F2 lo = f2(fminf(r1.x, r2.x), fminf(r1.y, r2.y)), hi = f2(fmaxf(r1.x, r2.x), fmaxf(r1.y, r2.y));
        const F2 zlo = mul2s(fma2(lo, d.z, o.z), c), zhi = mul2s(fma2(hi, d.z, o.z), c);
        const bool oklo_a = (lo.x > 1e-6f) && (fabsf(zlo.x) < 1.000001f);
        const bool oklo_b = (lo.y > 1e-6f) && (fabsf(zlo.y) < 1.000001f);
        const bool okhi_a = (hi.x > 1e-6f) && (fabsf(zhi.x) < 1.000001f);
        const bool okhi_b = (hi.y > 1e-6f) && (fabsf(zhi.y) < 1.000001f);
        t = f2(oklo_a ? lo.x : hi.x, oklo_b ? lo.y : hi.y);
        va = (disc.x >= 0.0f) && (oklo_a || okhi_a);
        vb = (disc.y >= 0.0f) && (oklo_b || okhi_b);
    }
    va = va && (act & 1u); vb = vb && (act & 2u);
    if (!(va || vb)) return 0u;
    // shape-level rule on the element-frame hit point (geom/spherics.py:40-46, geom/cylindrics.py:31-37)
    const F2 hx = fma2(t, D.x, P.x), hy = fma2(t, D.y, P.y);
    if (SHAPE == RTT_SHAPE_SPHERIC_FACE) {
        const F2 rr = fma2(hx, hx, mul2(hy, hy));
        va = va && (rr.x <= R.f[D_HB0SQ]); vb = vb && (rr.y <= R.f[D_HB0SQ]);
    } else {
        const float* hb = R.f + RTT_F_HB;                                // slack pre-added by prepare_row
        va = va && (hx.x <= hb[1]) && (hx.x >= hb[0]) && (hy.x <= hb[3]) && (hy.x >= hb[2]);
        vb = vb && (hx.y <= hb[1]) && (hx.y >= hb[0]) && (hy.y <= hb[3]) && (hy.y >= hb[2]);
    }
    if (!(va || vb)) return 0u;
    // ---- interaction: hit point, normal (primitives.py:330-343, 378-395), Snell (phys/std.py:123-145) ----
    const P3 hl = along2(o, t, d);
    P3 g;                                                                // gradient of the implicit form
    g.x = FULL ? mul2s(hl.x, tc) : bc(0.0f);
    g.y = mul2s(hl.y, tc);
    g.z = fma2s(hl.z, tc1k, bc(-2.0f));
    const F2 l2 = FULL ? fma2(g.x, g.x, fma2(g.y, g.y, mul2(g.z, g.z))) : fma2(g.y, g.y, mul2(g.z, g.z));
    const bool deg = (va && !(l2.x > 1e-30f)) || (vb && !(l2.y > 1e-30f));
    const F2 ninvl = f2(-rsqrt_pk(l2.x), -rsqrt_pk(l2.y));
    P3 n;
    n.x = FULL ? mul2(g.x, ninvl) : bc(0.0f); n.y = mul2(g.y, ninvl); n.z = mul2(g.z, ninvl);
    const F2 dt = FULL ? fma2(d.z, n.z, fma2(d.y, n.y, mul2(d.x, n.x))) : fma2(d.z, n.z, mul2(d.y, n.y));
    const F2 mu = f2(dt.x < 0.0f ? mu_enter.x : mu_exit.x, dt.y < 0.0f ? mu_enter.y : mu_exit.y);
    const F2 one_m = fma2(mul2s(dt, -1.0f), dt, bc(1.0f));              // 1 - cos^2
    const F2 term = fma2(mul2(mul2s(mu, -1.0f), mu), one_m, bc(1.0f));  // 1 - mu^2 (1 - cos^2)
    const bool tir = (va && term.x < 0.0f) || (vb && term.y < 0.0f);
    if (deg || tir)
        return pair_row_scalar<K>(rows, r, P, D, I, act, mu_enter, mu_exit, no_aux(), no_aux(), dep);
    const F2 c2 = f2(sqrt_(fmaxf(term.x, 0.0f)), sqrt_(fmaxf(term.y, 0.0f)));
    // new_dir = mu d + q n with q = entering ? (mu |cos| - c2) : -(mu |cos| - c2)  ==  -mu cos -/+ c2
    const F2 cs = f2(dt.x < 0.0f ? -c2.x : c2.x, dt.y < 0.0f ? -c2.y : c2.y);
    const F2 q = fma2(mul2s(mu, -1.0f), dt, cs);
    P3 nd;
    nd.x = FULL ? fma2(q, n.x, mul2(mu, d.x)) : mul2(mu, d.x);
    nd.y = fma2(q, n.y, mul2(mu, d.y));
    nd.z = fma2(q, n.z, mul2(mu, d.z));
    // write back: a lane without a hit keeps its state (t = 0 moves nothing; direction selected)
    const F2 te = f2(va ? t.x : 0.0f, vb ? t.y : 0.0f);
    P = along2(P, te, D);
    D.x = f2(va ? nd.x.x : D.x.x, vb ? nd.x.y : D.x.y);
    D.y = f2(va ? nd.y.x : D.y.x, vb ? nd.y.y : D.y.y);
    D.z = f2(va ? nd.z.x : D.z.x, vb ? nd.z.y : D.z.y);
    return (va ? 1u : 0u) | (vb ? 2u : 0u);
}

// ---- bounded plane with Rs == I, packed: circular stop (APERTURE), disk / rectangle sensor (TRANSMIT) -----
// geom/primitives.py:124-136 (t = -o.z / d.z, |d.z| < 1e-6 -> 1e-8), geom/bounded.py:60-64, 77-82 with `invert`,
// t > 1e-6; phys/filter.py:24-33 / phys/std.py:227-235; elements/sensor.py:22-39.
template <int BOUND, int PHYS, bool SENSOR, class DEP>
RTT_HD unsigned pair_plane(const RowDev* rows, int r, P3& P, P3& D, F2& I, unsigned act, DEP& dep) {
    const RowDev& R = rows[r];
    const F2 ox = add2s(P.x, -R.f[RTT_F_TS]), oy = add2s(P.y, -R.f[RTT_F_TS + 1]), oz = add2s(P.z, -R.f[RTT_F_TS + 2]);
    const F2 safe = f2(fabsf(D.z.x) < 1e-6f ? 1e-8f : D.z.x, fabsf(D.z.y) < 1e-6f ? 1e-8f : D.z.y);
    const F2 t = mul2(mul2s(oz, -1.0f), f2(rcp_(safe.x), rcp_(safe.y)));
    const F2 hx = fma2(t, D.x, ox), hy = fma2(t, D.y, oy);
    const bool inv = R.i[RTT_I_INVERT] != 0;
    bool ka, kb;
    if (BOUND == RTT_BOUND_DISK) {
        const F2 rr = fma2(hx, hx, mul2(hy, hy));
        ka = rr.x <= R.f[D_SB0SQ]; kb = rr.y <= R.f[D_SB0SQ];
    } else {
        const float* sb = R.f + RTT_F_SB;
        ka = (fabsf(hx.x) <= sb[0]) && (fabsf(hy.x) <= sb[1]);
        kb = (fabsf(hx.y) <= sb[0]) && (fabsf(hy.y) <= sb[1]);
    }
    // NaN distances (0 * inf) fail `t > 1e-6` like the reference's mask-and-min; (k != inv) == (k xor inv)
    const bool va = (act & 1u) && (t.x > 1e-6f) && (t.x < rtt_inf()) && (ka != inv);
    const bool vb = (act & 2u) && (t.y > 1e-6f) && (t.y < rtt_inf()) && (kb != inv);
    if (!(va || vb)) return 0u;
    if (SENSOR) {
        const F2 hz = fma2(t, D.z, oz);
        const int slot = R.i[RTT_I_SENSOR];
        if (va) dep(0, slot, v3(hx.x, hy.x, hz.x), I.x);
        if (vb) dep(1, slot, v3(hx.y, hy.y, hz.y), I.y);
    }
    const F2 te = f2(va ? t.x : 0.0f, vb ? t.y : 0.0f);
    P = along2(P, te, D);
    if (PHYS == RTT_PHYS_APERTURE) {
        // filter.py:31-33: mask = inBounds(hit_local) WITHOUT the invert flag; a hit means (in-bounds xor invert)
        const float m = inv ? 0.0f : 1.0f;
        const F2 mm = f2(va ? m : 1.0f, vb ? m : 1.0f);
        D.x = mul2(D.x, mm); D.y = mul2(D.y, mm); D.z = mul2(D.z, mm);
        I = mul2(I, mm);
    }
    return (va ? 1u : 0u) | (vb ? 2u : 0u);
}

}  // namespace rtt
