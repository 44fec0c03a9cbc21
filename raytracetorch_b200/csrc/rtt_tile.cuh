// rtt_tile.cuh — per-ray arithmetic of the FAST sequential forward kernel ("frame-resident" form).
//
// Same algorithm as rtt_core.cuh (scene/sequential.py:12-36 row by row), reorganised for speed where
// the FAST variant's tolerance (points/directions 1e-5 relative, masks exact away from ties) allows:
//
//   * The ray state lives in the ELEMENT frame of the row being processed, not in the global frame.
//     All rows of one element share (R_e, T_e) (geom/shape.py:37-38), so the element pose is applied
//     once when the walk enters the element (one precomputed frame change `Xf` = previous frame ->
//     this frame), not once per row, and normals / new directions are never rotated back per row.
//     The last frame is mapped to the global frame once at the end.
//   * Directions are not renormalised per row (geom/shape.py:38 `F.normalize`): every physics of the
//     reference maps a unit (or zero) direction to a unit (or zero) direction up to rounding, so for
//     REGULAR rays (|d|^2 within 4e-6 of 1, or d == 0) the renormalisation is the identity to 1e-6.
//     Irregular rays (caller passed un-normalised directions) take the reference-order path of
//     rtt_core.cuh in the same kernel, so their results keep the reference's normalise-then-overshoot
//     behaviour (t measured along the unit direction, hit = p + t*d with the raw d).
//
// The surface-level pieces (solve_roots, select_root, surface/shape bounds, normal_local, physics) are
// the rtt_core.cuh functions, unchanged.  RTT_HD: also compiled by tests/hostsim for the CPU checks.
#pragma once
#include "rtt_core.cuh"

namespace rtt {

enum {
    DI_TILE_OP = 13   // index into RTT_TILE_SPECS (0 = generic)
};

// Row kinds with straight-line code in the frame-resident kernel: X(opcode, surface, surface bound,
// shape rule, physics, Rs == I, is-sensor).  The element rotation no longer matters per row.
#define RTT_TILE_SPECS(X)                                                                                     \
    X(1, RTT_SURF_QUADRIC, RTT_BOUND_HALF, RTT_SHAPE_SPHERIC_FACE, RTT_PHYS_SNELL, 1, 0)     /* lens face      */ \
    X(2, RTT_SURF_CYLINDER, RTT_BOUND_NONE, RTT_SHAPE_SPHERIC_EDGE, RTT_PHYS_BLOCK, 1, 0)    /* inked edge     */ \
    X(3, RTT_SURF_CYLINDER, RTT_BOUND_NONE, RTT_SHAPE_SPHERIC_EDGE, RTT_PHYS_SNELL, 1, 0)    /* clear edge     */ \
    X(4, RTT_SURF_QUADRIC_ZY, RTT_BOUND_HALF, RTT_SHAPE_CYL_FACE, RTT_PHYS_SNELL, 1, 0)      /* cyl. lens face */ \
    X(5, RTT_SURF_PLANE, RTT_BOUND_NONE, RTT_SHAPE_CYL_EDGE, RTT_PHYS_SNELL, 0, 0)           /* cyl. lens side */ \
    X(6, RTT_SURF_PLANE, RTT_BOUND_NONE, RTT_SHAPE_CYL_EDGE, RTT_PHYS_BLOCK, 0, 0)                                \
    X(7, RTT_SURF_PLANE, RTT_BOUND_DISK, RTT_SHAPE_NONE, RTT_PHYS_APERTURE, 1, 0)            /* circular stop  */ \
    X(8, RTT_SURF_PLANE, RTT_BOUND_DISK, RTT_SHAPE_NONE, RTT_PHYS_TRANSMIT, 1, 1)            /* disk sensor    */ \
    X(9, RTT_SURF_PLANE, RTT_BOUND_RECT, RTT_SHAPE_NONE, RTT_PHYS_TRANSMIT, 1, 1)            /* rect sensor    */ \
    X(10, RTT_SURF_QUADRIC, RTT_BOUND_HALF_DISK, RTT_SHAPE_NONE, RTT_PHYS_REFLECT, 1, 0)     /* sph. mirror    */ \
    X(11, RTT_SURF_PLANE, RTT_BOUND_NONE, RTT_SHAPE_POLY, RTT_PHYS_BLOCK, 0, 0)              /* box face       */

// policy of a tile spec: only bit 1 of ident (Rs == I) is consulted by the tile functions
template <int SURF, int BOUND, int SHAPE, int PHYS, int RS_IDENT, int SENSOR>
using KTile = KStatic<SURF, BOUND, SHAPE, PHYS, (RS_IDENT ? 3 : 1), SENSOR>;

RTT_HD int tile_opcode(const RowDev& R) {
    int op = 0;
    const int rs = (R.i[DI_IDENT] & 2) ? 1 : 0;
#define RTT_X(OP, SURF, BOUND, SHAPE, PHYS, RS_IDENT, SENSOR)                                          \
    if (R.i[RTT_I_SURF] == SURF && R.i[RTT_I_BOUND] == BOUND && R.i[RTT_I_SHAPE] == SHAPE &&        \
        R.i[RTT_I_PHYS] == PHYS && rs == RS_IDENT && (R.i[RTT_I_SENSOR] >= 0) == (SENSOR != 0)) op = OP;
    RTT_TILE_SPECS(RTT_X)
#undef RTT_X
    // specialised HALF-bounded conics assume the usual, non-inverted bound (lean root selection); an inverted one
    // takes the generic path
    if (R.i[RTT_I_BOUND] == RTT_BOUND_HALF && R.i[RTT_I_INVERT] != 0) op = 0;
#if defined(RTT_EXPERIMENT_GENERIC_ROWS)
    op = 0;
#endif
    return op;
}

// ---- frames -----------------------------------------------------------------------------------
// Frame of a row = its element pose (R_e, T_e); bare surfaces (shape NONE) live in the global frame.
// Xf maps coordinates of the previous row's frame to this row's frame:  p' = p @ M + c,  d' = d @ M.
struct Xf {
    float M[9];
    float c[3];
    int32_t kind;        // 0: same frame, 1: translation only (M == I), 2: general
    // lens-edge culling (see below): rows [r, r + run) can be skipped for a ray that passes the test
    int32_t run;         // 0: row r is not the start / inside of a cullable run
    float zlo, zhi;      // enclosure of the z range the edge rows accept (element frame)
    int32_t ctype;       // 1: z test only, 2: + rectangular cross-section b = (x0, x1, y0, y1), 3: + disc, b[0] = r^2
    float b[4];
    float q[4];          // ctype 2: the rectangle b as centre / half extents (cx, hx, cy, hy) for the packed cull test
    int32_t ctl;         // pair kernel: opcode | kind << 8 | run << 16 (set when the tile is staged)
    int32_t pad[1];
};

struct FramePose { float R[9]; float T[3]; };

RTT_HD FramePose frame_pose(const RowDev* R) {           // nullptr = global frame
    FramePose f;
    for (int a = 0; a < 9; ++a) f.R[a] = (a % 4 == 0) ? 1.0f : 0.0f;
    f.T[0] = f.T[1] = f.T[2] = 0.0f;
    if (R && R->i[RTT_I_SHAPE] != RTT_SHAPE_NONE) {
        for (int a = 0; a < 9; ++a) f.R[a] = R->f[RTT_F_RE + a];
        for (int a = 0; a < 3; ++a) f.T[a] = R->f[RTT_F_TE + a];
    }
    return f;
}

// p_new = ((p_old @ Ro^T + To) - Tn) @ Rn = p_old @ (Ro^T Rn) + (To - Tn) @ Rn
RTT_HD Xf make_xf(const RowDev* from, const RowDev* to) {
    const FramePose o = frame_pose(from), n = frame_pose(to);
    Xf x;
    x.run = 0; x.zlo = 0.0f; x.zhi = 0.0f; x.ctype = 0;
    x.b[0] = x.b[1] = x.b[2] = x.b[3] = 0.0f; x.q[0] = x.q[1] = x.q[2] = x.q[3] = 0.0f; x.ctl = 0; x.pad[0] = 0;
    bool same_R = true, same_T = true, o_ident = true, n_ident = true;
    for (int a = 0; a < 9; ++a) {
        same_R = same_R && (o.R[a] == n.R[a]);
        const float want = (a % 4 == 0) ? 1.0f : 0.0f;
        o_ident = o_ident && (o.R[a] == want);
        n_ident = n_ident && (n.R[a] == want);
    }
    for (int a = 0; a < 3; ++a) same_T = same_T && (o.T[a] == n.T[a]);
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            // (Ro^T Rn)[i][j] = sum_k Ro[k][i] Rn[k][j]; exact when either factor is the identity
            x.M[3 * i + j] = o_ident ? n.R[3 * i + j]
                           : (n_ident ? o.R[3 * j + i]
                                      : fma3(o.R[i], n.R[j], o.R[3 + i], n.R[3 + j], o.R[6 + i], n.R[6 + j]));
        }
    const V3 dT = v3(o.T[0] - n.T[0], o.T[1] - n.T[1], o.T[2] - n.T[2]);
    const V3 c = n_ident ? dT : mul_R(dT, n.R);
    x.c[0] = c.x; x.c[1] = c.y; x.c[2] = c.z;
    if (same_R) {                                         // R^T R == I: a pure shift in the shared frame
        for (int a = 0; a < 9; ++a) x.M[a] = (a % 4 == 0) ? 1.0f : 0.0f;
        x.kind = same_T ? 0 : 1;
    } else {
        x.kind = 2;
    }
    return x;
}

RTT_HD void apply_xf(const Xf& x, V3& p, V3& d) {
    if (x.kind == 0) return;
    const V3 c = ld3(x.c);
    if (x.kind == 1) { p = p + c; return; }
    p = mul_R(p, x.M) + c;
    d = mul_R(d, x.M);
}

// the same with the kind taken from the row's control word (kind != 0; general = kind is 2)
RTT_HD void apply_xf_kind(const Xf& x, bool general, V3& p, V3& d) {
    const V3 c = ld3(x.c);
    if (!general) { p = p + c; return; }
    p = mul_R(p, x.M) + c;
    d = mul_R(d, x.M);
}

// ---- lens-edge culling ---------------------------------------------------------------------------
// The edge rows of a lens (the cylinder of a spherical lens, geom/spherics.py:34-39; the four side planes
// of a cylindrical lens, geom/cylindrics.py:38-55) accept a hit only if its element-frame z lies between
// the two faces, i.e. inside an enclosure [zlo, zhi].  Two result-preserving shortcuts:
//   (z)  a ray that is beyond the enclosure and does not travel back towards it fails that rule for
//        every t > 1e-6 (h.z = p.z + t d.z is monotone in t);
//   (xy) the edge surfaces lie on / outside the boundary of the lens cross-section (rectangle or disc).
//        If the ray's xy position now and at the moment it leaves the z enclosure are both strictly inside
//        the (slightly shrunken) cross-section, the segment between them — the cross-section is convex —
//        crosses no edge surface, so any edge hit lies beyond the enclosure and fails the z rule.
// Either way the rows of the run report "no hit" exactly as the full test would; the kernel skips them
// when every lane of the warp passes.  Margins (1e-3) are far above fp32 rounding at scene scale.
constexpr float kCullMargin = 1e-3f;

struct EdgeInfo { float zlo, zhi; int ctype; float b[4]; };

RTT_HD bool edge_row_info(const RowDev& R, EdgeInfo& e) {
    const float* hb = R.f + RTT_F_HB;                      // after prepare_row (cyl. aperture slack pre-added)
    e.ctype = 1; e.b[0] = e.b[1] = e.b[2] = e.b[3] = 0.0f;
    if (R.i[RTT_I_SHAPE] == RTT_SHAPE_SPHERIC_EDGE) {
        e.zlo = hb[0] - kCullMargin; e.zhi = hb[1] + kCullMargin;
        if (R.i[RTT_I_SURF] == RTT_SURF_CYLINDER && (R.i[DI_IDENT] & 2) && R.f[RTT_F_TS] == 0.0f &&
            R.f[RTT_F_TS + 1] == 0.0f && R.f[RTT_F_RADIUS] > 4.0f * kCullMargin) {
            const float rin = R.f[RTT_F_RADIUS] - 2.0f * kCullMargin;
            e.ctype = 3; e.b[0] = rin * rin;
        }
        return true;
    }
    if (R.i[RTT_I_SHAPE] == RTT_SHAPE_CYL_EDGE) {
        const float y0 = hb[2], y1 = hb[3];
        const float ya = fabsf(y0) > fabsf(y1) ? y0 : y1;                       // largest |y| of the aperture
        const float yi = (y0 <= 0.0f && y1 >= 0.0f) ? 0.0f : (fabsf(y0) < fabsf(y1) ? y0 : y1);   // smallest |y|
        // sag is monotone in |y|: its extremes over the aperture sit at the two ends
        const float f0 = sag_at(hb[4], yi, hb[5]), f1 = sag_at(hb[4], ya, hb[5]);
        const float b0 = sag_at(hb[6], yi, hb[7]), b1 = sag_at(hb[6], ya, hb[7]);
        e.zlo = fminf(f0, f1) - kCullMargin; e.zhi = fmaxf(b0, b1) + kCullMargin;
        if (!(e.zlo == e.zlo && e.zhi == e.zhi)) return false;
        if (R.i[RTT_I_SURF] == RTT_SURF_PLANE) {
            // the plane (normal = third column of Rs, through Ts) must be vertical and must not cut the shrunken
            // aperture rectangle: its four corners lie on one side
            const float nx = R.f[RTT_F_RS + 2], ny = R.f[RTT_F_RS + 5], nz = R.f[RTT_F_RS + 8];
            const float m = 2.0f * kCullMargin;
            const float x0 = hb[0] + m, x1 = hb[1] - m, yy0 = hb[2] + m, yy1 = hb[3] - m;
            if (fabsf(nz) <= 1e-6f && x0 < x1 && yy0 < yy1) {
                const float tx = R.f[RTT_F_TS], ty = R.f[RTT_F_TS + 1];
                const float s00 = nx * (x0 - tx) + ny * (yy0 - ty), s01 = nx * (x0 - tx) + ny * (yy1 - ty);
                const float s10 = nx * (x1 - tx) + ny * (yy0 - ty), s11 = nx * (x1 - tx) + ny * (yy1 - ty);
                const bool neg = s00 < 0.0f && s01 < 0.0f && s10 < 0.0f && s11 < 0.0f;
                const bool pos = s00 > 0.0f && s01 > 0.0f && s10 > 0.0f && s11 > 0.0f;
                if (neg || pos) { e.ctype = 2; e.b[0] = x0; e.b[1] = x1; e.b[2] = yy0; e.b[3] = yy1; }
            }
        }
        return true;
    }
    return false;
}

// Fills xf[r].run / zlo / zhi / ctype / b for the run of cullable rows that starts at r (bounds = union over the
// rest of the run, cross-section test only if every row of the run agrees on it).  Call after make_xf on all rows.
RTT_HD void edge_run_at(const RowDev* rows, int S, Xf* xf, int r) {
    EdgeInfo e;
    if (!edge_row_info(rows[r], e)) return;
    int run = 1;
    for (int m = r + 1; m < S && xf[m].kind == 0; ++m, ++run) {
        EdgeInfo e2;
        if (!edge_row_info(rows[m], e2)) break;
        e.zlo = fminf(e.zlo, e2.zlo); e.zhi = fmaxf(e.zhi, e2.zhi);
        const bool same = e.ctype == e2.ctype && e.b[0] == e2.b[0] && e.b[1] == e2.b[1] && e.b[2] == e2.b[2] &&
                          e.b[3] == e2.b[3];
        if (!same) e.ctype = 1;
    }
    xf[r].run = run; xf[r].zlo = e.zlo; xf[r].zhi = e.zhi; xf[r].ctype = e.ctype;
    for (int a = 0; a < 4; ++a) xf[r].b[a] = e.b[a];
    xf[r].q[0] = 0.5f * (e.b[0] + e.b[1]); xf[r].q[1] = 0.5f * (e.b[1] - e.b[0]);
    xf[r].q[2] = 0.5f * (e.b[2] + e.b[3]); xf[r].q[3] = 0.5f * (e.b[3] - e.b[2]);
}

RTT_HD bool edge_culled(const Xf& x, V3 p, V3 d) {
    // bitwise `&` / `|` on the compares throughout: predicate chains, no short-circuit branches (per-lane data)
    if (((p.z > x.zhi) & (d.z >= 0.0f)) | ((p.z < x.zlo) & (d.z <= 0.0f))) return true;       // (z)
    if (x.ctype < 2 || d.z == 0.0f) return false;
    // (xy): time at which the ray leaves the z enclosure, stretched a little (a longer segment is conservative)
    const float zt = d.z > 0.0f ? x.zhi : x.zlo;
    const float te = fmaf(div_(zt - p.z, d.z), 1.0001f, 1e-4f);
    const float qx = fmaf(te, d.x, p.x), qy = fmaf(te, d.y, p.y);
    if (x.ctype == 2)
        return (p.x > x.b[0]) & (p.x < x.b[1]) & (p.y > x.b[2]) & (p.y < x.b[3]) &
               (qx > x.b[0]) & (qx < x.b[1]) & (qy > x.b[2]) & (qy < x.b[3]);
    return ((p.x * p.x + p.y * p.y) < x.b[0]) & ((qx * qx + qy * qy) < x.b[0]);
}

// ---- box culling (non-sequential search) ----------------------------------------------------------
// A face row of a convex polyhedron accepts a hit only if the hit point lies within 1e-4 of the inside of every
// sibling half-space (geom/shape.py:122-132), i.e. inside the slightly inflated solid.  For a box (three
// antiparallel, mutually orthogonal pairs of face planes) that solid fits in a sphere, so a ray whose line stays
// outside the sphere — or which starts outside and points away — cannot hit any of the six rows.  The sphere is
// derived from the face planes themselves and kept in the GLOBAL frame (the non-sequential state's frame).
struct NsCull {
    int32_t run;             // > 0 at the first row of a verified box: rows [r, r + run) can be skipped together
    float cx, cy, cz, r2;    // bounding sphere, global frame
    float ex, ey, ez;        // its centre in the element frame (shape_in_bounds works there)
};

RTT_HD NsCull box_cull_info(const RowDev* rows, int S, int r) {
    NsCull c;
    c.run = 0; c.cx = c.cy = c.cz = c.r2 = 0.0f; c.ex = c.ey = c.ez = 0.0f;
    const RowDev& R0 = rows[r];
    if (R0.i[RTT_I_SHAPE] != RTT_SHAPE_POLY || R0.i[RTT_I_POLY_FIRST] != r || R0.i[RTT_I_POLY_COUNT] != 6 || r + 6 > S)
        return c;
    V3 n[6]; float off[6];
    for (int f = 0; f < 6; ++f) {
        const RowDev& R = rows[r + f];
        if (R.i[RTT_I_SURF] != RTT_SURF_PLANE || R.i[RTT_I_SHAPE] != RTT_SHAPE_POLY || R.i[RTT_I_POLY_FIRST] != r ||
            R.i[RTT_I_BOUND] != RTT_BOUND_NONE || R.i[RTT_I_PHYS] == RTT_PHYS_LINEAR) return c;
        n[f] = v3(R.f[RTT_F_RS + 2], R.f[RTT_F_RS + 5], R.f[RTT_F_RS + 8]);     // plane normal (third column of Rs)
        off[f] = dot(n[f], ld3(R.f + RTT_F_TS));                                 // signed offset of the plane
    }
    // pair the faces: partner = the antiparallel one
    int axis_a[3], axis_b[3], na = 0;
    bool used[6] = {false, false, false, false, false, false};
    for (int f = 0; f < 6; ++f) {
        if (used[f]) continue;
        int partner = -1;
        for (int g = f + 1; g < 6; ++g)
            if (!used[g] && dot(n[f], n[g]) < -0.99999f) { partner = g; break; }
        if (partner < 0 || na == 3) return c;
        used[f] = used[partner] = true;
        axis_a[na] = f; axis_b[na] = partner; ++na;
    }
    if (na != 3) return c;
    for (int a = 0; a < 3; ++a) {
        const float len = norm3(n[axis_a[a]].x, n[axis_a[a]].y, n[axis_a[a]].z);
        if (fabsf(len - 1.0f) > 1e-5f) return c;
        for (int b = a + 1; b < 3; ++b)
            if (fabsf(dot(n[axis_a[a]], n[axis_a[b]])) > 1e-5f) return c;
    }
    // centre and half extents along the three axes (element frame)
    V3 ce = v3(0.0f, 0.0f, 0.0f);
    float h2 = 0.0f;
    for (int a = 0; a < 3; ++a) {
        const V3 na_ = n[axis_a[a]];
        const float o1 = off[axis_a[a]], o2 = -off[axis_b[a]];                   // partner's offset along +n_a
        const float mid = 0.5f * (o1 + o2), half = 0.5f * fabsf(o1 - o2);
        ce = ce + mid * na_;
        h2 += half * half;
    }
    const float rad = sqrt_(h2) * 1.01f + 2e-3f;                                 // inflated: 1e-4 slack + rounding
    const V3 cg = mul_RT(ce, R0.f + RTT_F_RE) + ld3(R0.f + RTT_F_TE);            // element -> global
    if (!(rad == rad) || !((cg.x + cg.y + cg.z) - (cg.x + cg.y + cg.z) == 0.0f)) return c;
    c.run = 6; c.cx = cg.x; c.cy = cg.y; c.cz = cg.z; c.r2 = rad * rad;
    c.ex = ce.x; c.ey = ce.y; c.ez = ce.z;
    return c;
}

// true iff the ray (p, d) — d of any length, possibly zero — cannot touch the sphere
RTT_HD bool sphere_missed(const NsCull& c, V3 p, V3 d) {
    const V3 m = v3(c.cx - p.x, c.cy - p.y, c.cz - p.z);
    const float mm = dot(m, m), md = dot(m, d), dd = dot(d, d);
    if (mm > c.r2 && md <= 0.0f) return true;                                    // outside and not approaching
    return (mm * dd - md * md) > c.r2 * dd * 1.001f;                             // line passes outside
}

// A NaN / inf coordinate reaches every component of the reference's `[N,3] @ [3,3]` poses (NaN * 0 = NaN), so such
// a ray hits nothing and leaves the trace untouched; the kernels skip identity rotations, hence the explicit test.
RTT_HD bool finite_ray(V3 p, V3 d) {
    const float s = ((p.x + p.y) + p.z) + ((d.x + d.y) + d.z);
    return (s - s) == 0.0f;
}

// |d|^2 close enough to 1 (or exactly 0) for the per-row renormalisation to be the identity to 1e-6
RTT_HD bool regular_dir(V3 d) {
    const float l2 = fma3(d.x, d.x, d.y, d.y, d.z, d.z);
    return (l2 == 0.0f) || (fabsf(l2 - 1.0f) <= 4e-6f);
}

// ---- lean root selection for lens faces (RTT_TILE_LEAN: FAST builds) ------------------------------
// A HALF-bounded, non-inverted conic (every lens face): the reference masks each of the two roots with
// (t > 1e-6) & (|z c| < 1 + 1e-6) and takes the smaller survivor (geom/bounded.py:20-36).  Same decision, fewer
// selects: order the roots, test both, keep the lower valid one; the A ~ 0 fallback (geom/primitives.py:305-313) is a
// rarely taken branch instead of a blend.  Root values are those of solve_roots (identical expressions).
template <class K>
RTT_HD bool conic_half_hit(const RowDev& R, V3 o, V3 d, float& t) {
    const float c = R.f[RTT_F_C], c1k = R.f[D_C1K];
    const float tc = 2.0f * c, tc1k = 2.0f * c1k;
    float A, B, Cq;
    if (K::surf(R) == RTT_SURF_QUADRIC) {
        A = c * (d.x * d.x + d.y * d.y) + c1k * (d.z * d.z);
        B = (tc * (o.x * d.x + o.y * d.y) + (tc1k * o.z) * d.z) - 2.0f * d.z;
        Cq = (c * (o.x * o.x + o.y * o.y) + c1k * (o.z * o.z)) - 2.0f * o.z;
    } else {
        A = c * (d.y * d.y) + c1k * (d.z * d.z);
        B = (tc * (o.y * d.y) + (tc1k * o.z) * d.z) - 2.0f * d.z;
        Cq = (c * (o.y * o.y) + c1k * (o.z * o.z)) - 2.0f * o.z;
    }
    if (fabsf(A) < 1e-6f) {                                             // flat face / ray along a generator
        const float Bs = (fabsf(B) < 1e-6f) ? 1e-6f : B;
        t = div_(-Cq, Bs);
        return (t > 1e-6f) & (fabsf(fmaf(t, d.z, o.z) * c) < 1.000001f);
    }
    const float disc = B * B - (4.0f * A) * Cq;
    const float sq = sqrt_(fabsf(disc));
    const float inv = rcp_(2.0f * A);
    const float r1 = (-B - sq) * inv, r2 = (-B + sq) * inv;
    const float lo = fminf(r1, r2), hi = fmaxf(r1, r2);
    const bool oklo = (lo > 1e-6f) & (fabsf(fmaf(lo, d.z, o.z) * c) < 1.000001f);
    const bool okhi = (hi > 1e-6f) & (fabsf(fmaf(hi, d.z, o.z) * c) < 1.000001f);
    t = oklo ? lo : hi;
    return (disc >= 0.0f) & (oklo | okhi);
}

// ---- one row ------------------------------------------------------------------------------------
// Distance along the ray with every validity rule of the sequential walk (surface bounds, t > 1e-6,
// shape-level rule); (pe, de) are in the row's frame.  Returns true iff the ray interacts with the row.
template <class K = KDyn>
RTT_HD bool tile_test(const RowDev* rows, int r, V3 pe, V3 de, float& t) {
    const RowDev& R = rows[r];
    const bool rs_ident = (K::ident(R) & 2) != 0;
    V3 o, dd;
    if (K::surf(R) == RTT_SURF_PLANE && !rs_ident) {
        // a plane only needs the z components of the surface-frame ray (third column of Rs)
        const V3 q = pe - ld3(R.f + RTT_F_TS);
        const float* Rs = R.f + RTT_F_RS;
        o = v3(0.0f, 0.0f, fma3(q.x, Rs[2], q.y, Rs[5], q.z, Rs[8]));
        dd = v3(0.0f, 0.0f, fma3(de.x, Rs[2], de.y, Rs[5], de.z, Rs[8]));
        if (K::bound(R) != RTT_BOUND_NONE) {              // bounded planes test x, y of the local hit
            o = rot_fwd(q, Rs, false);
            dd = rot_fwd(de, Rs, false);
        }
    } else {
        o = rot_fwd(pe - ld3(R.f + RTT_F_TS), R.f + RTT_F_RS, rs_ident);
        dd = rot_fwd(de, R.f + RTT_F_RS, rs_ident);
    }
    bool valid;
#if defined(RTT_TILE_LEAN)
    if (K::specialised() && (K::surf(R) == RTT_SURF_QUADRIC || K::surf(R) == RTT_SURF_QUADRIC_ZY) &&
        K::bound(R) == RTT_BOUND_HALF) {                                // tile_opcode: never an inverted bound here
        valid = conic_half_hit<K>(R, o, dd, t);
    } else
#endif
    {
        const Roots q = solve_roots<K>(R, o, dd);
        int which;
        t = select_root<K>(R, q, o, dd, &which);
        valid = t < rtt_inf();
    }
    if (K::shape(R) != RTT_SHAPE_NONE && valid) valid = shape_in_bounds<K>(rows, r, along(pe, t, de));
    return valid;
}

// The interaction of a ray known to hit row R at distance t: new position / direction in the row's
// frame, intensity factor, and the surface-frame hit point (sensor records, aperture physics).
template <class K = KDyn>
RTT_HD void tile_interact(const RowDev& R, V3 pe, V3 de, float t, float mu_enter, float mu_exit,
                          V3& np, V3& nd, float& mod, V3& hit_local, PhysAux aux = no_aux()) {
    const bool rs_ident = (K::ident(R) & 2) != 0;
    const V3 o = rot_fwd(pe - ld3(R.f + RTT_F_TS), R.f + RTT_F_RS, rs_ident);
    const V3 dd = rot_fwd(de, R.f + RTT_F_RS, rs_ident);
    hit_local = along(o, t, dd);                                         // primitives.py:81
    float nlen;
    const V3 nl = normal_local<K>(R, hit_local, &nlen);
    const V3 n = (K::surf(R) == RTT_SURF_PLANE && !rs_ident)
                     ? v3(R.f[RTT_F_RS + 2], R.f[RTT_F_RS + 5], R.f[RTT_F_RS + 8])
                     : rot_bwd(nl, R.f + RTT_F_RS, rs_ident);            // element-frame normal
    np = along(pe, t, de);                                               // shape.py:81 in the element frame
    nd = physics<K>(R, hit_local, de, n, mu_enter, mu_exit, &mod, aux);
}

}  // namespace rtt
