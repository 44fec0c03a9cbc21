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
    return op;
}

// ---- frames -----------------------------------------------------------------------------------
// Frame of a row = its element pose (R_e, T_e); bare surfaces (shape NONE) live in the global frame.
// Xf maps coordinates of the previous row's frame to this row's frame:  p' = p @ M + c,  d' = d @ M.
struct Xf {
    float M[9];
    float c[3];
    int32_t kind;        // 0: same frame, 1: translation only (M == I), 2: general
    int32_t pad[3];
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
    x.pad[0] = x.pad[1] = x.pad[2] = 0;
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

// |d|^2 close enough to 1 (or exactly 0) for the per-row renormalisation to be the identity to 1e-6
RTT_HD bool regular_dir(V3 d) {
    const float l2 = fma3(d.x, d.x, d.y, d.y, d.z, d.z);
    return (l2 == 0.0f) || (fabsf(l2 - 1.0f) <= 4e-6f);
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
    const Roots q = solve_roots<K>(R, o, dd);
    int which;
    t = select_root<K>(R, q, o, dd, &which);
    bool valid = t < rtt_inf();
    if (K::shape(R) != RTT_SHAPE_NONE && valid) valid = shape_in_bounds<K>(rows, r, along(pe, t, de));
    return valid;
}

// The interaction of a ray known to hit row R at distance t: new position / direction in the row's
// frame, intensity factor, and the surface-frame hit point (sensor records, aperture physics).
template <class K = KDyn>
RTT_HD void tile_interact(const RowDev& R, V3 pe, V3 de, float t, float mu_enter, float mu_exit,
                          V3& np, V3& nd, float& mod, V3& hit_local) {
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
    nd = physics<K>(R, hit_local, de, n, mu_enter, mu_exit, &mod);
}

}  // namespace rtt
