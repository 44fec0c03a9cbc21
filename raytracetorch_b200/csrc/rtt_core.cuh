// rtt_core.cuh — per-ray arithmetic of the ray-propagation path (forward and adjoint).
//
// Pure functions of (table row, ray state): no memory traffic policy, no launch logic —
// that lives in rtt_kernels.inl.  Everything is RTT_HD so the same source also compiles
// for the host in tests/hostsim (a development checker that runs the kernel arithmetic on
// the CPU against the torch oracle; it is never loaded by the product package).
//
// The formulas restate the reference step by step (file:line under /root/reference are
// given at each function) and keep its operation ORDER, because in EXACT mode (compiled
// with -fmad=false) every fp32 rounding step of the reference's eager ops is reproduced:
// hit masks, root selection and sensor bins then agree bit for bit on identity-rotation
// scenes.  FAST mode is the same source with FMA contraction enabled.
#pragma once

#include <stdint.h>
#include <math.h>
#include "../../include/rtt_b200.h"

#if defined(__CUDACC__)
#define RTT_HD __host__ __device__ __forceinline__
#else
#define RTT_HD inline
#endif

namespace rtt {

// Division / square root.  EXACT: IEEE (div.rn / sqrt.rn), like the reference's eager ops.  FAST
// (RTT_APPROX, device only): MUFU approximations (rcp / sqrt .approx.ftz, <= 2 ulp), one reciprocal
// shared by the three components of a vector; IEEE fp32 division costs ~10 issue slots plus a
// slow-path CALL for zero / inf / denormal operands, which dead rays (dir == 0) and missed roots
// (t == inf) hit all the time: ~40 % of the forward kernel's instructions in the first profile
// (profiles/r1_c2_seq_fwd_baseline.md).  Special values behave like IEEE where the algorithm
// relies on them: x*rcp(0) = inf or NaN (0*inf), which keeps "axis-parallel ray on a cylinder edge
// => NaN => miss" (geom/primitives.py:212-231).
#if defined(RTT_APPROX) && defined(__CUDA_ARCH__)
RTT_HD float rcp_(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
RTT_HD float sqrt_(float x) { float r; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
RTT_HD float div_(float a, float b) { return a * rcp_(b); }
RTT_HD float rsqrt_(float x) { float r; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
#elif defined(RTT_OUTLINE_IEEE) && defined(__CUDA_ARCH__)
// EXACT device build with the IEEE division / square root sequences (~15 instructions + a slow-path call each, at
// ~60 sites of the non-sequential kernel) behind ONE out-of-line copy each: smaller code for the instruction cache.
__device__ __noinline__ float rtt_ieee_div(float a, float b) { return a / b; }
__device__ __noinline__ float rtt_ieee_sqrt(float x) { return sqrtf(x); }
RTT_HD float rcp_(float x) { return rtt_ieee_div(1.0f, x); }
RTT_HD float sqrt_(float x) { return rtt_ieee_sqrt(x); }
RTT_HD float div_(float a, float b) { return rtt_ieee_div(a, b); }
#else
RTT_HD float rcp_(float x) { return 1.0f / x; }
RTT_HD float sqrt_(float x) { return sqrtf(x); }
RTT_HD float div_(float a, float b) { return a / b; }
#endif

// ---- row record as staged in shared memory ------------------------------------------------
// f[0..40] is the caller's table row; f[41..47] and i[11] are derived once per block.
enum {
    D_C1K = 41,       // c*(1+k)                 geom/primitives.py:280
    D_MU_ENTER = 42,  // ior_out/ior_in          phys/std.py:132
    D_MU_EXIT = 43,   // ior_in/ior_out
    D_R2 = 44,        // radius**2               geom/primitives.py:161,214
    D_SB0SQ = 45,     // sb[0]**2                geom/bounded.py:64,157
    D_HB0SQ = 46,     // hb[0]**2                geom/spherics.py:44
    D_ACC_SLOT = 47,  // sequential adjoint: index of the row's private gradient slots (as a float), -1 = none
    D_SAME_ELEM = 47, // non-sequential forward (same slot, other kernel): 1 = same element pose as the previous row
    DI_IDENT = 11,    // bit0: Re == I, bit1: Rs == I (exact compare)
    DI_OPCODE = 12    // index of the matching KStatic specialisation (RTT_ROW_SPECS), 0 = generic
};

// Row kinds that get straight-line code: X(opcode, surface, surface bound, shape rule, physics, ident, is-sensor).
// These cover every row of the reference's lens / stop / sensor elements in the poses users
// build them with (untilted, or tilted as a whole element); anything else runs the generic path.
#define RTT_ROW_SPECS(X)                                                                                \
    X(1, RTT_SURF_QUADRIC, RTT_BOUND_HALF, RTT_SHAPE_SPHERIC_FACE, RTT_PHYS_SNELL, 3, 0)    /* lens face        */ \
    X(2, RTT_SURF_QUADRIC, RTT_BOUND_HALF, RTT_SHAPE_SPHERIC_FACE, RTT_PHYS_SNELL, 2, 0)    /* tilted lens face */ \
    X(3, RTT_SURF_CYLINDER, RTT_BOUND_NONE, RTT_SHAPE_SPHERIC_EDGE, RTT_PHYS_BLOCK, 3, 0)   /* inked lens edge  */ \
    X(4, RTT_SURF_CYLINDER, RTT_BOUND_NONE, RTT_SHAPE_SPHERIC_EDGE, RTT_PHYS_BLOCK, 2, 0)                          \
    X(5, RTT_SURF_CYLINDER, RTT_BOUND_NONE, RTT_SHAPE_SPHERIC_EDGE, RTT_PHYS_SNELL, 3, 0)   /* clear lens edge  */ \
    X(6, RTT_SURF_QUADRIC_ZY, RTT_BOUND_HALF, RTT_SHAPE_CYL_FACE, RTT_PHYS_SNELL, 3, 0)     /* cyl. lens face   */ \
    X(7, RTT_SURF_QUADRIC_ZY, RTT_BOUND_HALF, RTT_SHAPE_CYL_FACE, RTT_PHYS_SNELL, 2, 0)                            \
    X(8, RTT_SURF_PLANE, RTT_BOUND_NONE, RTT_SHAPE_CYL_EDGE, RTT_PHYS_SNELL, 1, 0)          /* cyl. lens side   */ \
    X(9, RTT_SURF_PLANE, RTT_BOUND_NONE, RTT_SHAPE_CYL_EDGE, RTT_PHYS_SNELL, 0, 0)                                 \
    X(10, RTT_SURF_PLANE, RTT_BOUND_NONE, RTT_SHAPE_CYL_EDGE, RTT_PHYS_BLOCK, 1, 0)                                \
    X(11, RTT_SURF_PLANE, RTT_BOUND_NONE, RTT_SHAPE_CYL_EDGE, RTT_PHYS_BLOCK, 0, 0)                                \
    X(12, RTT_SURF_PLANE, RTT_BOUND_DISK, RTT_SHAPE_NONE, RTT_PHYS_APERTURE, 3, 0)          /* circular stop    */ \
    X(13, RTT_SURF_PLANE, RTT_BOUND_DISK, RTT_SHAPE_NONE, RTT_PHYS_TRANSMIT, 3, 1)          /* disk sensor      */ \
    X(14, RTT_SURF_PLANE, RTT_BOUND_RECT, RTT_SHAPE_NONE, RTT_PHYS_TRANSMIT, 3, 1)          /* rect sensor      */ \
    X(15, RTT_SURF_QUADRIC, RTT_BOUND_HALF_DISK, RTT_SHAPE_NONE, RTT_PHYS_REFLECT, 3, 0)    /* spherical mirror */ \
    X(16, RTT_SURF_PLANE, RTT_BOUND_NONE, RTT_SHAPE_POLY, RTT_PHYS_BLOCK, 1, 0)             /* box face         */

// Subset of RTT_ROW_SPECS that also gets a specialised ADJOINT (code size: the reverse sweep is ~3x
// the forward interaction).  Absorbing rows (inked edges, box faces) kill the ray, so live-ray
// gradients rarely cross them; they and every other kind use the generic adjoint.
#if defined(RTT_EXPERIMENT_ADJ_MIN)
#define RTT_ROW_SPECS_ADJ(X)                                                                            \
    X(1, RTT_SURF_QUADRIC, RTT_BOUND_HALF, RTT_SHAPE_SPHERIC_FACE, RTT_PHYS_SNELL, 3, 0)                           \
    X(13, RTT_SURF_PLANE, RTT_BOUND_DISK, RTT_SHAPE_NONE, RTT_PHYS_TRANSMIT, 3, 1)
#else
#define RTT_ROW_SPECS_ADJ(X)                                                                            \
    X(1, RTT_SURF_QUADRIC, RTT_BOUND_HALF, RTT_SHAPE_SPHERIC_FACE, RTT_PHYS_SNELL, 3, 0)                           \
    X(2, RTT_SURF_QUADRIC, RTT_BOUND_HALF, RTT_SHAPE_SPHERIC_FACE, RTT_PHYS_SNELL, 2, 0)                           \
    X(6, RTT_SURF_QUADRIC_ZY, RTT_BOUND_HALF, RTT_SHAPE_CYL_FACE, RTT_PHYS_SNELL, 3, 0)                            \
    X(7, RTT_SURF_QUADRIC_ZY, RTT_BOUND_HALF, RTT_SHAPE_CYL_FACE, RTT_PHYS_SNELL, 2, 0)                            \
    X(12, RTT_SURF_PLANE, RTT_BOUND_DISK, RTT_SHAPE_NONE, RTT_PHYS_APERTURE, 3, 0)                                 \
    X(13, RTT_SURF_PLANE, RTT_BOUND_DISK, RTT_SHAPE_NONE, RTT_PHYS_TRANSMIT, 3, 1)                                 \
    X(14, RTT_SURF_PLANE, RTT_BOUND_RECT, RTT_SHAPE_NONE, RTT_PHYS_TRANSMIT, 3, 1)
#endif

// ---- row-kind policies -----------------------------------------------------------------------
// Every per-row function below is a template over a policy K that answers "what kind of row is
// this?".  KDyn reads the kinds from the table row at run time (generic path: any scene the
// compiler can flatten).  KStatic<...> fixes them at compile time, so the switches fold away and
// a lens face, a lens edge, a stop, a sensor ... each become straight-line code; the kernels
// dispatch once per row on an opcode computed when the table is staged (classify_row).
struct KDyn {
    static RTT_HD int surf(const struct RowDev& R);
    static RTT_HD int bound(const struct RowDev& R);
    static RTT_HD int shape(const struct RowDev& R);
    static RTT_HD int phys(const struct RowDev& R);
    static RTT_HD int ident(const struct RowDev& R);
    static RTT_HD bool sensor(const struct RowDev& R);
    static RTT_HD bool specialised() { return false; }
};
template <int SURF, int BOUND, int SHAPE, int PHYS, int IDENT, int SENSOR = 0>
struct KStatic {
    static RTT_HD bool sensor(const struct RowDev&) { return SENSOR != 0; }
    static RTT_HD bool specialised() { return true; }
    static RTT_HD int surf(const struct RowDev&) { return SURF; }
    static RTT_HD int bound(const struct RowDev&) { return BOUND; }
    static RTT_HD int shape(const struct RowDev&) { return SHAPE; }
    static RTT_HD int phys(const struct RowDev&) { return PHYS; }
    static RTT_HD int ident(const struct RowDev&) { return IDENT; }
};

struct RowDev {
    float f[RTT_ROW_F];
    int32_t i[RTT_ROW_I];
};
RTT_HD int KDyn::surf(const RowDev& R) { return R.i[RTT_I_SURF]; }
RTT_HD int KDyn::bound(const RowDev& R) { return R.i[RTT_I_BOUND]; }
RTT_HD int KDyn::shape(const RowDev& R) { return R.i[RTT_I_SHAPE]; }
RTT_HD int KDyn::phys(const RowDev& R) { return R.i[RTT_I_PHYS]; }
RTT_HD int KDyn::ident(const RowDev& R) { return R.i[DI_IDENT]; }
RTT_HD bool KDyn::sensor(const RowDev& R) { return R.i[RTT_I_SENSOR] >= 0; }

struct V3 { float x, y, z; };

RTT_HD V3 v3(float x, float y, float z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
RTT_HD V3 operator+(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
RTT_HD V3 operator-(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
RTT_HD V3 operator*(float s, V3 a) { return v3(s * a.x, s * a.y, s * a.z); }
RTT_HD V3 operator-(V3 a) { return v3(-a.x, -a.y, -a.z); }
// torch.sum(a*b, dim=1): products rounded, accumulated left to right
RTT_HD float dot(V3 a, V3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
// p + t*d with t broadcast (geom/primitives.py:80-81)
RTT_HD V3 along(V3 p, float t, V3 d) { return v3(p.x + t * d.x, p.y + t * d.y, p.z + t * d.z); }

RTT_HD float rtt_inf() { return INFINITY; }
RTT_HD float rtt_nan() { return NAN; }
RTT_HD bool is_nan(float x) { return x != x; }

// Explicit fused multiply-add (honoured in EXACT mode too).  The reference's `[N,3] @ [3,3]`
// runs in the CPU BLAS as the chain fma(a2,R2j, fma(a1,R1j, a0*R0j)) and torch.norm over a
// 3-vector as fma(z,z, fma(y,y, x*x)) (measured on the reference host), so
// these two patterns are written as FMA chains in both variants.
RTT_HD float fma3(float a0, float b0, float a1, float b1, float a2, float b2) {
    return fmaf(a2, b2, fmaf(a1, b1, a0 * b0));
}
// a @ R   (row vector times matrix; geom/transform.py:92-93)
RTT_HD V3 mul_R(V3 a, const float* R) {
    return v3(fma3(a.x, R[0], a.y, R[3], a.z, R[6]),
              fma3(a.x, R[1], a.y, R[4], a.z, R[7]),
              fma3(a.x, R[2], a.y, R[5], a.z, R[8]));
}
// a @ R^T (geom/primitives.py:94, geom/shape.py:85)
RTT_HD V3 mul_RT(V3 a, const float* R) {
    return v3(fma3(a.x, R[0], a.y, R[1], a.z, R[2]),
              fma3(a.x, R[3], a.y, R[4], a.z, R[5]),
              fma3(a.x, R[6], a.y, R[7], a.z, R[8]));
}
// torch.norm / F.normalize over one 3-vector
RTT_HD float norm3(float x, float y, float z) { return sqrt_(fma3(x, x, y, y, z, z)); }
RTT_HD V3 ld3(const float* p) { return v3(p[0], p[1], p[2]); }
// v / s, component-wise (one reciprocal in FAST mode)
RTT_HD V3 div3(V3 v, float s) {
#if defined(RTT_APPROX) && defined(__CUDA_ARCH__)
    const float r = rcp_(s);
    return v3(v.x * r, v.y * r, v.z * r);
#else
    return v3(div_(v.x, s), div_(v.y, s), div_(v.z, s));
#endif
}

// Derived per-row constants; the same fp32 operations the reference performs on 0-dim tensors.
RTT_HD void prepare_row(RowDev& R) {
    const float c = R.f[RTT_F_C], k = R.f[RTT_F_K];
    R.f[D_C1K] = c * (1.0f + k);
    R.f[D_MU_ENTER] = R.f[RTT_F_IOR_OUT] / R.f[RTT_F_IOR_IN];
    R.f[D_MU_EXIT] = R.f[RTT_F_IOR_IN] / R.f[RTT_F_IOR_OUT];
    R.f[D_R2] = R.f[RTT_F_RADIUS] * R.f[RTT_F_RADIUS];
    R.f[D_SB0SQ] = R.f[RTT_F_SB] * R.f[RTT_F_SB];
    R.f[D_HB0SQ] = R.f[RTT_F_HB] * R.f[RTT_F_HB];
    if (R.i[RTT_I_SHAPE] == RTT_SHAPE_CYL_FACE || R.i[RTT_I_SHAPE] == RTT_SHAPE_CYL_EDGE) {
        // geom/cylindrics.py:31-37: x_min-1e-5 <= x <= x_max+1e-5 (same fp32 sums, done once per row)
        R.f[RTT_F_HB + 0] = R.f[RTT_F_HB + 0] - 1e-5f; R.f[RTT_F_HB + 1] = R.f[RTT_F_HB + 1] + 1e-5f;
        R.f[RTT_F_HB + 2] = R.f[RTT_F_HB + 2] - 1e-5f; R.f[RTT_F_HB + 3] = R.f[RTT_F_HB + 3] + 1e-5f;
    }
    int ident = 0;
    const float* Re = R.f + RTT_F_RE;
    const float* Rs = R.f + RTT_F_RS;
    bool ie = true, is = true;
    for (int a = 0; a < 9; ++a) {
        const float want = (a % 4 == 0) ? 1.0f : 0.0f;
        ie = ie && (Re[a] == want);
        is = is && (Rs[a] == want);
    }
    if (ie) ident |= 1;
    if (is) ident |= 2;
    R.i[DI_IDENT] = ident;
    int op = 0;
#define RTT_X(OP, SURF, BOUND, SHAPE, PHYS, IDENT, SENSOR)                                            \
    if (R.i[RTT_I_SURF] == SURF && R.i[RTT_I_BOUND] == BOUND && R.i[RTT_I_SHAPE] == SHAPE &&        \
        R.i[RTT_I_PHYS] == PHYS && ident == IDENT && (R.i[RTT_I_SENSOR] >= 0) == (SENSOR != 0)) op = OP;
    RTT_ROW_SPECS(RTT_X)
#undef RTT_X
#if defined(RTT_EXPERIMENT_GENERIC_ROWS)
    op = 0;                                                             // A/B build: every row on the generic path
#endif
    R.i[DI_OPCODE] = op;
}

// ---- poses ---------------------------------------------------------------------------------
RTT_HD V3 rot_fwd(V3 a, const float* R, bool ident) { return ident ? a : mul_R(a, R); }
RTT_HD V3 rot_bwd(V3 a, const float* R, bool ident) { return ident ? a : mul_RT(a, R); }

// F.normalize(v, p=2, dim=1, eps=1e-12)  (rays/ray.py:25)
RTT_HD V3 normalize12(V3 v, float* len_out) {
    const float s = norm3(v.x, v.y, v.z);
    const float den = fmaxf(s, 1e-12f);
    *len_out = s;
    // v/1 == v and 0/den == 0 bit for bit (signed zeros included): skip the IEEE division, whose
    // slow path these two operand patterns (unit directions, dead rays) would take every time
    if (den == 1.0f || (v.x == 0.0f && v.y == 0.0f && v.z == 0.0f)) return v;
    return div3(v, den);
}

// ---- surface-level bounds (geom/bounded.py) ------------------------------------------------
template <class K = KDyn>
RTT_HD bool surface_in_bounds(const RowDev& R, V3 h) {
    const float* sb = R.f + RTT_F_SB;
    switch (K::bound(R)) {
        case RTT_BOUND_DISK:                                            // :60-64
            return (h.x * h.x + h.y * h.y) <= R.f[D_SB0SQ];
        case RTT_BOUND_RECT:                                            // :77-82
            return (fabsf(h.x) <= sb[0]) & (fabsf(h.y) <= sb[1]);
        case RTT_BOUND_ELLIPSE: {                                       // :98-106
            const float u = h.x * sb[2] - h.y * sb[3];
            const float v = h.x * sb[3] + h.y * sb[2];
            const float a = div_(u, sb[0]), b = div_(v, sb[1]);
            return (a * a + b * b) <= 1.0f;
        }
        case RTT_BOUND_HALF:                                            // :123-127, :171-174
            return fabsf(h.z * R.f[RTT_F_C]) < 1.000001f;
        case RTT_BOUND_HALF_DISK:                                       // :151-159
            return (fabsf(h.z * R.f[RTT_F_C]) < 1.000001f) & ((h.x * h.x + h.y * h.y) <= R.f[D_SB0SQ]);
        case RTT_BOUND_NAPPE:                                           // :208-217 (slope in the c slot)
            return (h.z * R.f[RTT_F_C]) >= -1e-6f;
        default:
            return true;
    }
}

// ---- candidate roots (geom/primitives.py) --------------------------------------------------
struct Roots {
    float t1, t2;     // t2 unused for planes (n == 1)
    int n;
    // quantities the adjoint reuses
    float A, B, C, sq;
    bool lin;
};

template <class K = KDyn>
RTT_HD Roots solve_roots(const RowDev& R, V3 o, V3 d) {
    Roots q;
    q.n = 2; q.A = q.B = q.C = q.sq = 0.0f; q.lin = false;
    const float inf = rtt_inf();
    switch (K::surf(R)) {
        case RTT_SURF_PLANE: {                                          // :124-136
            const float safe = (fabsf(d.z) < 1e-6f) ? 1e-8f : d.z;
            q.t1 = div_(-o.z, safe); q.t2 = inf; q.n = 1; q.B = safe;
            return q;
        }
        case RTT_SURF_SPHERE: {                                         // :155-184 (a == 1 assumed)
            const float b = 2.0f * dot(o, d);
            const float cc = dot(o, o) - R.f[D_R2];
            const float disc = b * b - 4.0f * cc;
            const bool ok = disc >= 0.0f;
            const float sq = sqrt_(ok ? disc : 0.0f);
            q.t1 = ok ? (-b - sq) * 0.5f : inf;        // x/2 == x*0.5 exactly
            q.t2 = ok ? (-b + sq) * 0.5f : inf;
            q.B = b; q.sq = sq;
            return q;
        }
        case RTT_SURF_CYLINDER: {                                       // :201-231 (no A==0 guard)
            const float A = d.x * d.x + d.y * d.y;
            const float B = 2.0f * (o.x * d.x + o.y * d.y);
            const float Cq = (o.x * o.x + o.y * o.y) - R.f[D_R2];
            const float disc = B * B - (4.0f * A) * Cq;
            const bool ok = disc >= 0.0f;
            const float sq = sqrt_(fabsf(disc));
            const float den = 2.0f * A;
            if (den == 0.0f) {
                // axis-parallel ray (every collimated bundle): x / +0 = +-inf, 0 / 0 = NaN, spelled out so
                // the IEEE slow path is not taken; NaN => miss downstream exactly as in the reference
                const float n1 = -B - sq, n2 = -B + sq;
                q.t1 = ok ? (n1 > 0.0f ? inf : (n1 < 0.0f ? -inf : rtt_nan())) : inf;
                q.t2 = ok ? (n2 > 0.0f ? inf : (n2 < 0.0f ? -inf : rtt_nan())) : inf;
            } else {
                q.t1 = ok ? div_(-B - sq, den) : inf;
                q.t2 = ok ? div_(-B + sq, den) : inf;
            }
            q.A = A; q.B = B; q.C = Cq; q.sq = sq;
            return q;
        }
        case RTT_SURF_CONE: {                                           // :416-468, z^2 = k^2 (x^2 + y^2)
            const float k2 = R.f[RTT_F_C] * R.f[RTT_F_C];
            const float A = d.z * d.z - k2 * (d.x * d.x + d.y * d.y);
            const float B = 2.0f * (o.z * d.z - k2 * (o.x * d.x + o.y * d.y));
            const float Cq = o.z * o.z - k2 * (o.x * o.x + o.y * o.y);
            const float disc = B * B - (4.0f * A) * Cq;
            const bool ok = disc >= 0.0f;
            const bool lin = fabsf(A) < 1e-6f;
            const float sq = sqrt_(ok ? disc : 0.0f);
            const float den = 2.0f * (lin ? 1.0f : A);
            const float r1 = div_(-B - sq, den), r2 = div_(-B + sq, den);
            const float Bs = (fabsf(B) < 1e-6f) ? 1e-6f : B;
            const float tl = lin ? div_(-Cq, Bs) : 0.0f;
            q.t1 = lin ? tl : (ok ? r1 : inf);
            q.t2 = lin ? tl : (ok ? r2 : inf);
            q.A = A; q.B = B; q.C = Cq; q.sq = sq; q.lin = lin;
            return q;
        }
        default: {                                                      // conics :266-320, :356-376
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
            const float disc = B * B - (4.0f * A) * Cq;
            const bool ok = disc >= 0.0f;
            const bool lin = fabsf(A) < 1e-6f;
            const float sq = sqrt_(fabsf(disc));
            const float As = lin ? 1.0f : A;
            const float den = 2.0f * As;
            const float r1 = div_(-B - sq, den), r2 = div_(-B + sq, den);
            const float Bs = (fabsf(B) < 1e-6f) ? 1e-6f : B;
            const float tl = lin ? div_(-Cq, Bs) : 0.0f;
            q.t1 = lin ? tl : (ok ? r1 : inf);
            q.t2 = lin ? tl : (ok ? r2 : inf);
            q.A = A; q.B = B; q.C = Cq; q.sq = sq; q.lin = lin;
            return q;
        }
    }
}

// Smallest admissible root.  Unbounded: geom/primitives.py:28-36; bounded: geom/bounded.py:20-36.
// NaN propagates like torch.min (a NaN candidate that was not masked makes the result NaN).
// *which = index of the selected root (0/1).
template <class K = KDyn>
RTT_HD float select_root(const RowDev& R, const Roots& q, V3 o, V3 d, int* which) {
    const float inf = rtt_inf();
    float t1 = q.t1, t2 = q.t2;
    if (K::bound(R) == RTT_BOUND_NONE) {
        if (t1 <= 1e-6f) t1 = inf;
        if ((q.n == 2) & (t2 <= 1e-6f)) t2 = inf;
    } else {
        const bool inv = R.i[RTT_I_INVERT] != 0;
        bool k1 = surface_in_bounds<K>(R, along(o, t1, d));
        if (inv) k1 = !k1;
        if ((t1 <= 1e-6f) | !k1) t1 = inf;
        if (q.n == 2) {
            bool k2 = surface_in_bounds<K>(R, along(o, t2, d));
            if (inv) k2 = !k2;
            if ((t2 <= 1e-6f) | !k2) t2 = inf;
        }
    }
    if (q.n == 1) { *which = 0; return t1; }
    if (is_nan(t1) | is_nan(t2)) { *which = 0; return rtt_nan(); }
    if (t2 < t1) { *which = 1; return t2; }
    *which = 0;
    return t1;
}

// ---- shape-level validity (geom/shape.py:47-55) ---------------------------------------------
RTT_HD float sag_at(float c, float h, float tz) {                       // geom/bounded.py:129-139
    const float h2 = h * h;
    float term = 1.0f - (c * c) * h2;
    term = term > 0.0f ? term : 0.0f;
    return div_(c * h2, 1.0f + sqrt_(term)) + tz;
}

template <class K = KDyn>
RTT_HD bool shape_in_bounds(const RowDev* rows, int r, V3 h) {
    const RowDev& R = rows[r];
    const float* hb = R.f + RTT_F_HB;
    switch (K::shape(R)) {
        case RTT_SHAPE_SPHERIC_FACE:                                    // geom/spherics.py:40-46
            return (h.x * h.x + h.y * h.y) <= R.f[D_HB0SQ];
        case RTT_SHAPE_SPHERIC_EDGE:                                    // geom/spherics.py:34-39
            return (h.z >= hb[0]) & (h.z <= hb[1]);
        case RTT_SHAPE_CYL_FACE:
        case RTT_SHAPE_CYL_EDGE: {                                      // geom/cylindrics.py:23-55
            // `&`, not `&&`: four compares on one predicate chain instead of four short-circuit branches
            const bool ap = (h.x <= hb[1]) & (h.x >= hb[0]) & (h.y <= hb[3]) & (h.y >= hb[2]);       // slack pre-added
            if (K::shape(R) == RTT_SHAPE_CYL_FACE) return ap;
            if (!ap) return false;                                      // outside the aperture: the sag tests cannot save it
            const float zf = sag_at(hb[4], h.y, hb[5]);
            const float zb = sag_at(hb[6], h.y, hb[7]);
            return (h.z >= zf + 1e-4f) & (h.z <= zb - 1e-4f) & ap;
        }
        case RTT_SHAPE_POLY: {                                          // geom/shape.py:122-132
            const int first = R.i[RTT_I_POLY_FIRST], cnt = R.i[RTT_I_POLY_COUNT];
            if (R.f[D_SB0SQ] > 0.0f) {
                // a verified box (non-sequential kernel staging, rtt_tile.cuh::box_cull_info): a point that passes the
                // sibling half-space tests lies in the box inflated by 1e-4, hence inside its bounding sphere — a point
                // outside the sphere fails them, without evaluating the five planes
                const float ux = h.x - R.f[RTT_F_C], uy = h.y - R.f[RTT_F_K], uz = h.z - R.f[RTT_F_RADIUS];
                if ((ux * ux + uy * uy) + uz * uz > R.f[D_SB0SQ]) return false;
            }
            bool ok = true;
            for (int m = first; m < first + cnt; ++m) {
                if (m == r) continue;
                const float* Rm = rows[m].f + RTT_F_RS;                 // ROW 2 of R (reference quirk)
                const float* Tm = rows[m].f + RTT_F_TS;
                const float lz = (Rm[6] * (h.x - Tm[0]) + Rm[7] * (h.y - Tm[1])) + Rm[8] * (h.z - Tm[2]);
                ok = ok && (lz < 1e-4f);
                if (!ok) break;
            }
            return ok;
        }
        default:
            return true;
    }
}

// ---- intersection of one ray with one row ---------------------------------------------------
struct Frames {
    V3 pe, de;     // element frame (de NOT renormalised); == global for bare surfaces
    V3 den;        // renormalised element-frame direction
    float len;     // |de|
    V3 o, dd;      // surface frame
};

template <class K = KDyn>
RTT_HD Frames to_frames(const RowDev& R, V3 p, V3 d) {
    Frames F;
    const int ident = K::ident(R);
    if (K::shape(R) == RTT_SHAPE_NONE) {                           // geom/primitives.py:49
        F.pe = p; F.de = d; F.den = d; F.len = 1.0f;
    } else {                                                            // geom/shape.py:37-38
        F.pe = rot_fwd(p - ld3(R.f + RTT_F_TE), R.f + RTT_F_RE, ident & 1);
        F.de = rot_fwd(d, R.f + RTT_F_RE, ident & 1);
        F.den = normalize12(F.de, &F.len);
    }
    F.o = rot_fwd(F.pe - ld3(R.f + RTT_F_TS), R.f + RTT_F_RS, ident & 2);
    F.dd = rot_fwd(F.den, R.f + RTT_F_RS, ident & 2);
    return F;
}

// Distance with the surface-level rules only; returns true iff `t < inf` (false for NaN).
// reuse_elem: the element-frame part of F (pe, de, den, len) already holds this ray in THIS row's element frame
// (the previous row probed belongs to the same element: bit-identical pose, hence bit-identical values), so only
// the surface-frame part is recomputed — the nearest-hit search shares one element pose among a lens's or a box's rows.
template <class K = KDyn>
RTT_HD bool intersect_t(const RowDev* rows, int r, V3 p, V3 d, Frames& F, Roots& q, float& t, int& which,
                        bool reuse_elem = false) {
    const RowDev& R = rows[r];
    if (reuse_elem && K::shape(R) != RTT_SHAPE_NONE) {
        const int ident = K::ident(R);
        F.o = rot_fwd(F.pe - ld3(R.f + RTT_F_TS), R.f + RTT_F_RS, ident & 2);
        F.dd = rot_fwd(F.den, R.f + RTT_F_RS, ident & 2);
    } else {
        F = to_frames<K>(R, p, d);
    }
    q = solve_roots<K>(R, F.o, F.dd);
    t = select_root<K>(R, q, F.o, F.dd, &which);
    return t < rtt_inf();
}
// Shape-level rule of a hit at distance t (geom/shape.py:47-55); true for bare surfaces.
template <class K = KDyn>
RTT_HD bool shape_ok(const RowDev* rows, int r, const Frames& F, float t) {
    if (K::shape(rows[r]) == RTT_SHAPE_NONE) return true;
    return shape_in_bounds<K>(rows, r, along(F.pe, t, F.de));           // un-normalised de: shape.py:47
}
// Distance with every validity rule; returns true iff `t < inf` and valid.
// WITH_SHAPE=false is the Element.forward variant (geom/shape.py:61-87: no shape-level rule).
template <bool WITH_SHAPE, class K = KDyn>
RTT_HD bool intersect(const RowDev* rows, int r, V3 p, V3 d, Frames& F, Roots& q, float& t, int& which) {
    bool valid = intersect_t<K>(rows, r, p, d, F, q, t, which);
    if (WITH_SHAPE && valid) valid = shape_ok<K>(rows, r, F, t);
    return valid;
}

// ---- normals (geom/primitives.py:138-143, 186-187, 233-241, 330-343, 378-395) ---------------
template <class K = KDyn>
RTT_HD V3 normal_local(const RowDev& R, V3 h, float* len_out) {
    *len_out = 1.0f;
    switch (K::surf(R)) {
        case RTT_SURF_PLANE: return v3(0.0f, 0.0f, 1.0f);
        case RTT_SURF_SPHERE: return div3(h, R.f[RTT_F_RADIUS]);
        case RTT_SURF_CYLINDER: { const V3 q = div3(v3(h.x, h.y, 0.0f), R.f[RTT_F_RADIUS]); return v3(q.x, q.y, 0.0f); }
        case RTT_SURF_CONE: {                                           // :470-494, +z at the vertex / flat limit
            const float k2 = R.f[RTT_F_C] * R.f[RTT_F_C];
            const V3 raw = v3(-k2 * h.x, -k2 * h.y, h.z);
            const float len = norm3(raw.x, raw.y, raw.z);
            *len_out = len;
            if (!(len > 1e-8f)) return v3(0.0f, 0.0f, 1.0f);
            return div3(raw, len + 1e-8f);
        }
        default: {
            const float tc = 2.0f * R.f[RTT_F_C], tc1k = 2.0f * R.f[D_C1K];
            const float nx = (K::surf(R) == RTT_SURF_QUADRIC) ? tc * h.x : 0.0f;
            const float ny = tc * h.y;
            const float nz = tc1k * h.z - 2.0f;
#if defined(RTT_APPROX) && defined(__CUDA_ARCH__)
            // FAST: one rsqrt instead of sqrt + add + rcp (drops the reference's +1e-8 on a length of ~2: 5e-9 relative).
            // The squared length is clamped like the lean adjoint's (rtt_lean.cuh) instead of branching to the slow form
            // below: a conic's gradient (2cx, 2cy, 2c(1+k)z - 2) vanishes only at the centre of the quadric, never on it,
            // and the branch cost five predicated-off instructions at every hit.
            {
                const float l2 = fmaxf(fma3(nx, nx, ny, ny, nz, nz), 1e-30f);
                const float inv = rsqrt_(l2);
                *len_out = l2 * inv;
                return v3(-nx * inv, -ny * inv, -nz * inv);
            }
#else
            const float len = norm3(nx, ny, nz);
            const float den = len + 1e-8f;
            *len_out = len;
            return -div3(v3(nx, ny, nz), den);
#endif
        }
    }
}

template <class K = KDyn>
RTT_HD V3 normal_global(const RowDev& R, V3 nl) {
    const int ident = K::ident(R);
    // plane: nl == (0,0,1), so nl @ Rs^T is the third column of Rs (the FMA chain adds exact zeros)
    V3 n = (K::surf(R) == RTT_SURF_PLANE && !(ident & 2)) ? v3(R.f[RTT_F_RS + 2], R.f[RTT_F_RS + 5], R.f[RTT_F_RS + 8])
                                                           : rot_bwd(nl, R.f + RTT_F_RS, ident & 2);
    if (K::shape(R) != RTT_SHAPE_NONE) n = rot_bwd(n, R.f + RTT_F_RE, ident & 1);
    return n;
}

// ---- per-interaction extras of the stochastic Fresnel physics -------------------------------------
struct PhysAux {
    float ni, no;      // (ior_in, ior_out) this ray uses at the row (row values or the wavelength LUT's)
    float u;           // uniform [0, 1) draw of (ray, row, bounce)
};
RTT_HD PhysAux no_aux() { PhysAux a; a.ni = a.no = 1.0f; a.u = 0.0f; return a; }

// Fresnel decision (phys/std.py:177-199): true = reflect.  R = 1 under total internal reflection.
RTT_HD bool fresnel_reflects(float cos_i, float mu, float n1, float n2, float u, float* cos_t_out) {
    const float sin2_t = (mu * mu) * (1.0f - cos_i * cos_i);
    const float ct = sqrt_(fmaxf(1.0f - sin2_t, 0.0f));
    *cos_t_out = ct;
    if (sin2_t > 1.0f) return u < 1.0f;
    const float n1ci = n1 * cos_i, n2ct = n2 * ct, n1ct = n1 * ct, n2ci = n2 * cos_i;
    const float rs = div_(n1ci - n2ct, (n1ci + n2ct) + 1e-8f), rp = div_(n1ct - n2ci, (n1ct + n2ci) + 1e-8f);
    return u < 0.5f * (rs * rs + rp * rp);
}

// ---- physics (phys/std.py, phys/filter.py) --------------------------------------------------
// mu_enter = ior_out/ior_in, mu_exit = ior_in/ior_out (per wavelength when a LUT is present).
template <class K = KDyn>
RTT_HD V3 physics(const RowDev& R, V3 hl, V3 d, V3 n, float mu_enter, float mu_exit, float* mod,
                  PhysAux aux = no_aux()) {
    *mod = 1.0f;
    switch (K::phys(R)) {
        case RTT_PHYS_FRESNEL: {                                        // std.py:146-224
            const float dt = dot(d, n);
            const bool entering = dt < 0.0f;
            const float ci = fabsf(dt);
            const float n1 = entering ? aux.ni : aux.no, n2 = entering ? aux.no : aux.ni;
            const float mu = entering ? mu_enter : mu_exit;             // == n2 / n1
            float ct;
            if (fresnel_reflects(ci, mu, n1, n2, aux.u, &ct)) {
                const float tw = 2.0f * dt;
                return v3(d.x - tw * n.x, d.y - tw * n.y, d.z - tw * n.z);
            }
            const float qf = mu * ci - ct;
            const float qs = entering ? qf : -qf;
            return v3(mu * d.x + qs * n.x, mu * d.y + qs * n.y, mu * d.z + qs * n.z);
        }
        case RTT_PHYS_BLOCK:                                            // std.py:243-254
            *mod = 0.0f;
            return v3(0.0f, 0.0f, 0.0f);
        case RTT_PHYS_REFLECT: {                                        // std.py:97-108
            const float tw = 2.0f * dot(d, n);
            return v3(d.x - tw * n.x, d.y - tw * n.y, d.z - tw * n.z);
        }
        case RTT_PHYS_APERTURE: {                                       // filter.py:24-33
            const float m = surface_in_bounds<K>(R, hl) ? 1.0f : 0.0f;
            *mod = m;
            return v3(d.x * m, d.y * m, d.z * m);
        }
        case RTT_PHYS_SNELL: {                                          // std.py:123-145
            const float dt = dot(d, n);
            const bool entering = dt < 0.0f;
            const float c1 = fabsf(dt);
            const float mu = entering ? mu_enter : mu_exit;
            const float term = 1.0f - (mu * mu) * (1.0f - c1 * c1);
            if (term < 0.0f) {                                          // total internal reflection
                const float tw = 2.0f * dt;
                return v3(d.x - tw * n.x, d.y - tw * n.y, d.z - tw * n.z);
            }
            const float c2 = sqrt_(term > 0.0f ? term : 0.0f);
            const float qf = mu * c1 - c2;
            const float qs = entering ? qf : -qf;                       // qf * (-n) == (-qf) * n bit for bit
            return v3(mu * d.x + qs * n.x, mu * d.y + qs * n.y, mu * d.z + qs * n.z);
        }
        case RTT_PHYS_LINEAR: {                                         // std.py:72-88, Linear.transform = plane pose
            const V3 dl = mul_R(d, R.f + RTT_F_RS);                     // the reference multiplies even by an identity
            const float u = div_(dl.x, dl.z), v = div_(dl.y, dl.z);
            const float a = R.f[RTT_F_C] * hl.x + R.f[RTT_F_RADIUS] * u;
            const float b = R.f[RTT_F_K] * hl.y + R.f[RTT_F_IOR_IN] * v;
            float len;
            const V3 nl = normalize12(v3(a, b, 1.0f), &len);
            return mul_RT(nl, R.f + RTT_F_RS);
        }
        default:                                                        // Transmit std.py:227-235
            return d;
    }
}

// ---- one complete interaction (Element.forward, elements/parent.py:44-58) -------------------
struct Step {
    V3 hit_global, hit_local, normal, new_dir;
    float mod, t;
};

template <class K = KDyn>
RTT_HD Step interact(const RowDev& R, const Frames& F, float t, V3 p, V3 d, float mu_enter, float mu_exit,
                     PhysAux aux = no_aux()) {
    Step s;
    s.t = t;
    s.hit_local = along(F.o, t, F.dd);                                  // primitives.py:81
    float nlen;
    s.normal = normal_global<K>(R, normal_local<K>(R, s.hit_local, &nlen));
    s.hit_global = along(p, t, d);                                      // shape.py:81 / primitives.py:80
    s.new_dir = physics<K>(R, s.hit_local, d, s.normal, mu_enter, mu_exit, &s.mod, aux);
    return s;
}

// ---- sensor binning (restating gui/workbench.py:615-624 in fp32; see rtt_b200.h) ------------
RTT_HD bool sensor_bin(float x, float y, float x0, float y0, float sx, float sy, int W, int H, int* ix, int* iy) {
    const float fx = floorf((x - x0) * sx);
    const float fy = floorf((y - y0) * sy);
    if (!((fx >= 0.0f) & (fx < (float)W) & (fy >= 0.0f) & (fy < (float)H))) return false;
    *ix = (int)fx; *iy = (int)fy;
    return true;
}

// =============================================================================================
// Ray sources (rays/bundle.py:30-171, render/camera.py:39-72): rays generated from a counter-based
// RNG instead of being read from memory.  Integer part (Philox4x32-10) is bit-identical on host
// and device; the float part uses the accurate libm calls (sincosf / acosf / sqrtf).
// =============================================================================================
// The source record is a template parameter: rtt_source_t itself (host checker) or the kernels' by-value
// copy of it (rtt_kernels_decl.h SourceDev); both have the fields kind, a[4], width, height, pose, seed,
// first, state.
RTT_HD void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t out[4]) {
    for (int round = 0; round < 10; ++round) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        const uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        const uint32_t n1 = (uint32_t)p1;
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        const uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
// [0, 1) with 24 random bits, like torch.rand in fp32
RTT_HD float u01(uint32_t x) { return (float)(x >> 8) * (1.0f / 16777216.0f); }

// Uniform draw of the Fresnel decision of (ray i, row r, bounce b) under the table's seed (rtt_b200.h).
RTT_HD unsigned long long table_seed(const RowDev* rows) {
    return ((unsigned long long)(uint32_t)rows[0].i[RTT_I_RNG_HI] << 32) | (uint32_t)rows[0].i[RTT_I_RNG_LO];
}
RTT_HD float fresnel_u(unsigned long long seed, long long i, int r, int b) {
    uint32_t out[4];
    philox4x32_10((uint32_t)i, (uint32_t)((unsigned long long)i >> 32), (uint32_t)(r + 256 * b), 0x4672u,
                  (uint32_t)seed, (uint32_t)(seed >> 32), out);
    return u01(out[0]);
}
// Extras of one interaction; the draw is only made for Fresnel rows.
template <class K = KDyn>
RTT_HD PhysAux make_aux(const RowDev* rows, const RowDev& R, float ni, float no, long long i, int r, int b) {
    PhysAux a;
    a.ni = ni; a.no = no; a.u = 0.0f;
    if (K::phys(R) == RTT_PHYS_FRESNEL) a.u = fresnel_u(table_seed(rows), i, r, b);
    return a;
}
// torch.linspace(start, end, steps)[i]: symmetric evaluation from both ends
RTT_HD float linspace_at(float start, float end, int steps, int i) {
    if (steps <= 1) return start;
    const float step = (end - start) / (float)(steps - 1);
    return (i < steps / 2) ? fmaf(step, (float)i, start) : fmaf(-step, (float)(steps - 1 - i), end);
}

struct SourceKey { unsigned long long key, base; };
template <class SRC>
RTT_HD SourceKey source_key(const SRC& s) {
    SourceKey k;
    k.key = s.state ? s.state[0] : s.seed;
    k.base = (unsigned long long)s.first + (s.state ? s.state[1] : 0ull);
    return k;
}

// Ray i of the source, in the global frame, direction normalised (Rays.__post_init__, rays/ray.py:25).
template <class SRC>
RTT_HD void source_ray(const SRC& s, SourceKey k, long long i, V3& p, V3& d) {
    const unsigned long long ctr = k.base + (unsigned long long)i;
    uint32_t r[4];
    philox4x32_10((uint32_t)ctr, (uint32_t)(ctr >> 32), 0u, 0u, (uint32_t)k.key, (uint32_t)(k.key >> 32), r);
    const float u0 = u01(r[0]), u1 = u01(r[1]);
    const float* R = s.pose;
    const V3 T = ld3(s.pose + 9);
    V3 pl = v3(0.0f, 0.0f, 0.0f), dl = v3(0.0f, 0.0f, 1.0f);
    float len;
    switch (s.kind) {
        case RTT_SRC_DISK: {                                            // rays/bundle.py:40-56, 83-100
            const float th = fmaf(u0, s.a[3] - s.a[2], s.a[2]);
            const float rr = sqrtf(fmaf(u1, s.a[1] - s.a[0], s.a[0]));
            float sn, cs;
            sincosf(th, &sn, &cs);
            pl = v3(rr * cs, rr * sn, 0.0f);
            break;
        }
        case RTT_SRC_LINE:                                              // rays/bundle.py:103-118
            pl = v3(fmaf(u0, 2.0f * s.a[0], -s.a[0]), 0.0f, 0.0f);
            break;
        case RTT_SRC_FAN: {                                             // rays/bundle.py:121-140
            const float th = fmaf(u0, 2.0f * s.a[0], -s.a[0]);
            float sn, cs;
            sincosf(th, &sn, &cs);
            dl = v3(0.0f, sn, cs);
            break;
        }
        case RTT_SRC_POINT: {                                           // rays/bundle.py:58-80, 143-170
            const float ph = acosf(fmaf(-2.0f, fmaf(u0, s.a[1] - s.a[0], s.a[0]), 1.0f));
            const float th = fmaf(u1, s.a[3] - s.a[2], s.a[2]);
            float sp, cp, st, ct;
            sincosf(ph, &sp, &cp);
            sincosf(th, &st, &ct);
            dl = v3(ct * sp, st * sp, cp);
            break;
        }
        default: {                                                      // RTT_SRC_CAMERA: render/camera.py:39-72
            const long long npix = (long long)s.width * s.height;
            const long long g = (long long)ctr;
            const long long pix = g % npix, smp = g / npix;
            const int px = (int)(pix % s.width), py = (int)(pix / s.width);
            float x = linspace_at(-s.a[0], s.a[0], s.width, px);
            float y = linspace_at(s.a[1], -s.a[1], s.height, py);
            if (smp > 0) {                                              // extension: jittered sub-pixel samples
                x = fmaf(u0 - 0.5f, 2.0f * s.a[0] / (float)(s.width > 1 ? s.width - 1 : 1), x);
                y = fmaf(u1 - 0.5f, 2.0f * s.a[1] / (float)(s.height > 1 ? s.height - 1 : 1), y);
            }
            const V3 dg = v3(fmaf(y, R[3], fmaf(x, R[0], R[6])), fmaf(y, R[4], fmaf(x, R[1], R[7])), fmaf(y, R[5], fmaf(x, R[2], R[8])));
            p = T;
            d = normalize12(dg, &len);
            return;
        }
    }
    p = mul_RT(pl, R) + T;                                              // geom/transform.py:262-268
    d = normalize12(mul_RT(dl, R), &len);
}

// =============================================================================================
// Adjoint
// =============================================================================================
// Gradient of one row's differentiable entries, in table_f order [0, RTT_N_DIFF).
struct RowGrad {
    float g[RTT_N_DIFF];
};

RTT_HD void zero(RowGrad& G) {
    for (int a = 0; a < RTT_N_DIFF; ++a) G.g[a] = 0.0f;
}

// y = a @ R : ga += gy @ R^T ; gR[i][j] += a[i]*gy[j]
RTT_HD V3 adj_mul_R(V3 a, V3 gy, const float* R, bool ident, float* gR, bool want_gR) {
    if (want_gR) {
        gR[0] += a.x * gy.x; gR[1] += a.x * gy.y; gR[2] += a.x * gy.z;
        gR[3] += a.y * gy.x; gR[4] += a.y * gy.y; gR[5] += a.y * gy.z;
        gR[6] += a.z * gy.x; gR[7] += a.z * gy.y; gR[8] += a.z * gy.z;
    }
    return ident ? gy : mul_RT(gy, R);
}
// y = a @ R^T : ga += gy @ R ; gR[j][i] += gy[j]*a[i]
RTT_HD V3 adj_mul_RT(V3 a, V3 gy, const float* R, bool ident, float* gR, bool want_gR) {
    if (want_gR) {
        gR[0] += gy.x * a.x; gR[1] += gy.x * a.y; gR[2] += gy.x * a.z;
        gR[3] += gy.y * a.x; gR[4] += gy.y * a.y; gR[5] += gy.y * a.z;
        gR[6] += gy.z * a.x; gR[7] += gy.z * a.y; gR[8] += gy.z * a.z;
    }
    return ident ? gy : mul_R(gy, R);
}

// Reverse of `interact` for a ray known to have hit row R with incoming state (p, d).
// Upstream: g_pos (d/d new_pos), g_dir (d/d new_dir), g_hl_up (d/d hit_local, sensor record),
// g_n_up (d/d normal) and g_t_up (d/d t) for callers of the single-surface op.
// Produces d/d p, d/d d and accumulates the row's parameter gradients into G.
// Intensity is handled by the caller (g_I_in = g_I_out * mod; mod is returned).
// (ni, no) are the (ior_in, ior_out) this ray used (row values, or the wavelength LUT's).
template <class K = KDyn>
RTT_HD void interact_adjoint(const RowDev& R, V3 p, V3 d, float ni, float no, float mu_enter, float mu_exit,
                             V3 g_pos, V3 g_dir, V3 g_hl_up, V3 g_n_up, float g_t_up,
                             V3& g_p, V3& g_d, float& mod_out, RowGrad& G, int want, float fresnel_u = 0.0f) {
    // ---- recompute the forward pieces (same selections as the forward pass) ----
    const Frames F = to_frames<K>(R, p, d);
    const Roots q = solve_roots<K>(R, F.o, F.dd);
    int which;
    const float t = select_root<K>(R, q, F.o, F.dd, &which);
    const V3 hl = along(F.o, t, F.dd);
    float nlen;
    const V3 nl = normal_local<K>(R, hl, &nlen);
    const int ident = K::ident(R);
    const bool has_shape = K::shape(R) != RTT_SHAPE_NONE;
    const V3 n_e = rot_bwd(nl, R.f + RTT_F_RS, ident & 2);              // element-frame normal
    const V3 n = has_shape ? rot_bwd(n_e, R.f + RTT_F_RE, ident & 1) : n_e;
    const bool w_pose_e = (want & RTT_FLAG_GRAD_POSE_E) != 0;
    const bool w_pose_s = (want & RTT_FLAG_GRAD_POSE_S) != 0;

    g_p = g_pos;                                                        // hit_global = p + t d
    g_d = v3(t * g_pos.x, t * g_pos.y, t * g_pos.z);
    float g_t = dot(g_pos, d) + g_t_up;
    V3 g_n = v3(0.0f, 0.0f, 0.0f);
    V3 g_hl = g_hl_up;
    float mod = 1.0f;

    // ---- physics ----
    int phys_eff = K::phys(R);
    if (K::phys(R) == RTT_PHYS_FRESNEL) {
        // the branch is a non-differentiable choice (std.py:190-193): differentiate the one the forward pass took
        const float dt = dot(d, n);
        const bool entering = dt < 0.0f;
        float ct;
        const bool refl = fresnel_reflects(fabsf(dt), entering ? mu_enter : mu_exit, entering ? ni : no,
                                           entering ? no : ni, fresnel_u, &ct);
        phys_eff = refl ? RTT_PHYS_REFLECT : RTT_PHYS_SNELL;
    }
    switch (phys_eff) {
        case RTT_PHYS_BLOCK:
            mod = 0.0f;
            break;
        case RTT_PHYS_APERTURE: {
            const float m = surface_in_bounds<K>(R, hl) ? 1.0f : 0.0f;
            mod = m;
            g_d = g_d + m * g_dir;
            break;
        }
        case RTT_PHYS_REFLECT: {
            const float dn = dot(d, n), gn = dot(g_dir, n);
            g_d = g_d + (g_dir - (2.0f * gn) * n);
            g_n = -2.0f * (gn * d + dn * g_dir);
            break;
        }
        case RTT_PHYS_SNELL: {
            const float dt = dot(d, n);
            const bool entering = dt < 0.0f;
            const float c1 = fabsf(dt);
            const float mu = entering ? mu_enter : mu_exit;
            const float one_m = 1.0f - c1 * c1;
            const float term = 1.0f - (mu * mu) * one_m;
            if (term < 0.0f) {
                const float gn = dot(g_dir, n);
                g_d = g_d + (g_dir - (2.0f * gn) * n);
                g_n = -2.0f * (gn * d + dt * g_dir);
            } else {
                const float c2 = sqrt_(term > 0.0f ? term : 0.0f);
                const float sgn = entering ? 1.0f : -1.0f;
                const float qf = mu * c1 - c2;
                float g_mu = dot(g_dir, d);
                const float g_q = sgn * dot(g_dir, n);
                g_d = g_d + mu * g_dir;
                g_n = (qf * sgn) * g_dir;
                g_mu += c1 * g_q;
                float g_c1 = mu * g_q;
                const float g_term = (term > 0.0f) ? div_(-g_q, 2.0f * c2) : 0.0f;
                g_mu += g_term * (-2.0f * mu * one_m);
                g_c1 += g_term * (2.0f * mu * mu * c1);
                const float g_dt = (dt < 0.0f) ? -g_c1 : ((dt > 0.0f) ? g_c1 : 0.0f);
                g_d = g_d + g_dt * n;
                g_n = g_n + g_dt * d;
                if (want & RTT_FLAG_GRAD_IOR) {
                    if (entering) {                                     // mu = no/ni
                        G.g[RTT_F_IOR_OUT] += div_(g_mu, ni);
                        G.g[RTT_F_IOR_IN] -= div_(g_mu * no, ni * ni);
                    } else {                                            // mu = ni/no
                        G.g[RTT_F_IOR_IN] += div_(g_mu, no);
                        G.g[RTT_F_IOR_OUT] -= div_(g_mu * ni, no * no);
                    }
                }
            }
            break;
        }
        case RTT_PHYS_LINEAR: {
            // out = normalize(Cx hl.x + Dx u, Cy hl.y + Dy v, 1) @ Rs^T with (u, v) = (dl.x, dl.y) / dl.z, dl = d @ Rs
            const float* Rs = R.f + RTT_F_RS;
            const V3 dl = mul_R(d, Rs);
            const float iz = rcp_(dl.z);
            const float u = dl.x * iz, v = dl.y * iz;
            const float Cx = R.f[RTT_F_C], Cy = R.f[RTT_F_K], Dx = R.f[RTT_F_RADIUS], Dy = R.f[RTT_F_IOR_IN];
            const float a = Cx * hl.x + Dx * u, b = Cy * hl.y + Dy * v;
            const float inv = rcp_(norm3(a, b, 1.0f));
            const V3 nl2 = v3(a * inv, b * inv, inv);
            const V3 g_nl2 = adj_mul_RT(nl2, g_dir, Rs, false, G.g + RTT_F_RS, w_pose_s);
            const float pr = dot(nl2, g_nl2);
            const float g_a = (g_nl2.x - nl2.x * pr) * inv, g_b = (g_nl2.y - nl2.y * pr) * inv;
            if (want & RTT_FLAG_GRAD_CK) { G.g[RTT_F_C] += g_a * hl.x; G.g[RTT_F_K] += g_b * hl.y; }
            if (want & RTT_FLAG_GRAD_RADIUS) G.g[RTT_F_RADIUS] += g_a * u;
            if (want & RTT_FLAG_GRAD_IOR) G.g[RTT_F_IOR_IN] += g_b * v;
            g_hl.x += g_a * Cx; g_hl.y += g_b * Cy;
            const float g_u = g_a * Dx, g_v = g_b * Dy;
            const V3 g_dl = v3(g_u * iz, g_v * iz, -(g_u * u + g_v * v) * iz);
            g_d = g_d + adj_mul_R(d, g_dl, Rs, false, G.g + RTT_F_RS, w_pose_s);
            break;
        }
        default:
            g_d = g_d + g_dir;
            break;
    }
    mod_out = mod;
    g_n = g_n + g_n_up;

    // ---- normal: n = (nl @ Rs^T) @ Re^T ----
    V3 g_ne = g_n;
    if (has_shape) g_ne = adj_mul_RT(n_e, g_n, R.f + RTT_F_RE, ident & 1, G.g + RTT_F_RE, w_pose_e);
    const V3 g_nl = adj_mul_RT(nl, g_ne, R.f + RTT_F_RS, ident & 2, G.g + RTT_F_RS, w_pose_s);

    // ---- normal_local(hl) ----
    switch (K::surf(R)) {
        case RTT_SURF_PLANE: break;
        case RTT_SURF_SPHERE: {
            const float rr = R.f[RTT_F_RADIUS];
            g_hl = g_hl + div3(g_nl, rr);
            G.g[RTT_F_RADIUS] -= div_(dot(g_nl, hl), rr * rr);
            break;
        }
        case RTT_SURF_CYLINDER: {
            const float rr = R.f[RTT_F_RADIUS];
            { const float ir = rcp_(rr); g_hl.x += g_nl.x * ir; g_hl.y += g_nl.y * ir; }
            G.g[RTT_F_RADIUS] -= div_(g_nl.x * hl.x + g_nl.y * hl.y, rr * rr);
            break;
        }
        case RTT_SURF_CONE: {
            if (nlen > 1e-8f) {                                         // nl = raw / (|raw| + 1e-8), raw = (-k2 x, -k2 y, z)
                const float sl = R.f[RTT_F_C], k2 = sl * sl;
                const V3 raw = v3(-k2 * hl.x, -k2 * hl.y, hl.z);
                const float den = nlen + 1e-8f;
                const float s = div_(dot(raw, g_nl), den * den * nlen);
                const V3 gq = div3(g_nl, den);
                const V3 g_raw = v3(gq.x - raw.x * s, gq.y - raw.y * s, gq.z - raw.z * s);
                g_hl.x += -k2 * g_raw.x; g_hl.y += -k2 * g_raw.y; g_hl.z += g_raw.z;
                G.g[RTT_F_C] += -(hl.x * g_raw.x + hl.y * g_raw.y) * (2.0f * sl);
            }
            break;
        }
        default: {
            const float c = R.f[RTT_F_C], k = R.f[RTT_F_K];
            const float tc = 2.0f * c, tc1k = 2.0f * R.f[D_C1K];
            const bool full = K::surf(R) == RTT_SURF_QUADRIC;
            const V3 raw = v3(full ? tc * hl.x : 0.0f, tc * hl.y, tc1k * hl.z - 2.0f);
            const float den = nlen + 1e-8f;
            // nl = -raw/den, den = |raw| + 1e-8
            const float rg = dot(raw, g_nl);
            const float s = (nlen > 0.0f) ? div_(rg, den * den * nlen) : 0.0f;
            const V3 gq = div3(g_nl, den);
            const V3 g_raw = v3(-gq.x + raw.x * s, -gq.y + raw.y * s, -gq.z + raw.z * s);
            float g_tc = g_raw.y * hl.y;
            if (full) { g_hl.x += tc * g_raw.x; g_tc += g_raw.x * hl.x; }
            g_hl.y += tc * g_raw.y;
            g_hl.z += tc1k * g_raw.z;
            const float g_tc1k = g_raw.z * hl.z;
            G.g[RTT_F_C] += 2.0f * g_tc + 2.0f * (1.0f + k) * g_tc1k;
            G.g[RTT_F_K] += 2.0f * c * g_tc1k;
            break;
        }
    }

    // ---- hit_local = o + t dd ----
    V3 g_o = g_hl;
    V3 g_dd = t * g_hl;
    g_t += dot(g_hl, F.dd);

    // ---- t = selected root ----
    const V3 o = F.o, dd = F.dd;
    switch (K::surf(R)) {
        case RTT_SURF_PLANE: {
            const float safe = q.B;
            const float g_t_s = div_(g_t, safe);
            g_o.z += -g_t_s;
            if (!(fabsf(dd.z) < 1e-6f)) g_dd.z += -g_t_s * t;      // d(-oz/dz)/d dz = -t/dz
            break;
        }
        case RTT_SURF_SPHERE: {
            const float sg = which ? 1.0f : -1.0f;
            const float b = q.B, sq = q.sq;
            const float isq = rcp_(sq);
            const float g_b = g_t * (-1.0f + sg * b * isq) * 0.5f;
            const float g_cc = -g_t * sg * isq;
            g_o = g_o + (2.0f * g_b) * dd + (2.0f * g_cc) * o;
            g_dd = g_dd + (2.0f * g_b) * o;
            G.g[RTT_F_RADIUS] += g_cc * (-2.0f * R.f[RTT_F_RADIUS]);
            break;
        }
        case RTT_SURF_CYLINDER: {
            const float gC = -g_t * rcp_(2.0f * q.A * t + q.B);
            const float gB = gC * t, gA = gB * t;
            g_dd.x += gA * 2.0f * dd.x + gB * 2.0f * o.x;
            g_dd.y += gA * 2.0f * dd.y + gB * 2.0f * o.y;
            g_o.x += gB * 2.0f * dd.x + gC * 2.0f * o.x;
            g_o.y += gB * 2.0f * dd.y + gC * 2.0f * o.y;
            G.g[RTT_F_RADIUS] += gC * (-2.0f * R.f[RTT_F_RADIUS]);
            break;
        }
        case RTT_SURF_CONE: {
            const float sl = R.f[RTT_F_C], k2 = sl * sl;
            float gA, gB, gC;
            if (q.lin) {
                const float Bs = (fabsf(q.B) < 1e-6f) ? 1e-6f : q.B;
                gA = 0.0f;
                gC = div_(-g_t, Bs);
                gB = (fabsf(q.B) < 1e-6f) ? 0.0f : div_(g_t * q.C, Bs * Bs);
            } else {
                gC = -g_t * rcp_(2.0f * q.A * t + q.B);
                gB = gC * t; gA = gB * t;
            }
            // A = dz^2 - k2 (dx^2+dy^2), B = 2 (oz dz - k2 (ox dx + oy dy)), C = oz^2 - k2 (ox^2+oy^2)
            g_dd.x += -2.0f * k2 * (gA * dd.x + gB * o.x);
            g_dd.y += -2.0f * k2 * (gA * dd.y + gB * o.y);
            g_dd.z += 2.0f * (gA * dd.z + gB * o.z);
            g_o.x += -2.0f * k2 * (gB * dd.x + gC * o.x);
            g_o.y += -2.0f * k2 * (gB * dd.y + gC * o.y);
            g_o.z += 2.0f * (gB * dd.z + gC * o.z);
            const float g_k2 = -(gA * (dd.x * dd.x + dd.y * dd.y) + 2.0f * gB * (o.x * dd.x + o.y * dd.y) +
                                 gC * (o.x * o.x + o.y * o.y));
            G.g[RTT_F_C] += g_k2 * (2.0f * sl);
            break;
        }
        default: {
            const float c = R.f[RTT_F_C], k = R.f[RTT_F_K], c1k = R.f[D_C1K];
            float gA, gB, gC;
            if (q.lin) {
                const float Bs = (fabsf(q.B) < 1e-6f) ? 1e-6f : q.B;
                gA = 0.0f;
                gC = div_(-g_t, Bs);
                gB = (fabsf(q.B) < 1e-6f) ? 0.0f : div_(g_t * q.C, Bs * Bs);
            } else {
                gC = -g_t * rcp_(2.0f * q.A * t + q.B);
                gB = gC * t; gA = gB * t;
            }
            const bool full = K::surf(R) == RTT_SURF_QUADRIC;
            const float ox = full ? o.x : 0.0f, dx = full ? dd.x : 0.0f;
            const float g_u = gA * (dx * dx + dd.y * dd.y) + gB * 2.0f * (ox * dx + o.y * dd.y) + gC * (ox * ox + o.y * o.y);
            const float g_v = gA * (dd.z * dd.z) + gB * 2.0f * (o.z * dd.z) + gC * (o.z * o.z);
            if (full) {
                g_dd.x += gA * 2.0f * c * dx + gB * 2.0f * c * ox;
                g_o.x += gB * 2.0f * c * dx + gC * 2.0f * c * ox;
            }
            g_dd.y += gA * 2.0f * c * dd.y + gB * 2.0f * c * o.y;
            g_o.y += gB * 2.0f * c * dd.y + gC * 2.0f * c * o.y;
            g_dd.z += gA * 2.0f * c1k * dd.z + gB * (2.0f * c1k * o.z - 2.0f);
            g_o.z += gB * 2.0f * c1k * dd.z + gC * (2.0f * c1k * o.z - 2.0f);
            G.g[RTT_F_C] += g_u + (1.0f + k) * g_v;
            G.g[RTT_F_K] += c * g_v;
            break;
        }
    }

    // ---- surface pose: o = (pe - Ts) @ Rs ; dd = den @ Rs ----
    const V3 pes = F.pe - ld3(R.f + RTT_F_TS);
    const V3 g_pes = adj_mul_R(pes, g_o, R.f + RTT_F_RS, ident & 2, G.g + RTT_F_RS, w_pose_s);
    const V3 g_den = adj_mul_R(F.den, g_dd, R.f + RTT_F_RS, ident & 2, G.g + RTT_F_RS, w_pose_s);
    if (w_pose_s) { G.g[RTT_F_TS] -= g_pes.x; G.g[RTT_F_TS + 1] -= g_pes.y; G.g[RTT_F_TS + 2] -= g_pes.z; }

    if (!has_shape) {
        g_p = g_p + g_pes;
        g_d = g_d + g_den;
        return;
    }
    // ---- renormalisation: den = de / max(|de|, 1e-12) ----
    V3 g_de;
    if (F.len > 1e-12f) {
        const float pr = dot(F.den, g_den);
        g_de = div3(v3(g_den.x - F.den.x * pr, g_den.y - F.den.y * pr, g_den.z - F.den.z * pr), F.len);
    } else {
        g_de = v3(g_den.x * 1e12f, g_den.y * 1e12f, g_den.z * 1e12f);
    }
    // ---- element pose: pe = (p - Te) @ Re ; de = d @ Re ----
    const V3 pte = p - ld3(R.f + RTT_F_TE);
    const V3 g_pte = adj_mul_R(pte, g_pes, R.f + RTT_F_RE, ident & 1, G.g + RTT_F_RE, w_pose_e);
    const V3 g_dg = adj_mul_R(d, g_de, R.f + RTT_F_RE, ident & 1, G.g + RTT_F_RE, w_pose_e);
    if (w_pose_e) { G.g[RTT_F_TE] -= g_pte.x; G.g[RTT_F_TE + 1] -= g_pte.y; G.g[RTT_F_TE + 2] -= g_pte.z; }
    g_p = g_p + g_pte;
    g_d = g_d + g_dg;
}

}  // namespace rtt
