/*
 * rtt_b200 — C ABI of the B200 (sm_100a) ray-propagation kernels.
 *
 * This is the drop-in boundary for the batched ray-trace hot path of RayTraceTorch.
 * The reference is pure Python/PyTorch and has no FFI of its own; each entry point below
 * replaces one Python seam of the reference (file:line under /root/reference) and is what
 * a binding written by the reference's maintainer (ctypes) calls.
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless it says
 *     "host"; all float arrays are contiguous fp32, int arrays int32, masks uint64/uint8
 *   - the caller owns every buffer; the library allocates nothing that outlives a call
 *     and keeps no state that influences a result: no caches, no environment variables, the
 *     surface table is read from device memory and staged in shared memory by every thread
 *     block, so calls are re-entrant (the one process-wide variable is the diagnostic launch
 *     counter behind rtt_launch_count())
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, nothing
 *     synchronises
 *   - return value: 0 on success, a cudaError_t (>0) from the launch, or a negative
 *     RTT_E_* code for argument errors; rtt_error_string() explains any of them
 *   - there is NO CPU path: without a CUDA device every compute entry returns an error
 *
 * Surface table (built by raytracetorch_b200/table.py from the live nn.Parameters):
 *   table_f [n_rows, RTT_ROW_F] fp32, table_i [n_rows, RTT_ROW_I] int32; one row per
 *   (element, surface index) in the reference's flattening order
 *   (scene/base.py:116-123, scene/sequential.py:17-19).  Layout: see the enums below and
 *   raytracetorch_b200/codes.py (rtt_layout_query() lets a binding verify both agree).
 */
#ifndef RTT_B200_H
#define RTT_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- row layout (floats) ---- */
enum {
    RTT_ROW_F = 48, RTT_F_RE = 0, RTT_F_TE = 9, RTT_F_RS = 12, RTT_F_TS = 21,
    RTT_F_C = 24, RTT_F_K = 25, RTT_F_RADIUS = 26, RTT_F_IOR_IN = 27, RTT_F_IOR_OUT = 28,
    RTT_N_DIFF = 29, RTT_F_SB = 29, RTT_F_HB = 33, RTT_ROW_G = 48
};
/* ---- row layout (ints) ---- */
enum {
    RTT_ROW_I = 16, RTT_I_SURF = 0, RTT_I_BOUND = 1, RTT_I_INVERT = 2, RTT_I_SHAPE = 3,
    RTT_I_PHYS = 4, RTT_I_SENSOR = 5, RTT_I_POLY_FIRST = 6, RTT_I_POLY_COUNT = 7,
    RTT_I_ELEM = 8, RTT_I_SIDX = 9, RTT_I_FLAGS = 10
};
enum { RTT_SURF_PLANE = 0, RTT_SURF_QUADRIC, RTT_SURF_QUADRIC_ZY, RTT_SURF_CYLINDER, RTT_SURF_SPHERE,
       RTT_SURF_CONE };      /* double cone z^2 = slope^2 (x^2 + y^2), slope in f[RTT_F_C] (geom/primitives.py:398-494) */
enum { RTT_BOUND_NONE = 0, RTT_BOUND_DISK, RTT_BOUND_RECT, RTT_BOUND_ELLIPSE, RTT_BOUND_HALF, RTT_BOUND_HALF_DISK,
       RTT_BOUND_NAPPE };    /* one nappe of a cone: z * slope >= -1e-6 (geom/bounded.py:189-217) */
enum { RTT_SHAPE_NONE = 0, RTT_SHAPE_SPHERIC_FACE, RTT_SHAPE_SPHERIC_EDGE, RTT_SHAPE_CYL_FACE,
       RTT_SHAPE_CYL_EDGE, RTT_SHAPE_POLY, RTT_SHAPE_OPEN };
enum { RTT_PHYS_TRANSMIT = 0, RTT_PHYS_SNELL, RTT_PHYS_REFLECT, RTT_PHYS_BLOCK, RTT_PHYS_APERTURE,
       RTT_PHYS_LINEAR, RTT_PHYS_FRESNEL };
/* RTT_PHYS_FRESNEL (phys/std.py:146-224): reflect with probability R (unpolarised Fresnel reflectance, 1 under total
 * internal reflection), else refract.  The reference draws torch.rand_like; here the uniform number of (ray i, row r,
 * bounce b) is Philox4x32-10(counter = {i, r + 256 b, "Fr"}, key = the 64-bit seed stored in ints
 * [RTT_I_RNG_LO, RTT_I_RNG_HI] of table row 0), so a trace and its adjoint take the same branch and the parity
 * with the reference is statistical.  i counts from the start of the launch (plus the source's `first`). */
enum { RTT_I_RNG_LO = 14, RTT_I_RNG_HI = 15 };
/* RTT_PHYS_LINEAR (phys/std.py:35-88, the ideal thin lens / mirror elements of elements/ideal.py): a plane row
 * whose otherwise unused scalar slots carry the ray-transfer coefficients — f[RTT_F_C] = Cx, f[RTT_F_K] = Cy,
 * f[RTT_F_RADIUS] = Dx, f[RTT_F_IOR_IN] = Dy — with the matching gradient flags (CK, RADIUS, IOR). */
enum { RTT_FLAG_GRAD_POSE_E = 1, RTT_FLAG_GRAD_POSE_S = 2, RTT_FLAG_GRAD_CK = 4,
       RTT_FLAG_GRAD_RADIUS = 8, RTT_FLAG_GRAD_IOR = 16 };
enum { RTT_MAX_ROWS = 64, RTT_MAX_SENSORS = 4, RTT_MAX_WAVELENGTHS = 8, RTT_MAX_BOUNCES = 255 };

/* ---- error codes ---- */
enum { RTT_OK = 0, RTT_E_ARG = -1, RTT_E_ROWS = -2, RTT_E_NO_DEVICE = -3, RTT_E_ALIGN = -4, RTT_E_SENSOR = -5 };

/* arithmetic mode: FAST lets the compiler contract a*b+c into FMA and uses reciprocal
 * multiplies; EXACT keeps every rounding step of the reference's eager fp32 ops (separately
 * rounded mul/add, IEEE div/sqrt) — used by the parity tests to get bit-exact masks. */
enum { RTT_MODE_FAST = 0, RTT_MODE_EXACT = 1 };
/* Optional hint OR-ed into `mode` of rtt_trace_seq_bwd: the caller guarantees that no row of the table requests
 * pose gradients (RTT_FLAG_GRAD_POSE_E / _POSE_S) — the usual lens-design case (curvatures, conic constants, radii,
 * indices).  The adjoint then runs a build without the pose-gradient code (72 instead of 80+ registers, four blocks
 * per SM); pose flags present in the table are ignored under this hint.  Every other entry ignores the bit. */
enum { RTT_MODE_SCALAR_GRADS = 0x100, RTT_MODE_ARITH_MASK = 0xff };
/* Optional kernel-build selector in bits 16..23 of `mode` (0 = the library's default choice for the table): which
 * compiled instantiation of the sequential forward / adjoint kernel runs.  Results are the same for every value
 * (FAST builds agree to rounding); it exists for performance sweeps and A/B measurements, and it is the ONLY tuning
 * input — the library reads no environment variables.
 *   rtt_trace_seq_fwd (FAST): 0 / 12 = tile kernel, 2 rays/thread, one persistent block of 1024 threads per SM
 *     (default); 7 / 13 = the same in four / two waves of blocks; 6 = 7 with a barrier per tile; 8 = two blocks of 512; 1 = tile, 1 ray/thread, 4 blocks of 256 per SM; 2 = tile,
 *     2 rays, 3 blocks; 3 = tile, 2 rays, 4 blocks; 5 = tile, 1 ray, 5 blocks; 9 = per-ray kernel in the reference's
 *     operation order; 16 = packed ray pairs (f32x2 arithmetic) with bulk-async (TMA) ray streaming, 3 blocks/SM;
 *     17 = same, 4 blocks/SM; 18 = packed pairs with plain global loads / stores
 *   rtt_trace_seq_bwd: low 3 bits 2 / 3 / 4 = resident blocks per SM the adjoint build is compiled for; bit 8 (value 8) =
 *     no lean path (every ray through the general adjoint, csrc/rtt_lean.cuh); bit 16 = the lean build in 256-thread
 *     blocks whatever the launch size (default: one 1024-thread block per SM from 2^25 rays on)
 *   rtt_trace_nonseq_fwd: 0 = one block of 1024 threads per SM (EXACT: with a barrier per bounce trip); 7 = 256-thread
 *     blocks running free (the round-1 kernel); 1, 3..6 = other barrier placements (A/B) */
enum { RTT_MODE_TUNE_SHIFT = 16, RTT_MODE_TUNE_MASK = 0xff0000 };
/* rtt_trace_nonseq_fwd / _bwd only: run the FAST arithmetic (see there); every other entry ignores the bit. */
enum { RTT_MODE_NONSEQ_FAST = 0x400 };

/* Sensor image request for one sensor slot.  Bin rule (restating the fixed-range
 * histogram of gui/workbench.py:615-624 in fp32):
 *   ix = floor((x - x0) * sx), iy = floor((y - y0) * sy), kept iff 0<=ix<W and 0<=iy<H,
 *   image[channel, iy, ix] += weight, weight = ray intensity BEFORE the sensor
 *   (elements/sensor.py:36), channel = wavelength index (0 without a wavelength LUT). */
typedef struct {
    float* image;      /* [channels, height, width] fp32, accumulated into (caller zeroes), or NULL */
    float* record;     /* [record_hits, n, 4] (hit_local x,y,z, weight): the k-th interaction of ray i with
                          this sensor goes to record[k][i] for k < record_hits; NULL = no records       */
    int32_t height, width, channels;
    float x0, y0, sx, sy;
    int32_t record_hits; /* K >= 1 (0 is read as 1).  A sequential trace visits a sensor row once (K = 1);
                            in a non-sequential trace a ray can cross a sensor several times — and in the
                            reference it routinely re-hits the plane it just left (t > 1e-6 at an fp32 ulp
                            of ~8e-6), which its hit lists record as two entries                       */
    uint8_t* count;    /* [n] number of interactions of ray i with this sensor (saturates at 255), written
                          by the non-sequential trace only, or NULL                                       */
} rtt_sensor_t;

/* Ray source: the bundle is GENERATED inside the kernels instead of being read from memory
 * (rays/bundle.py:30-171 `Bundle.sample`, render/camera.py:39-72 `Camera.generate_rays`).
 * Ray i of a launch draws Philox4x32-10(counter = first + i [+ state[1]], key = seed [state[0]]),
 * so a trace and its adjoint (and every rank of a sharded bundle) regenerate identical rays
 * from 16 bytes of state: a 1e9-ray trace reads no ray input from HBM.
 *   local sample       DISK  : pos = (r cos th, r sin th, 0), th ~ U(a[2],a[3]), r = sqrt(U(a[0],a[1])); dir = +z
 *                      LINE  : pos = (U(-a[0],a[0]), 0, 0); dir = +z
 *                      FAN   : pos = 0; dir = (0, sin th, cos th), th ~ U(-a[0],a[0])
 *                      POINT : pos = 0; dir = (cos th sin ph, sin th sin ph, cos ph),
 *                              ph = acos(1 - 2 U(a[0],a[1])), th ~ U(a[2],a[3])
 *   local -> global    pos @ R^T + T, dir @ R^T  (geom/transform.py:245-276), dir renormalised (rays/ray.py:25)
 *                      CAMERA: pixel (first+i) mod (W*H), sample (first+i) div (W*H); x = linspace(-a[0],a[0],W)[px],
 *                              y = linspace(a[1],-a[1],H)[py] (+ U(-.5,.5) pixel jitter for sample > 0);
 *                              dir = normalise(x*R[0:3] + y*R[3:6] + R[6:9]), pos = T   (rows of R = right, up, forward) */
enum { RTT_SRC_DISK = 0, RTT_SRC_LINE = 1, RTT_SRC_FAN = 2, RTT_SRC_POINT = 3, RTT_SRC_CAMERA = 4 };
typedef struct {
    int32_t kind;            /* RTT_SRC_*                                                           */
    float a[4];              /* kind parameters, see above                                          */
    int32_t width, height;   /* CAMERA only                                                         */
    const float* pose;       /* DEVICE [12]: R row-major [9], T [3]                                 */
    uint64_t seed;           /* Philox key                                                          */
    int64_t first;           /* counter of this launch's ray 0 (shard offset)                       */
    const uint64_t* state;   /* optional DEVICE {key, counter}: replaces `seed`, is added to `first`
                                (lets a captured CUDA graph draw fresh rays on every replay)        */
    float intensity;         /* intensity of every generated ray (Rays.initialize default 1)        */
    float wavelength;        /* wavelength of every generated ray (default 0)                       */
} rtt_source_t;

typedef struct {
    const float* f;        /* [n_rows, RTT_ROW_F] */
    const int32_t* i;      /* [n_rows, RTT_ROW_I] */
    int32_t n_rows;
    int32_t n_lut;         /* L sample wavelengths, 0 = no per-wavelength index          */
    const float* lut;      /* [L, n_rows, 2] (ior_in, ior_out) or NULL                   */
    const float* lut_w;    /* [L] sample wavelengths, same unit as the rays' wavelength  */
} rtt_table_t;

/* Library identity / layout handshake. which: 0 ROW_F, 1 ROW_I, 2 ROW_G, 3 MAX_ROWS,
 * 4 N_DIFF, 5 MAX_SENSORS, 6 MAX_WAVELENGTHS, 7 MAX_BOUNCES; returns -1 for unknown. */
int rtt_version(void);
int rtt_layout_query(int which);
const char* rtt_error_string(int code);
/* Number of kernels this library has launched since load (host counter; bench.py reports it). */
int64_t rtt_launch_count(void);

/* SequentialScene.simulate(rays) -> rays          (scene/sequential.py:12-36)
 * Every ray walks all n_rows rows once, in order, state in registers.  Rays that miss a
 * row are left untouched; dead rays keep walking (reference behaviour).
 *   in_*        : pos/dir [n,3], intensity [n], wavelength [n] (may be NULL iff n_lut==0)
 *   out_*       : same shapes (may alias the inputs); all three NULL = the final rays are not written (goal
 *                 evaluations that only read sensor records / images)
 *   hitmask     : [n] uint64, bit r set iff the ray interacted with row r (may be NULL)
 *   sensors     : HOST array of n_sensors requests indexed by the rows' sensor slot
 *   source      : NULL, or a ray source that replaces the four in_* arrays (which may then be NULL);
 *                 with a source the out_* arrays may be NULL too (sensor records / images only)  */
int rtt_trace_seq_fwd(const float* in_pos, const float* in_dir, const float* in_intensity,
                      const float* in_wavelength, const rtt_source_t* source,
                      float* out_pos, float* out_dir, float* out_intensity, uint64_t* hitmask,
                      const rtt_table_t* table, const rtt_sensor_t* sensors, int32_t n_sensors,
                      int64_t n, int32_t mode, void* stream);

/* Adjoint of rtt_trace_seq_fwd (replaces torch autograd over scene/sequential.py:12-36 as
 * driven by tests/test_optimize_singlet.py:66-116 / optim/goals.py:144-187).  Recomputes the
 * forward per ray from the ORIGINAL inputs and the recorded hitmask, then sweeps the rows in
 * reverse.
 *   g_out_*     : upstream gradients of out_pos/out_dir/out_intensity ([n,3],[n,3],[n]); NULL = zero
 *   g_record    : HOST array [n_sensors] of device pointers [n,4]: upstream gradient of each
 *                 sensor record (d/d hit_local xyz, d/d weight); entries or the array may be NULL
 *   g_in_*      : gradients w.r.t. the input rays (NULL to skip)
 *   g_table     : [n_rows, RTT_ROW_G] fp32, ACCUMULATED into (caller zeroes); entries
 *                 [0, RTT_N_DIFF) are d/d table_f; NULL to skip
 *   g_lut       : [L, n_rows, 2] accumulated gradient of the wavelength LUT, or NULL
 *   mode        : RTT_MODE_FAST / RTT_MODE_EXACT, optionally | RTT_MODE_SCALAR_GRADS             */
int rtt_trace_seq_bwd(const float* in_pos, const float* in_dir, const float* in_intensity,
                      const float* in_wavelength, const rtt_source_t* source, const uint64_t* hitmask,
                      const float* g_out_pos, const float* g_out_dir, const float* g_out_intensity,
                      const float* const* g_record,
                      float* g_in_pos, float* g_in_dir, float* g_in_intensity,
                      float* g_table, float* g_lut,
                      const rtt_table_t* table, int32_t n_sensors,
                      int64_t n, int32_t mode, void* stream);

/* Scene.simulate() / step() / ray_cast()             (scene/base.py:129-235)
 * Per ray: up to nbounces times { stop if intensity<=0; nearest valid hit over ALL rows
 * (first row wins ties, NaN distance anywhere = no hit); stop if none; interact }.
 *   hit_seq     : [n, nbounces] uint8, row index per executed bounce, 255 = none (may be NULL)
 *   n_hits      : [n] uint8 number of executed bounces (may be NULL)
 * Sensors: every sensor interaction is accumulated into the image; `record` keeps the first
 * `record_hits` interactions per ray and `count` their total number.
 * The arithmetic bits of `mode` are ignored: the non-sequential trace (and its adjoint) run the EXACT arithmetic.
 * Whether a ray re-hits the surface it is leaving is decided by the reference's t > 1e-6 rule at the fp32 ulp of
 * scene-scale coordinates, i.e. by its exact rounding sequence; FMA contraction or approximate division change hit
 * sequences on ~20 % of the rays.  A caller who accepts that noise (the hit sequences of rays that do NOT depend on
 * the threshold are unchanged, see tests) opts in to the FAST arithmetic (MUFU division / square root, FMA
 * contraction: ~1.4x the throughput) with RTT_MODE_NONSEQ_FAST; trace and adjoint must use the same setting. */
int rtt_trace_nonseq_fwd(const float* in_pos, const float* in_dir, const float* in_intensity,
                         const float* in_wavelength, const rtt_source_t* source,
                         float* out_pos, float* out_dir, float* out_intensity,
                         uint8_t* hit_seq, uint8_t* n_hits,
                         const rtt_table_t* table, const rtt_sensor_t* sensors, int32_t n_sensors,
                         int32_t nbounces, int64_t n, int32_t mode, void* stream);

/* Adjoint of rtt_trace_nonseq_fwd: replays the recorded hit sequence (no search), then
 * reverse sweep.  Same gradient conventions as rtt_trace_seq_bwd; g_record[slot] is the upstream
 * gradient [record_hits[slot], n, 4] of that sensor's records (record_hits: HOST array, NULL = all 1). */
int rtt_trace_nonseq_bwd(const float* in_pos, const float* in_dir, const float* in_intensity,
                         const float* in_wavelength, const rtt_source_t* source,
                         const uint8_t* hit_seq, int32_t nbounces,
                         const float* g_out_pos, const float* g_out_dir, const float* g_out_intensity,
                         const float* const* g_record, const int32_t* record_hits,
                         float* g_in_pos, float* g_in_dir, float* g_in_intensity,
                         float* g_table, float* g_lut,
                         const rtt_table_t* table, int32_t n_sensors,
                         int64_t n, int32_t mode, void* stream);

/* Element.intersectTest(rays) -> [n, K]              (elements/parent.py:30-42,
 * geom/shape.py:25-59, geom/primitives.py:38-57): distances of rows [row0, row0+k) with all
 * validity rules applied, +inf (or NaN, as the reference) on a miss.  t_out is [n, k]. */
int rtt_intersect_test(const float* in_pos, const float* in_dir, float* t_out,
                       const rtt_table_t* table, int32_t row0, int32_t k,
                       int64_t n, int32_t mode, void* stream);

/* Element.forward(rays, surf_idx) -> (new_pos, new_dir, intensity_mult)
 *                                                     (elements/parent.py:44-58)
 * One row, NO shape-level validity (the reference applies it only in intersectTest,
 * geom/shape.py:52 vs :61-87); rays that miss produce inf/NaN exactly like the reference.
 * Optional extra outputs (NULL to skip): hit_local [n,3], t [n], normal [n,3]. */
int rtt_surface_step_fwd(const float* in_pos, const float* in_dir, const float* in_wavelength,
                         float* new_pos, float* new_dir, float* mod,
                         float* hit_local, float* t_out, float* normal,
                         const rtt_table_t* table, int32_t row,
                         int64_t n, int32_t mode, void* stream);

/* Adjoint of rtt_surface_step_fwd.  Upstream gradients may be NULL (= zero). */
int rtt_surface_step_bwd(const float* in_pos, const float* in_dir, const float* in_wavelength,
                         const float* g_new_pos, const float* g_new_dir, const float* g_hit_local,
                         const float* g_t, const float* g_normal,
                         float* g_in_pos, float* g_in_dir, float* g_table, float* g_lut,
                         const rtt_table_t* table, int32_t row,
                         int64_t n, int32_t mode, void* stream);

/* Bundle.sample(N) -> Rays on the device          (rays/bundle.py:30-37, render/camera.py:39-72)
 * Materialises the rays of `source` (the same rays the traces of the same `mode` generate in-kernel from it).
 * pos/dir [n,3], intensity [n]; wavelength [n] may be NULL. */
int rtt_sample_bundle(const rtt_source_t* source, float* pos, float* dir, float* intensity, float* wavelength,
                      int64_t n, int32_t mode, void* stream);

/* Sensor reductions of the optimisation goals      (optim/goals.py:42-96, 99-187;
 * elements/sensor.py:67-176), on raw sensor records rec[m] = (x, y, z, w) (w == 0: no hit).
 * Two-stage deterministic sums: per-block partials in `work` (DEVICE, RTT_SPOT_WORK floats, zeroed
 * once by the caller; the kernels leave it zeroed), last block adds them in a fixed order.
 *
 * rtt_spot_moments:      out[0..3] = (sum w, sum w x, sum w y, #{w > 0}); active_only != 0 drops w <= 0
 *                        (optim/goals.py:165) — SpotTargetLoss takes every recorded hit (:76-88).
 * rtt_spot_moments_bwd:  g_rec[m] = (g[1] w, g[2] w, 0, g[0] + g[1] x + g[2] y) (0 where dropped); g = DEVICE [3]
 * rtt_spot_size_fwd:     with W = max(mom[0], 1e-12), centre c = target (DEVICE [2]) or (mom[1], mom[2]) / W:
 *                        out[0] = sum_i sqrt(q_i), q_i = |xy_i - c|^2 w_i / W over w_i > 0 (optim/goals.py:176-183);
 *                        out[1..2] = d out[0] / d c (needed by the backward)
 * rtt_spot_size_bwd:     g_rec[m] = g_loss[0] * d out[0] / d rec[m], through c and W unless target is given */
enum { RTT_SPOT_WORK = 4 * 1024 + 4 };
int rtt_spot_moments(const float* rec, int64_t m, int32_t active_only, float* out4, float* work, void* stream);
int rtt_spot_moments_bwd(const float* rec, int64_t m, int32_t active_only, const float* g3, float* g_rec, void* stream);
int rtt_spot_size_fwd(const float* rec, int64_t m, const float* mom4, const float* target_xy, float* out3,
                      float* work, void* stream);
int rtt_spot_size_bwd(const float* rec, int64_t m, const float* mom4, const float* target_xy, const float* out3,
                      const float* g_loss, float* g_rec, void* stream);

/* Per-id sensor moments: Sensor.getSpotSizeParallel_xy (elements/sensor.py:87-176: isin + sort + searchsorted +
 * scatter_add over the hit lists) as one pass per reduction over the dense records.
 *   rec      : [m,4] sensor records (x, y, z, w), w == 0 = no hit (every sum below is weighted by w)
 *   ids      : [m] int8 ray ids (rays/ray.py:17)
 *   group_of : DEVICE int32 [256], group_of[id + 128] = index of `id` in the caller's query list, -1 = not queried
 *   n_groups : K = number of queried ids, 1..256
 *   work     : DEVICE scratch of RTT_SPOT_ID_WORK floats, zero before the first call (left zeroed), one per stream
 * rtt_spot_id_moments:  out [K,4] = per group (sum w, sum w x, sum w y, #{w > 0})
 * rtt_spot_id_size:     with centres [K,2] (centroids or targets) and norm order p >= 1:
 *                       out [K,4] = (sum w (|dx|^p + |dy|^p), sum w p |dx|^(p-1) sgn dx, sum w p |dy|^(p-1) sgn dy, 0);
 *                       the reference's result is out[k][0] / (2 max-or-one(sum w)), entries 1, 2 feed the adjoint
 * rtt_spot_id_size_bwd: coef [K,8] = (cx, cy, a, bx, by, sW, 0, 0) per group ->
 *                       g_rec[i] = (a w (p|dx|^(p-1) sgn dx - bx), a w (p|dy|^(p-1) sgn dy - by), 0,
 *                                   a (|dx|^p + |dy|^p - bx dx - by dy - sW)), zero for ids that were not queried */
enum { RTT_SPOT_ID_WORK = 296 * 256 * 4 + 4 };
int rtt_spot_id_moments(const float* rec, const int8_t* ids, int64_t m, const int32_t* group_of, int32_t n_groups,
                        float* out, float* work, void* stream);
int rtt_spot_id_size(const float* rec, const int8_t* ids, int64_t m, const int32_t* group_of, int32_t n_groups,
                     const float* centres, float norm_ord, float* out, float* work, void* stream);
int rtt_spot_id_size_bwd(const float* rec, const int8_t* ids, int64_t m, const int32_t* group_of, int32_t n_groups,
                         const float* coef, float norm_ord, float* g_rec, void* stream);

/* Measurement helper (bench.py): launches a pure-FMA kernel (8 independent chains per thread,
 * 8 blocks of 256 threads per SM, `iters` x 64 FMAs per thread) on `stream` and returns the FLOPs
 * it executes (FMA = 2), or a negative code.  Timed with CUDA events by the caller, this is the
 * measured FP32 peak the roofline of the FP32-issue-bound traces is reported against.
 * `scratch`: one device float (never written in practice). */
int64_t rtt_probe_fp32(int32_t iters, float* scratch, void* stream);

/* Renderer.render_3d (render/camera.py:191-257): per ray the nearest hit over all table rows (the caller passes the
 * table of the renderable, non-aperture elements), the winner's global normal (Shape.forward, geom/shape.py:61-87) and
 * the colour clamp(base_rgb[row] * (0.3 + 0.7 |n . light|), 0, 1); rays without a hit get `background`.
 *   in_pos / in_dir : [n,3] pixel rays, or NULL with `source` = a CAMERA rtt_source_t (rays generated in the kernel)
 *   base_rgb        : DEVICE [n_rows,3]; light_dir, background: HOST float[3] (light_dir normalised by the caller)
 *   out_rgb         : [n,3];  out_row: [n] uint8 winning row, 255 = background, or NULL
 * Always runs the EXACT arithmetic (the nearest-hit decision is the non-sequential search's, see rtt_trace_nonseq_fwd). */
int rtt_render_shade(const float* in_pos, const float* in_dir, const rtt_source_t* source, const rtt_table_t* table,
                     const float* base_rgb, const float* light_dir, const float* background,
                     float* out_rgb, uint8_t* out_row, int64_t n, int32_t mode, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* RTT_B200_H */
