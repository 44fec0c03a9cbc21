"""Numeric codes and row layout of the flat surface table.

Single source of truth on the Python side; ``include/rtt_b200.h`` carries the same
numbers for the C-ABI / CUDA side and ``tests/test_cabi.py`` checks the two agree
(the shared library exports them through ``rtt_layout_query``).

One table row = one ``(element, surface index)`` pair, in the order of the
reference's flattening (``scene/base.py:116-123`` and the double loop in
``scene/sequential.py:17-19``).
"""

# ---- float part of a row -------------------------------------------------------
ROW_F = 48          # floats per row
F_RE = 0            # 3x3 element (shape) rotation, row-major     geom/shape.py:37
F_TE = 9            # element translation
F_RS = 12           # 3x3 surface rotation, row-major             geom/primitives.py:49
F_TS = 21           # surface translation
F_C = 24            # quadric curvature                           geom/primitives.py:263
F_K = 25            # quadric conic constant                      geom/primitives.py:264
F_RADIUS = 26       # Sphere / Cylinder radius                    geom/primitives.py:153,199
F_IOR_IN = 27       # RefractSnell.ior_in                         phys/std.py:120
F_IOR_OUT = 28      # RefractSnell.ior_out                        phys/std.py:121
N_DIFF = 29         # entries [0, N_DIFF) receive gradients from the adjoint kernel
F_SB = 29           # 4 floats: surface-level bound parameters (non-differentiable selections)
F_HB = 33           # 8 floats: shape-level bound parameters
ROW_G = 48          # floats per row of the gradient table returned by the adjoint (== ROW_F)

# ---- int part of a row ---------------------------------------------------------
ROW_I = 16
I_SURF = 0          # surface kind
I_BOUND = 1         # surface-level bound kind
I_INVERT = 2        # SurfaceBounded.invert                        geom/bounded.py:29
I_SHAPE = 3         # shape-level bound kind (0 = element's shape is a bare Surface)
I_PHYS = 4          # surface function kind
I_SENSOR = 5        # sensor slot (>=0) or -1
I_POLY_FIRST = 6    # first sibling row of a convex polyhedron
I_POLY_COUNT = 7    # number of sibling planes
I_ELEM = 8          # element index   (scene/base.py map_to_element)
I_SIDX = 9          # surface index   (scene/base.py map_to_surface)
I_FLAGS = 10        # bit field, see FLAG_*

SURF_PLANE, SURF_QUADRIC, SURF_QUADRIC_ZY, SURF_CYLINDER, SURF_SPHERE, SURF_CONE = 0, 1, 2, 3, 4, 5
BOUND_NONE, BOUND_DISK, BOUND_RECT, BOUND_ELLIPSE, BOUND_HALF, BOUND_HALF_DISK, BOUND_NAPPE = 0, 1, 2, 3, 4, 5, 6
# SURF_CONE rows (geom/primitives.py:398-494) keep the slope in the F_C slot; BOUND_NAPPE = SingleCone
(SHAPE_NONE, SHAPE_SPHERIC_FACE, SHAPE_SPHERIC_EDGE, SHAPE_CYL_FACE, SHAPE_CYL_EDGE,
 SHAPE_POLY, SHAPE_OPEN) = 0, 1, 2, 3, 4, 5, 6
PHYS_TRANSMIT, PHYS_SNELL, PHYS_REFLECT, PHYS_BLOCK, PHYS_APERTURE, PHYS_LINEAR, PHYS_FRESNEL = 0, 1, 2, 3, 4, 5, 6
I_RNG_LO, I_RNG_HI = 14, 15   # row 0 only: 64-bit seed of the Fresnel reflect / refract draws
# PHYS_LINEAR rows (phys/std.py:35-88) keep Cx, Cy, Dx, Dy in the F_C, F_K, F_RADIUS, F_IOR_IN slots

# gradient-request flags (host knows them from requires_grad; no device sync needed)
FLAG_GRAD_POSE_E = 1 << 0   # Re / Te
FLAG_GRAD_POSE_S = 1 << 1   # Rs / Ts
FLAG_GRAD_CK = 1 << 2       # c, k
FLAG_GRAD_RADIUS = 1 << 3
FLAG_GRAD_IOR = 1 << 4

MAX_ROWS = 64               # rows staged in shared memory per launch
MAX_SENSORS = 4
MAX_WAVELENGTHS = 8
MAX_BOUNCES = 255           # hit sequence is stored one byte per bounce

# ray source kinds (rtt_source_t.kind) and the scratch size of the goal reductions
SRC_DISK, SRC_LINE, SRC_FAN, SRC_POINT, SRC_CAMERA = 0, 1, 2, 3, 4
SPOT_ID_WORK = 296 * 256 * 4 + 4    # floats of scratch per stream of the per-id reductions (RTT_SPOT_ID_WORK)
SPOT_WORK = 4 * 1024 + 4
