"""Algorithmic work of one trace, for roofline reporting (bench.py, DESIGN.md section "Roofline").

Counting rule: add / sub / mul / abs / neg = 1, FMA = 2, div / sqrt = 1,
compares / selects / moves = 0.  The per-piece counts below are hand counts of the operations in
``csrc/rtt_core.cuh`` (the function each entry covers is named); they describe the ALGORITHM —
what any implementation of the reference's formulas must compute once per ray and table row —
not the instructions a particular build issues.

Bytes: a sequential trace reads pos 12 + dir 12 + intensity 4 (+ wavelength 4 when a per-wavelength
index table is in use) per ray and writes pos 12 + dir 12 + intensity 4 (+ 8 hit mask when
requested, + 16 per sensor record when hit lists are materialised): 56..84 B/ray, independent of
the number of rows.  The adjoint reads the same inputs + hit mask 8 + upstream gradients 28 and
writes input-ray gradients 28.
"""
from __future__ import annotations

from typing import Dict, Sequence

from . import codes as C

# ---- pieces ---------------------------------------------------------------------------------
POSE_SUB = 3          # p - T
POSE_ROT = 15         # one [3]@[3,3]: 3 x (mul + 2 fma)            mul_R / mul_RT
NORMALIZE = 10        # norm3 (mul + 2 fma + sqrt) + max + 3 div      normalize12
ALONG = 6             # p + t*d                                       along
ROOTS = {             # solve_roots
    C.SURF_PLANE: 2, C.SURF_SPHERE: 18, C.SURF_CYLINDER: 25, C.SURF_QUADRIC: 37, C.SURF_QUADRIC_ZY: 29,
    C.SURF_CONE: 38}
BOUND = {             # surface_in_bounds, per candidate root (ALONG added separately)
    C.BOUND_NONE: 0, C.BOUND_DISK: 3, C.BOUND_RECT: 2, C.BOUND_ELLIPSE: 11, C.BOUND_HALF: 2, C.BOUND_HALF_DISK: 5,
    C.BOUND_NAPPE: 1}
SHAPE = {             # shape_in_bounds (ALONG added separately); POLY is per sibling plane
    C.SHAPE_NONE: 0, C.SHAPE_SPHERIC_FACE: 3, C.SHAPE_SPHERIC_EDGE: 0, C.SHAPE_CYL_FACE: 4, C.SHAPE_CYL_EDGE: 28,
    C.SHAPE_POLY: 8, C.SHAPE_OPEN: 0}
NORMAL = {            # normal_local
    C.SURF_PLANE: 0, C.SURF_SPHERE: 3, C.SURF_CYLINDER: 2, C.SURF_QUADRIC: 18, C.SURF_QUADRIC_ZY: 15,
    C.SURF_CONE: 14}
PHYS = {              # physics
    C.PHYS_TRANSMIT: 0, C.PHYS_SNELL: 27, C.PHYS_REFLECT: 12, C.PHYS_BLOCK: 0, C.PHYS_APERTURE: 6,
    C.PHYS_LINEAR: 46}   # two [3]@[3,3], 2 div, 2 x (mul + mul + add), normalize
ADJOINT_FACTOR = 3.5  # reverse sweep ~2.5x the forward interaction + 1x recompute (interact_adjoint)


def _is_identity(row_f: Sequence[float], off: int) -> bool:
    want = (1.0, 0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0, 1.0)
    return all(float(row_f[off + a]) == want[a] for a in range(9))


def row_costs(row_f: Sequence[float], row_i: Sequence[int]) -> Dict[str, int]:
    """FLOPs of (test: distance + every validity rule) and (interact: hit, normal, physics) for one ray."""
    surf, bound, shape, phys = row_i[C.I_SURF], row_i[C.I_BOUND], row_i[C.I_SHAPE], row_i[C.I_PHYS]
    ident_e, ident_s = _is_identity(row_f, C.F_RE), _is_identity(row_f, C.F_RS)
    n_roots = 1 if surf == C.SURF_PLANE else 2
    test = POSE_SUB + (0 if ident_s else 2 * POSE_ROT)
    if shape != C.SHAPE_NONE:
        test += POSE_SUB + (0 if ident_e else 2 * POSE_ROT) + NORMALIZE
    test += ROOTS[surf]
    if bound != C.BOUND_NONE:
        test += n_roots * (ALONG + BOUND[bound])
    if shape != C.SHAPE_NONE:
        mult = max(int(row_i[C.I_POLY_COUNT]) - 1, 0) if shape == C.SHAPE_POLY else 1
        test += ALONG + SHAPE[shape] * mult
    interact = ALONG + NORMAL[surf] + ALONG + PHYS[phys]
    if surf != C.SURF_PLANE or not (ident_e and ident_s):
        interact += (0 if ident_s else POSE_ROT) + (0 if (ident_e or shape == C.SHAPE_NONE) else POSE_ROT)
    return dict(test=test, interact=interact)


def sequential_flops_per_ray(table_f, table_i, hit_fraction: Sequence[float]) -> float:
    """Forward FLOPs per ray: every row tested once, interaction on the fraction of rays that hit."""
    tot = 0.0
    for r in range(len(table_i)):
        c = row_costs(table_f[r], table_i[r])
        tot += c["test"] + float(hit_fraction[r]) * c["interact"]
    return tot


def nonsequential_flops_per_ray(table_f, table_i, searches_per_ray: float, hits_per_row: Sequence[float]) -> float:
    """Non-sequential forward: every executed nearest-hit search tests all rows (scene/base.py:164-176), every bounce
    then interacts with its winner.  ``searches_per_ray`` = executed searches per ray (bounces + the final miss),
    ``hits_per_row[r]`` = interactions per ray with row r."""
    tot = 0.0
    for r in range(len(table_i)):
        c = row_costs(table_f[r], table_i[r])
        tot += searches_per_ray * c["test"] + float(hits_per_row[r]) * c["interact"]
    return tot


def adjoint_flops_per_ray(table_f, table_i, hit_fraction: Sequence[float]) -> float:
    """Adjoint kernel: replay of the recorded interactions (roots + interaction, no validity rules)
    followed by the reverse sweep."""
    tot = 0.0
    for r in range(len(table_i)):
        c = row_costs(table_f[r], table_i[r])
        tot += float(hit_fraction[r]) * (c["test"] + c["interact"]) * ADJOINT_FACTOR
    return tot


# ---- the per-row constants the judge recomputes the roofline with -----------------------
# refracting / reflecting quadric row: 160 FLOP per ray-surface with a general pose, 115 with identity rotation;
# planar row (aperture / sensor / ideal element): 60; lens-edge row (cylinder or box / side plane): 40.
# These are deliberately coarser (and, for edge rows, LOWER) than the hand counts above: an edge row that the kernel
# culls is charged 40, not the 50-110 the full test costs, so `frac` by this counting does not reward skipped work.
SURVEY_QUADRIC_GENERAL, SURVEY_QUADRIC_IDENT, SURVEY_PLANAR, SURVEY_EDGE = 160.0, 115.0, 60.0, 40.0
SURVEY_ADJOINT_FACTOR = 3.5


def survey_row_flops(row_f: Sequence[float], row_i: Sequence[int]) -> float:
    surf, shape = row_i[C.I_SURF], row_i[C.I_SHAPE]
    if shape in (C.SHAPE_SPHERIC_EDGE, C.SHAPE_CYL_EDGE, C.SHAPE_POLY):
        return SURVEY_EDGE
    if surf == C.SURF_PLANE:
        return SURVEY_PLANAR
    ident = _is_identity(row_f, C.F_RE) and _is_identity(row_f, C.F_RS)
    return SURVEY_QUADRIC_IDENT if ident else SURVEY_QUADRIC_GENERAL


def survey_flops_per_ray(table_f, table_i) -> float:
    """Forward FLOPs per ray by SURVEY 8(d)'s constants: every row charged once per ray."""
    return float(sum(survey_row_flops(table_f[r], table_i[r]) for r in range(len(table_i))))


def survey_adjoint_flops_per_ray(table_f, table_i, hit_fraction: Sequence[float]) -> float:
    """Adjoint by SURVEY 8(d): 3.5 x the forward constant of every row a live ray interacted with."""
    return float(sum(float(hit_fraction[r]) * survey_row_flops(table_f[r], table_i[r]) * SURVEY_ADJOINT_FACTOR
                     for r in range(len(table_i))))


def sequential_bytes_per_ray(*, wavelength: bool, hitmask: bool, records: int = 0) -> int:
    return 28 + (4 if wavelength else 0) + 28 + (8 if hitmask else 0) + 16 * records


def adjoint_bytes_per_ray(*, wavelength: bool, records: int = 0) -> int:
    return 28 + (4 if wavelength else 0) + 8 + 28 + 28 + 16 * records


def fp32_peak_tflops(n_sm: int, sm_mhz: float) -> float:
    """n_SM x 128 FP32 lanes x 2 (FMA) x clock."""
    return n_sm * 128 * 2 * sm_mhz * 1e6 / 1e12
