"""Scene definitions shared by the golden generator and the parity tests.

Every builder takes a namespace ``ns`` with ``elements``, ``geom``, ``phys``, ``rays``
attributes, so the *same* definition can be instantiated from the unmodified reference
(``oracle/ref_loader.load_reference()``, golden generation only) and from this repo's
``raytracetorch_b200`` package (tests, bench).  Values follow SURVEY.md section 8(d).
"""
from __future__ import annotations

import math

import torch


def _T(ns, z=0.0, x=0.0, y=0.0, rot=None, **kw):
    return ns.geom.RayTransform(translation=[x, y, z], rotation=rot, **kw)


def _adhoc(ns, shape, make_fn):
    """Ad-hoc element the way the reference's scripts build them (tests/render_static.py:26-38)."""
    el = ns.elements.Element()
    el.shape = shape
    for _ in range(len(shape)):
        el.surface_functions.append(make_fn())
    return el


# ---- C1: singlet + sensor (tests/test_optimize_singlet.py:29-49 values) ----------------
def c1_singlet(ns, physical=False, grads=False, inked=True, fresnel=False):
    E = ns.elements
    glass, media = (1.0, 1.5168) if physical else (1.5168, 1.0)
    lens = E.SingletLens(c1=0.016667, c2=-0.00283, d=25.4, t=4.0, ior_glass=glass, ior_media=media,
                         inked=inked, c1_grad=grads, c2_grad=grads, fresnel=fresnel)
    sensor = E.Sensor(ns.geom.Disk(radius=20.0, transform=_T(ns, 100.0)))
    return [lens, sensor]


# ---- C2: crossed cylindrical pair + inverted stop + rectangular sensor ------------------
def c2_cylindrical(ns, grads=False):
    E = ns.elements
    l1 = E.CylSingletLens(0.02, -0.02, 20.0, 20.0, 4.0, 1.5, c1_grad=grads, transform=_T(ns, 0.0))
    l2 = E.CylSingletLens(0.03, 0.0, 20.0, 20.0, 3.0, 1.6,
                          transform=_T(ns, 10.0, rot=[0.0, 0.0, math.pi / 2],
                                       trans_grad=grads, rot_grad=grads))
    stop = E.CircularAperture(6.0, invert=True, transform=_T(ns, 15.0))
    sensor = E.Sensor(ns.geom.Rectangle(15.0, 15.0, transform=_T(ns, 40.0)))
    return [l1, l2, stop, sensor]


C2_WAVELENGTHS = (486.1, 587.6, 656.3)
C2_GLASS_SCALE = (1.5224 / 1.5168, 1.0, 1.5143 / 1.5168)   # F, d, C lines relative to d


# ---- C4: doublet + stop + triplet + singlet + 16:9 sensor (17 rows) ---------------------
def c4_camera_lens(ns, grads=False):
    E = ns.elements
    dbl = E.DoubletLens(1 / 31, -1 / 18.5, -1 / 140, d=20.0, t1=6.0, t2=2.5, ior_glass1=1.517, ior_glass2=1.648,
                        c1_grad=grads, c3_grad=grads, transform=_T(ns, 0.0))
    stop = E.CircularAperture(8.0, invert=True, transform=_T(ns, 8.0))
    trp = E.TripletLens(1 / 40, 1 / 30, 1 / 30, -1 / 40, d=20.0, t1=4.0, t2=2.0, t3=4.0,
                        ior_glass1=1.5, ior_glass2=1.6, ior_glass3=1.5, c2_grad=grads, transform=_T(ns, 20.0))
    sng = E.SingletLens(-0.01, 0.01, d=20.0, t=3.0, ior_glass=1.5, inked=True, transform=_T(ns, 32.0))
    sensor = E.Sensor(ns.geom.Rectangle(12.0, 6.75, transform=_T(ns, 45.0)))
    return [dbl, stop, trp, sng, sensor]


# ---- C5: non-sequential mirror / lens / absorbing box / stop / sensor (12 rows) ---------
def c5_nonsequential(ns):
    E, G, P = ns.elements, ns.geom, ns.phys
    mirror = E.SphericalMirror(c1=-1 / 80, d=30.0, diameter=30.0, transform=_T(ns, 60.0))
    lens = E.SingletLens(0.02, -0.02, d=20.0, t=4.0, ior_glass=1.0, ior_media=1.5, transform=_T(ns, 20.0))
    box = _adhoc(ns, G.Box(4.0, 4.0, 4.0, transform=_T(ns, 40.0, y=8.0)), P.Block)
    stop = E.CircularAperture(9.0, invert=True, transform=_T(ns, 10.0))
    sensor = E.Sensor(G.Disk(3.0, transform=_T(ns, 30.0)))
    return [mirror, lens, box, stop, sensor]


# ---- light pipe: the four mirrored side planes of a Box4Side (geom/shape.py:213-276), tilted, + sensor ----------
def x5_light_pipe(ns):
    """Scene scale ~0.1-1: the fp32 ulp of the coordinates (< 1e-7) is below the reference's t > 1e-6 rule, so no ray
    re-hits the wall it just left and the hit sequences do not depend on the last bit of the poses
    (at scale 20-70 a 1-ulp change of the tilted planes' rotation re-routes 3 rays in 4)."""
    E, G, P = ns.elements, ns.geom, ns.phys
    pipe = _adhoc(ns, G.Box4Side(0.06, 0.04, transform=_T(ns, 0.2, x=0.003, rot=[0.02, -0.03, 0.1])), P.Reflect)
    sensor = E.Sensor(G.Disk(0.12, transform=_T(ns, 0.7)))
    stop = _adhoc(ns, G.Disk(0.2, transform=_T(ns, 0.75)), P.Block)    # the four planes form an endless tube: end the paths
    return [pipe, sensor, stop]


# ---- benchmark scene of the reference (benchmarks/sim_benchmark.py:56-88) ---------------
def sim_benchmark_scene(ns):
    E, G = ns.elements, ns.geom
    lens = E.SingletLens(c1=0.05, c2=-0.05, d=10.0, t=3.0, ior_glass=1.5, ior_media=1.0)
    ap = E.CircularAperture(radius=5.0, transform=_T(ns, 0.0))
    sensor = E.Sensor(G.Disk(radius=6.0, transform=_T(ns, 19.0)))
    return [lens, ap, sensor]


# ---- extra coverage: tilted/decentred parts, every surface / bound / physics kind -------
def x1_mirrors(ns):
    """Folded path: parabolic mirror, cylindrical mirror, XZ parabola, elliptic + rect stops."""
    E, G = ns.elements, ns.geom
    m1 = E.ParabolicMirror(c1=-1 / 120, d=30.0, transform=_T(ns, 50.0, rot=[0.05, -0.03, 0.0]))
    m2 = E.CylindricalMirror(c1=1 / 90, d=30.0, transform=_T(ns, 5.0, y=4.0, rot=[-0.04, 0.0, 0.2]))
    m3 = E.ParabolicMirrorXZ(c1=-1 / 200, d=30.0, transform=_T(ns, 70.0))
    a1 = E.EllipticAperture(9.0, 6.0, rot=0.3, invert=True, transform=_T(ns, 40.0, x=0.5))
    a2 = E.RectangularAperture(7.0, 5.0, invert=False, transform=_T(ns, 30.0))
    sensor = E.Sensor(G.Rectangle(30.0, 30.0, transform=_T(ns, 20.0, rot=[0.0, 0.1, 0.0])))
    return [m1, a1, m2, m3, a2, sensor]


def x2_tilted_lenses(ns, grads=False):
    """Tilted + decentred singlet and doublet (general poses on both levels), un-inked edge,
    sphere + plane primitives with Reflect/Transmit."""
    E, G, P = ns.elements, ns.geom, ns.phys
    l1 = E.SingletLens(0.03, -0.02, d=18.0, t=5.0, ior_glass=1.0, ior_media=1.6, inked=False,
                       c1_grad=grads, t_grad=grads, ior_media_grad=grads,
                       transform=_T(ns, 0.0, x=0.7, y=-0.4, rot=[0.06, -0.05, 0.3],
                                    trans_grad=grads, rot_grad=grads))
    l2 = E.DoubletLens(1 / 45, -1 / 35, -1 / 90, d=16.0, t1=4.0, t2=2.0, ior_glass1=1.52, ior_glass2=1.67,
                       c2_grad=grads, ior_glass2_grad=grads,
                       transform=_T(ns, 14.0, x=-0.3, rot=[-0.03, 0.04, 0.0]))
    ball = _adhoc(ns, G.Sphere(3.0, transform=_T(ns, 30.0, x=4.0, y=3.0)), P.Reflect)
    window = _adhoc(ns, G.Plane(transform=_T(ns, 36.0, rot=[0.1, 0.0, 0.0])), P.Transmit)
    sensor = E.Sensor(G.Disk(25.0, transform=_T(ns, 60.0)))
    return [l1, l2, ball, window, sensor]


def x3_ideal(ns, grads=False):
    """Ideal (paraxial) elements, elements/ideal.py: a tilted + decentred thin lens with a finite aperture, an
    unbounded thin lens, a paraxial mirror (Linear physics keeps rays going towards +z), then a sensor."""
    E, G = ns.elements, ns.geom
    l1 = E.IdealThinLens(focal=60.0, focal_grad=grads, diameter=22.0,
                         transform=_T(ns, 0.0, x=0.4, y=-0.3, rot=[0.03, -0.02, 0.1], trans_grad=grads, rot_grad=grads))
    l2 = E.IdealThinLens(focal=-150.0, focal_grad=grads, transform=_T(ns, 12.0))
    IdealMirror = getattr(E, "IdealMirror", None) or E.ideal.IdealMirror    # the reference does not re-export it
    m = IdealMirror(radius_x=400.0, radius_y=250.0, radius_x_grad=grads, diameter=30.0, transform=_T(ns, 20.0))
    sensor = E.Sensor(G.Disk(25.0, transform=_T(ns, 45.0)))
    return [l1, l2, m, sensor]


def x4_cones(ns, grads=False):
    """Cone primitives (geom/primitives.py:398-494, geom/bounded.py:189-217) as bare surfaces: a reflecting axicon
    (one nappe, tilted + decentred), a transmitting double cone, a refracting shallow nappe, then a sensor."""
    G, P, E = ns.geom, ns.phys, ns.elements
    Cone = getattr(G, "Cone", None) or G.primitives.Cone                 # the reference does not re-export them
    SingleCone = getattr(G, "SingleCone", None) or G.bounded.SingleCone
    # every vertex sits outside the beam: rays through a cone's tip are chaotic in the reference itself (its fp32
    # and fp64 runs end 20 units apart), which would only measure noise
    axicon = _adhoc(ns, SingleCone(slope=0.35, slope_grad=grads,
                                   transform=_T(ns, 30.0, x=-13.0, rot=[0.02, -0.03, 0.0],
                                                trans_grad=grads, rot_grad=grads)), P.Reflect)
    dbl = _adhoc(ns, Cone(slope=2.5, slope_grad=grads, transform=_T(ns, 8.0, y=-14.0)), P.Transmit)
    shallow = _adhoc(ns, SingleCone(slope=0.08, slope_grad=grads, transform=_T(ns, 3.0, x=14.0)),
                     lambda: P.RefractSnell(1.0, 1.5))
    sensor = E.Sensor(G.Disk(90.0, transform=_T(ns, -20.0)))
    return [shallow, dbl, axicon, sensor]


RENDER_CAMERA = ((60.0, 45.0, -70.0), (0.0, 2.0, 35.0), (0.0, 1.0, 0.0), 24.0, 96, 64)


def render_setup(ns, device="cpu"):
    """Scene + camera of the render fixture (oracle/make_golden.py::gen_render_case, tests/test_goals.py): the C5
    elements seen from off-axis."""
    scene = ns.scene.Scene()
    for e in c5_nonsequential(ns):
        scene.add_element(e)
    scene._build_index_maps()
    cam = ns.render.Camera(*RENDER_CAMERA, device=device)
    return scene, cam


# ---- ray bundles --------------------------------------------------------------------------
def bundle_collimated(ns, n, radius, z, seed, tilt=None, ray_id=0):
    torch.manual_seed(seed)
    tr = ns.geom.RayTransformBundle(translation=[0.0, 0.0, z], rotation=tilt)
    return ns.rays.CollimatedDisk(radius, ray_id, transform=tr).sample(n)


def bundle_point(ns, n, na, pos, seed, ray_id=0):
    torch.manual_seed(seed)
    tr = ns.geom.RayTransformBundle(translation=list(pos))
    return ns.rays.PointSource(na, ray_id, transform=tr).sample(n)


CASES = {
    # name: (builder, kwargs, mode, bundle spec)
    "c1_singlet": (c1_singlet, {}, "seq", ("coll", 5.0, -10.0, None)),
    "c1_singlet_physical": (c1_singlet, {"physical": True}, "seq", ("coll", 5.0, -10.0, None)),
    "c1_singlet_wide": (c1_singlet, {"physical": True}, "seq", ("coll", 14.0, -10.0, [0.05, 0.02, 0.0])),
    "c1_singlet_clear_edge": (c1_singlet, {"physical": True, "inked": False}, "seq", ("point", 0.45, (0.0, 1.0, -25.0))),
    "c2_cylindrical": (c2_cylindrical, {}, "seq", ("coll", 8.0, -10.0, None)),
    "c2_cylindrical_tilt": (c2_cylindrical, {}, "seq", ("coll", 13.0, -10.0, [0.03, -0.04, 0.0])),
    "c4_camera_lens": (c4_camera_lens, {}, "seq", ("point", 0.06, (0.0, 0.0, -200.0))),
    "c4_camera_lens_field": (c4_camera_lens, {}, "seq", ("coll", 11.0, -10.0, [0.04, 0.06, 0.0])),
    "x1_mirrors": (x1_mirrors, {}, "seq", ("coll", 12.0, -10.0, [0.0, 0.02, 0.0])),
    "x2_tilted_lenses": (x2_tilted_lenses, {}, "seq", ("coll", 10.0, -12.0, [0.02, 0.03, 0.0])),
    "x4_cones": (x4_cones, {}, "seq", ("coll", 9.0, -10.0, [0.03, 0.02, 0.0])),
    "x3_ideal": (x3_ideal, {}, "seq", ("coll", 13.0, -10.0, [0.02, -0.03, 0.0])),
    "c5_nonsequential": (c5_nonsequential, {}, "nonseq", ("coll", 10.0, -5.0, None)),
    "sim_benchmark": (sim_benchmark_scene, {}, "nonseq", ("coll", 4.0, 0.0, None)),
    "x2_nonsequential": (x2_tilted_lenses, {}, "nonseq", ("coll", 10.0, -12.0, [0.02, 0.03, 0.0])),
    "x5_light_pipe": (x5_light_pipe, {}, "nonseq", ("point", 0.25, (0.002, -0.001, 0.05))),
}


GRAD_CASES = {
    # name: (builder, kwargs, bundle spec) — sequential scenes with trainable parameters
    "grad_c3_singlet": (c1_singlet, {"physical": True, "grads": True}, ("coll", 5.0, -10.0, None)),
    "grad_c3_singlet_ref_order": (c1_singlet, {"grads": True}, ("coll", 5.0, -10.0, None)),
    "grad_c2_cylindrical": (c2_cylindrical, {"grads": True}, ("coll", 8.0, -10.0, [0.01, 0.02, 0.0])),
    "grad_c4_camera_lens": (c4_camera_lens, {"grads": True}, ("coll", 7.0, -10.0, [0.02, 0.03, 0.0])),
    "grad_x2_tilted": (x2_tilted_lenses, {"grads": True}, ("coll", 7.0, -12.0, [0.02, 0.03, 0.0])),
    "grad_x4_cones": (x4_cones, {"grads": True}, ("coll", 7.0, -10.0, [0.03, 0.02, 0.0])),
    "grad_x3_ideal": (x3_ideal, {"grads": True}, ("coll", 9.0, -10.0, [0.02, -0.03, 0.0])),
}


def make_bundle(ns, spec, n, seed):
    if spec[0] == "coll":
        return bundle_collimated(ns, n, spec[1], spec[2], seed, tilt=spec[3])
    return bundle_point(ns, n, spec[1], spec[2], seed)
