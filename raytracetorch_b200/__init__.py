"""raytracetorch_b200 — B200-native (sm_100a) kernels for RayTraceTorch's batched
ray-propagation hot path, behind the reference's Element / Scene / Bundle / Rays API.

    import raytracetorch_b200 as rtt
    lens = rtt.elements.SingletLens(...); sensor = rtt.elements.Sensor(rtt.geom.Disk(...))
    scene = rtt.scene.SequentialScene([lens, sensor]).cuda()
    rays = rtt.rays.CollimatedDisk(5.0, 0, device="cuda").sample(10**7)
    scene.simulate(rays)            # one fused CUDA kernel; backward = one adjoint kernel

Sub-modules mirror the reference's packages: ``geom``, ``phys``, ``elements``, ``rays``,
``scene``, ``optim``, ``render``.  ``table`` is the scene compiler, ``ops`` the torch custom
ops over the C ABI (``include/rtt_b200.h``), ``dist`` the multi-GPU sharding helpers.
There is no CPU compute path.
"""
from . import codes, geom, phys, elements, rays, table, scene  # noqa: F401
from .table import Dispersion, compile_elements  # noqa: F401

__version__ = "0.1.0"


def __getattr__(name):
    # heavier / optional sub-modules are imported on first use
    if name in ("ops", "dist", "optim", "render", "sources"):
        import importlib
        return importlib.import_module(f"{__name__}.{name}")
    raise AttributeError(name)
