"""pytest configuration: markers, import paths, shared fixtures.

``-m "not gpu"`` : oracle vs golden fixtures, scene compiler, kernel arithmetic on the host
                   (tests/hostsim), C-ABI symbol checks, gloo world_size-2 sharding.
``-m gpu``       : the parity tests proper — CUDA kernels through the C ABI vs the oracle.
"""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "slow: long-running full-size case")


@pytest.fixture(scope="session")
def rtt_ns():
    """This repo's package as the ``ns`` namespace expected by tests/scenes.py."""
    import types
    import raytracetorch_b200 as rtt
    return types.SimpleNamespace(elements=rtt.elements, geom=rtt.geom, phys=rtt.phys, rays=rtt.rays,
                                 scene=rtt.scene)


# Every kernel-parity test runs on two back-ends that share the per-ray source
# (csrc/rtt_core.cuh): "host" = tests/hostsim (g++ build, CPU suite) and "gpu" = the CUDA
# library through its C ABI (tests/gpusim, `-m gpu` suite).
_BACKENDS = ["host", pytest.param("gpu", marks=pytest.mark.gpu)]


def _runner(backend, variant):
    if backend == "host":
        from hostsim import build
        return build({"pair_plain": "pair", "tile": "fast", "nonseq_fast": "fast"}.get(variant, variant))
    import torch
    if not torch.cuda.is_available():
        pytest.fail("gpu test selected but no CUDA device is visible (there is no CPU fallback)")
    from gpusim import GpuSim
    # mode word of the C ABI: arithmetic in bits 0..7 (0 FAST, 1 EXACT), kernel build in bits 16..23
    # (include/rtt_b200.h RTT_MODE_TUNE_*): "pair" = packed ray pairs with bulk-async ray streaming, "pair_plain" =
    # the same arithmetic with plain loads / stores, "tile" = the scalar frame-resident tile kernel
    # "nonseq_fast": the explicit opt-in of the non-sequential entries to the FAST arithmetic (RTT_MODE_NONSEQ_FAST)
    # "fast" = the default build (tile kernel, one persistent 1024-thread block per SM); "tile" = the same kernel in
    # 256-thread blocks
    return GpuSim({"exact": 1, "fast": 0, "pair": 16 << 16, "pair_plain": 18 << 16, "tile": 3 << 16,
                   "nonseq_fast": 0x400}[variant])


@pytest.fixture(params=_BACKENDS)
def run_exact(request):
    return _runner(request.param, "exact")


@pytest.fixture(params=_BACKENDS)
def run_fast(request):
    return _runner(request.param, "fast")


@pytest.fixture(params=_BACKENDS)
def runner_of(request):
    return lambda variant: _runner(request.param, variant)
