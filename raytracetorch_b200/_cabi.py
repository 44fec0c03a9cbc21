"""ctypes binding of the C ABI in ``include/rtt_b200.h``.

``load()`` opens ``raytracetorch_b200/librtt_b200.so`` — the CUDA library, built in-tree by
``csrc/build.sh`` — and nothing else.  There is no CPU path: if the library is missing the
import of any compute op raises ``RttLibraryMissing`` with the build command.

All functions take raw addresses (``tensor.data_ptr()``); ``0`` is the C ``NULL``.
"""
from __future__ import annotations

import ctypes as ct
import os
from typing import Optional, Sequence

_HERE = os.path.dirname(os.path.abspath(__file__))
# RTT_B200_LIB: an alternative build of librtt_b200.so (A/B experiments, scripts/gpu_ab_lib.sh); default = the in-tree library
LIB_PATH = os.environ.get("RTT_B200_LIB") or os.path.join(_HERE, "librtt_b200.so")

MODE_FAST, MODE_EXACT = 0, 1
MODE_SCALAR_GRADS = 0x100   # hint for rtt_trace_seq_bwd (include/rtt_b200.h): no row requests pose gradients
MODE_NONSEQ_FAST = 0x400    # rtt_trace_nonseq_fwd / _bwd: explicit opt-in to the FAST arithmetic


class RttLibraryMissing(RuntimeError):
    pass


class RttError(RuntimeError):
    pass


class SensorReq(ct.Structure):
    _fields_ = [("image", ct.c_void_p), ("record", ct.c_void_p),
                ("height", ct.c_int32), ("width", ct.c_int32), ("channels", ct.c_int32),
                ("x0", ct.c_float), ("y0", ct.c_float), ("sx", ct.c_float), ("sy", ct.c_float),
                ("record_hits", ct.c_int32), ("count", ct.c_void_p)]


class SourceReq(ct.Structure):
    """rtt_source_t"""
    _fields_ = [("kind", ct.c_int32), ("a", ct.c_float * 4), ("width", ct.c_int32), ("height", ct.c_int32),
                ("pose", ct.c_void_p), ("seed", ct.c_uint64), ("first", ct.c_int64), ("state", ct.c_void_p),
                ("intensity", ct.c_float), ("wavelength", ct.c_float)]


SRC_DISK, SRC_LINE, SRC_FAN, SRC_POINT, SRC_CAMERA = 0, 1, 2, 3, 4
SPOT_ID_WORK = 296 * 256 * 4 + 4     # floats of scratch per stream for the per-id reductions (RTT_SPOT_ID_WORK)
SPOT_WORK = 4 * 1024 + 4


def make_source(kind: int, a: Sequence[float], pose_ptr: int, *, seed: int = 0, first: int = 0, state_ptr: int = 0,
                width: int = 0, height: int = 0, intensity: float = 1.0, wavelength: float = 0.0) -> SourceReq:
    a = list(a) + [0.0] * (4 - len(a))
    return SourceReq(kind, (ct.c_float * 4)(*a), width, height, pose_ptr or None, seed & (2 ** 64 - 1), first,
                     state_ptr or None, intensity, wavelength)


class TableReq(ct.Structure):
    _fields_ = [("f", ct.c_void_p), ("i", ct.c_void_p), ("n_rows", ct.c_int32), ("n_lut", ct.c_int32),
                ("lut", ct.c_void_p), ("lut_w", ct.c_void_p)]


_P = ct.c_void_p
_SRC = ct.POINTER(SourceReq)
_SIGS = {
    "rtt_trace_seq_fwd": [_P, _P, _P, _P, _SRC, _P, _P, _P, _P, ct.POINTER(TableReq), ct.POINTER(SensorReq),
                          ct.c_int32, ct.c_int64, ct.c_int32, _P],
    "rtt_trace_seq_bwd": [_P, _P, _P, _P, _SRC, _P, _P, _P, _P, ct.POINTER(_P), _P, _P, _P, _P, _P,
                          ct.POINTER(TableReq), ct.c_int32, ct.c_int64, ct.c_int32, _P],
    "rtt_trace_nonseq_fwd": [_P, _P, _P, _P, _SRC, _P, _P, _P, _P, _P, ct.POINTER(TableReq), ct.POINTER(SensorReq),
                             ct.c_int32, ct.c_int32, ct.c_int64, ct.c_int32, _P],
    "rtt_trace_nonseq_bwd": [_P, _P, _P, _P, _SRC, _P, ct.c_int32, _P, _P, _P, ct.POINTER(_P), ct.POINTER(ct.c_int32),
                             _P, _P, _P, _P, _P,
                             ct.POINTER(TableReq), ct.c_int32, ct.c_int64, ct.c_int32, _P],
    "rtt_intersect_test": [_P, _P, _P, ct.POINTER(TableReq), ct.c_int32, ct.c_int32, ct.c_int64, ct.c_int32, _P],
    "rtt_surface_step_fwd": [_P, _P, _P, _P, _P, _P, _P, _P, _P, ct.POINTER(TableReq), ct.c_int32,
                             ct.c_int64, ct.c_int32, _P],
    "rtt_surface_step_bwd": [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, ct.POINTER(TableReq), ct.c_int32,
                             ct.c_int64, ct.c_int32, _P],
    "rtt_sample_bundle": [_SRC, _P, _P, _P, _P, ct.c_int64, ct.c_int32, _P],
    "rtt_spot_moments": [_P, ct.c_int64, ct.c_int32, _P, _P, _P],
    "rtt_spot_moments_bwd": [_P, ct.c_int64, ct.c_int32, _P, _P, _P],
    "rtt_spot_size_fwd": [_P, ct.c_int64, _P, _P, _P, _P, _P],
    "rtt_spot_size_bwd": [_P, ct.c_int64, _P, _P, _P, _P, _P, _P],
    "rtt_render_shade": [_P, _P, _SRC, ct.POINTER(TableReq), _P, ct.POINTER(ct.c_float), ct.POINTER(ct.c_float), _P, _P,
                         ct.c_int64, ct.c_int32, _P],
    "rtt_spot_id_moments": [_P, _P, ct.c_int64, _P, ct.c_int32, _P, _P, _P],
    "rtt_spot_id_size": [_P, _P, ct.c_int64, _P, ct.c_int32, _P, ct.c_float, _P, _P, _P],
    "rtt_spot_id_size_bwd": [_P, _P, ct.c_int64, _P, ct.c_int32, _P, ct.c_float, _P, _P],
}
# symbols every build of the CUDA library must export (tests/test_cabi.py checks them)
EXPORTS = tuple(_SIGS) + ("rtt_version", "rtt_layout_query", "rtt_error_string", "rtt_launch_count")


def make_table(f_ptr: int, i_ptr: int, n_rows: int, lut_ptr: int = 0, lut_w_ptr: int = 0, n_lut: int = 0) -> TableReq:
    return TableReq(f_ptr, i_ptr, n_rows, n_lut, lut_ptr or None, lut_w_ptr or None)


def make_sensors(reqs: Sequence[dict]):
    """reqs: dicts with image, record, count (addresses or 0), height, width, channels, x0, y0, sx, sy,
    record_hits."""
    if not reqs:
        return None, 0
    arr = (SensorReq * len(reqs))()
    for k, r in enumerate(reqs):
        arr[k] = SensorReq(r.get("image") or None, r.get("record") or None,
                           r.get("height", 0), r.get("width", 0), r.get("channels", 1),
                           r.get("x0", 0.0), r.get("y0", 0.0), r.get("sx", 0.0), r.get("sy", 0.0),
                           r.get("record_hits", 1), r.get("count") or None)
    return arr, len(reqs)


class RttLib:
    """Typed handle on one shared library exporting the rtt_* symbols."""

    def __init__(self, path: str):
        self.path = path
        self.dll = ct.CDLL(path)
        for name, sig in _SIGS.items():
            fn = getattr(self.dll, name)
            fn.argtypes, fn.restype = sig, ct.c_int
        if hasattr(self.dll, "rtt_error_string"):
            self.dll.rtt_error_string.argtypes, self.dll.rtt_error_string.restype = [ct.c_int], ct.c_char_p
            self.dll.rtt_layout_query.argtypes, self.dll.rtt_layout_query.restype = [ct.c_int], ct.c_int
            self.dll.rtt_launch_count.argtypes, self.dll.rtt_launch_count.restype = [], ct.c_int64
            self.dll.rtt_version.argtypes, self.dll.rtt_version.restype = [], ct.c_int
        if hasattr(self.dll, "rtt_probe_fp32"):
            self.dll.rtt_probe_fp32.argtypes = [ct.c_int32, ct.c_void_p, ct.c_void_p]
            self.dll.rtt_probe_fp32.restype = ct.c_int64

    def check(self, code: int, what: str):
        if code != 0:
            msg = self.dll.rtt_error_string(code).decode() if hasattr(self.dll, "rtt_error_string") else str(code)
            raise RttError(f"{what} failed: {msg} (code {code})")

    def call(self, name: str, *args):
        self.check(getattr(self.dll, name)(*args), name)

    def launch_count(self) -> int:
        return int(self.dll.rtt_launch_count())

    def layout(self, which: int) -> int:
        return int(self.dll.rtt_layout_query(which))


_lib: Optional[RttLib] = None


def load() -> RttLib:
    """The CUDA library (and only it).  Raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RttLibraryMissing(
                f"{LIB_PATH} not found. Build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                f"or `bash raytracetorch_b200/csrc/build.sh`. This package has no CPU path.")
        lib = RttLib(LIB_PATH)
        from . import codes as C
        want = (C.ROW_F, C.ROW_I, C.ROW_G, C.MAX_ROWS, C.N_DIFF, C.MAX_SENSORS, C.MAX_WAVELENGTHS, C.MAX_BOUNCES)
        got = tuple(lib.layout(k) for k in range(8))
        if got != want:
            raise RttError(f"librtt_b200.so layout {got} does not match codes.py {want}; rebuild the library")
        _lib = lib
    return _lib
