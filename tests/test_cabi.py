"""The C-ABI library: loads, exports every symbol include/rtt_b200.h declares, layout handshake,
argument validation — and NO CPU path (compute entries fail loudly without a device)."""
import ctypes as ct
import os
import re

import numpy as np
import pytest
import torch

from raytracetorch_b200 import _cabi, codes as C

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "rtt_b200.h")


def _declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(rtt_[a-z_0-9]+)\s*\(", src)))


def test_header_declares_the_expected_entry_points():
    names = _declared_symbols()
    for want in ("rtt_trace_seq_fwd", "rtt_trace_seq_bwd", "rtt_trace_nonseq_fwd", "rtt_trace_nonseq_bwd",
                 "rtt_intersect_test", "rtt_surface_step_fwd", "rtt_surface_step_bwd", "rtt_version",
                 "rtt_layout_query", "rtt_error_string", "rtt_launch_count"):
        assert want in names


def test_library_exports_every_declared_symbol():
    assert os.path.exists(_cabi.LIB_PATH), "build with python -c 'import __graft_entry__ as g; g.build()'"
    dll = ct.CDLL(_cabi.LIB_PATH)
    for name in _declared_symbols():
        assert hasattr(dll, name), f"{name} declared in include/rtt_b200.h but not exported"
    for name in _cabi.EXPORTS:
        assert name in _declared_symbols(), f"{name} bound by _cabi.py but not declared in the header"


def test_layout_handshake_matches_codes_py():
    lib = _cabi.load()
    want = (C.ROW_F, C.ROW_I, C.ROW_G, C.MAX_ROWS, C.N_DIFF, C.MAX_SENSORS, C.MAX_WAVELENGTHS, C.MAX_BOUNCES)
    assert tuple(lib.layout(k) for k in range(8)) == want
    assert lib.layout(99) == -1
    assert lib.dll.rtt_version() >= 100


def test_header_enums_match_codes_py():
    src = open(HEADER).read()
    vals = {}
    for body in re.findall(r"enum\s*\{(.*?)\}", src, flags=re.S):
        nxt = 0
        for item in body.split(","):
            item = item.strip()
            if not item:
                continue
            if "=" in item:
                k, v = (x.strip() for x in item.split("="))
                nxt = int(eval(v, {}, dict(vals)))
            else:
                k = item
            vals[k] = nxt
            nxt += 1
    for k, v in vals.items():
        if not k.startswith("RTT_") or k.startswith(("RTT_E_", "RTT_OK", "RTT_MODE_")):
            continue
        py = k[len("RTT_"):]
        assert hasattr(C, py), f"codes.py lacks {py}"
        assert getattr(C, py) == v, f"{k}: header {v} != codes.py {getattr(C, py)}"


def test_argument_validation_and_no_cpu_path():
    lib = _cabi.load()
    tf = np.zeros((1, C.ROW_F), np.float32)
    ti = np.zeros((1, C.ROW_I), np.int32)
    pos = np.zeros((4, 3), np.float32)
    inten = np.ones(4, np.float32)
    req = _cabi.make_table(tf.ctypes.data, ti.ctypes.data, 1)
    f = lib.dll.rtt_trace_seq_fwd
    # null table / null rays / too many rows: negative RTT_E_* codes, never a crash
    assert f(pos.ctypes.data, pos.ctypes.data, inten.ctypes.data, None, None, pos.ctypes.data, pos.ctypes.data,
             inten.ctypes.data, None, None, None, 0, 4, 0, None) == -1
    assert f(None, pos.ctypes.data, inten.ctypes.data, None, None, pos.ctypes.data, pos.ctypes.data,
             inten.ctypes.data, None, ct.byref(req), None, 0, 4, 0, None) == -1
    big = _cabi.make_table(tf.ctypes.data, ti.ctypes.data, C.MAX_ROWS + 1)
    assert f(pos.ctypes.data, pos.ctypes.data, inten.ctypes.data, None, None, pos.ctypes.data, pos.ctypes.data,
             inten.ctypes.data, None, ct.byref(big), None, 0, 4, 0, None) == -2
    assert b"rows" in lib.dll.rtt_error_string(-2)
    if not torch.cuda.is_available():
        # valid arguments, no device: must refuse (RTT_E_NO_DEVICE), not compute on the host
        code = f(pos.ctypes.data, pos.ctypes.data, inten.ctypes.data, None, None, pos.ctypes.data, pos.ctypes.data,
                 inten.ctypes.data, None, ct.byref(req), None, 0, 4, 0, None)
        assert code == -3
        assert b"no CPU path" in lib.dll.rtt_error_string(code)
        with pytest.raises(_cabi.RttError):
            lib.call("rtt_trace_seq_fwd", pos.ctypes.data, pos.ctypes.data, inten.ctypes.data, None, None,
                     pos.ctypes.data, pos.ctypes.data, inten.ctypes.data, None, ct.byref(req), None, 0, 4, 0, None)


def test_python_ops_refuse_cpu_tensors(rtt_ns):
    import raytracetorch_b200 as rtt
    import scenes
    els = scenes.c1_singlet(rtt_ns, physical=True)
    scene = rtt.scene.SequentialScene(els)
    rays = scenes.make_bundle(rtt_ns, ("coll", 5.0, -10.0, None), 16, 0)
    with pytest.raises(rtt.ops.NoCpuPathError):
        scene.simulate(rays)


def test_missing_library_fails_loudly(monkeypatch):
    monkeypatch.setattr(_cabi, "_lib", None)
    monkeypatch.setattr(_cabi, "LIB_PATH", "/nonexistent/librtt_b200.so")
    with pytest.raises(_cabi.RttLibraryMissing):
        _cabi.load()
