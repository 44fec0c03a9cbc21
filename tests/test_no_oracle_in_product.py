"""The oracle is test infrastructure: the product package must never import, call or link it,
and must not carry a CPU fallback of its own."""
import ast
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "raytracetorch_b200")


def _py_files():
    for base, _dirs, files in os.walk(PKG):
        for f in files:
            if f.endswith(".py"):
                yield os.path.join(base, f)


def test_product_sources_never_mention_the_oracle_or_hostsim():
    for path in _py_files():
        tree = ast.parse(open(path).read())
        for node in ast.walk(tree):
            mods = []
            if isinstance(node, ast.Import):
                mods = [a.name for a in node.names]
            elif isinstance(node, ast.ImportFrom):
                mods = [node.module or ""]
            for m in mods:
                assert not m.split(".")[0] in ("oracle", "hostsim", "tests"), f"{path} imports {m}"
    for base, _dirs, files in os.walk(os.path.join(PKG, "csrc")):
        for f in files:
            if f.endswith((".cu", ".cuh", ".h", ".inl", ".sh")):
                txt = open(os.path.join(base, f)).read()
                assert "oracle/" not in txt and "#include \"../../tests" not in txt, f


def test_importing_the_product_does_not_load_the_oracle():
    code = ("import sys, raytracetorch_b200 as r, raytracetorch_b200.ops, raytracetorch_b200.scene;"
            "bad=[m for m in sys.modules if m.split('.')[0] in ('oracle','hostsim')];"
            "assert not bad, bad; print('ok')")
    out = subprocess.run([sys.executable, "-c", code], cwd=ROOT, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    assert "ok" in out.stdout


def test_cabi_binding_opens_only_the_cuda_library():
    src = open(os.path.join(PKG, "_cabi.py")).read()
    assert src.count("CDLL(") == 1 and "librtt_b200.so" in src
    assert "hostsim" not in src.replace("tests/hostsim", "")
