"""Pin the CPU oracle and the scene compiler to the UNMODIFIED reference.

The fixtures under tests/golden/ were produced by oracle/make_golden.py running the real
reference (SequentialScene.simulate / Scene.step / torch autograd) in the build container.
These tests need neither the reference nor a GPU.
"""
import types

import numpy as np
import pytest
import torch

import parity
import scenes
from oracle import trace_oracle as O


def _run_oracle(d):
    tf, ti = torch.from_numpy(d["table_f"]), d["table_i"].tolist()
    p, dd, inten = parity.inputs_t(d)
    if str(d["mode"]) == "seq":
        return O.trace_sequential(tf, ti, p, dd, inten)
    return O.trace_nonsequential(tf, ti, p, dd, inten, int(d["nbounces"]))


@pytest.mark.parametrize("name", parity.golden_names())
def test_oracle_reproduces_reference_bit_for_bit(name):
    """fp32 oracle == fp32 reference, every ray, every bit (13 scenes: all surface kinds,
    bounds, physics, tilted poses, sequential and non-sequential)."""
    d = parity.load(name)
    o = _run_oracle(d)
    for k in ("pos", "dir", "intensity"):
        np.testing.assert_array_equal(o[k].numpy(), d[f"f32_{k}"], err_msg=f"{name}:{k}")
    if str(d["mode"]) == "nonseq":
        np.testing.assert_array_equal(o["seq"].numpy(), d["f32_seq"])


@pytest.mark.parametrize("name", parity.forward_names("seq"))
def test_oracle_sensor_records_match_reference(name):
    """Sensor hit lists (hit_local, intensity BEFORE the sensor: elements/sensor.py:35-37)."""
    d = parity.load(name)
    o = _run_oracle(d)
    if "f32_sensor0_loc" not in d.files:
        pytest.skip("scene has no sensor")
    mask, hl, w = o["sensor"][0]
    np.testing.assert_array_equal(hl.numpy(), d["f32_sensor0_loc"])
    np.testing.assert_array_equal(w.numpy(), d["f32_sensor0_w"])


@pytest.mark.parametrize("name", parity.forward_names("seq"))
def test_oracle_fp64_tracks_reference_fp64(name):
    """Same oracle in double == the reference in double to rounding: the restatement has no
    fp32-specific accident in it."""
    d = parity.load(name)
    tf = torch.from_numpy(d["table_f"]).double()
    p, dd, inten = (t.double() for t in parity.inputs_t(d))
    o = O.trace_sequential(tf, d["table_i"].tolist(), p, dd, inten)
    live = d["f64_intensity"] > 0
    # the table was built in fp32, the reference's fp64 run re-derived its poses in fp64:
    # agreement is therefore at fp32-parameter level, not 1e-15
    agree = (o["intensity"].numpy() > 0) == live
    assert agree.mean() > 0.995
    sel = live & agree
    assert parity.vec_rel(o["pos"].numpy()[sel], d["f64_pos"][sel]).max() < 5e-4
    assert np.median(parity.vec_rel(o["pos"].numpy()[sel], d["f64_pos"][sel])) < 1e-6


def _mirror_case(rtt_ns, name):
    if name in scenes.CASES:
        builder, kw, _mode, _b = scenes.CASES[name]
    else:
        builder, kw, _b = scenes.GRAD_CASES[name]
    return builder(rtt_ns, **kw)


@pytest.mark.parametrize("name", parity.golden_names())
def test_mirror_classes_compile_to_reference_table(rtt_ns, name):
    """This repo's Element/Shape/Surface classes flatten to the very table the scene compiler
    produced from the reference's own objects (row order scene/base.py:116-123)."""
    import raytracetorch_b200 as rtt
    d = parity.load(name)
    tab = rtt.compile_elements(_mirror_case(rtt_ns, name))
    np.testing.assert_array_equal(tab.i.numpy()[:, :10], d["table_i"][:, :10])
    np.testing.assert_array_equal(tab.f.detach().numpy(), d["table_f"])


@pytest.mark.parametrize("name", parity.golden_names(grads=True))
def test_oracle_autograd_matches_reference_autograd(rtt_ns, name):
    """d loss / d (every trainable Parameter, input pos/dir/intensity): table built from the
    mirror classes + oracle autograd vs the reference's own backward pass."""
    import raytracetorch_b200 as rtt
    builder, kw, _ = scenes.GRAD_CASES[name]
    d = parity.load(name)
    els = builder(rtt_ns, **kw)
    holder = torch.nn.Module()
    holder.elements = torch.nn.ModuleList(els)
    tab = rtt.compile_elements(els)
    p, dd, inten = parity.inputs_t(d)
    for t in (p, dd, inten):
        t.requires_grad_(True)
    o = O.trace_sequential(tab.f, tab.i_host, p, dd, inten)
    loss = parity.golden_loss(o["pos"], o["dir"], o["intensity"])
    loss.backward()
    assert abs(float(loss.detach()) - float(d["f32_loss"])) <= 1e-6 * abs(float(d["f32_loss"]))
    assert parity.grad_rel(p.grad.numpy(), d["f32_g_pos"]) < 1e-6
    assert parity.grad_rel(dd.grad.numpy(), d["f32_g_dir"]) < 1e-6
    assert parity.grad_rel(inten.grad.numpy(), d["f32_g_intensity"]) < 1e-6
    params = dict(holder.named_parameters())
    gold = [k[len("f32_gp::"):] for k in d.files if k.startswith("f32_gp::")]
    assert gold, "fixture holds no parameter gradients"
    for k in gold:
        assert k in params, f"mirror classes lack parameter {k}"
        g = params[k].grad
        g = np.zeros_like(d["f32_gp::" + k]) if g is None else g.numpy()
        ref = d["f32_gp::" + k]
        if np.linalg.norm(ref) == 0:
            assert np.linalg.norm(g) == 0
        else:
            # 5e-5, or the reference's own fp32-vs-fp64 distance where its fp32 gradient is that noisy
            # (summation-order noise of an ill-conditioned rot_vec gradient, x3_ideal)
            noise = parity.grad_rel(ref, d["f64_gp::" + k].astype(np.float32))
            assert parity.grad_rel(g, ref) < max(5e-5, 2.0 * noise), k
