"""Adjoint kernel alone (CUDA events) and a kernel table of the forward+adjoint step (run on the GPU box).

    python scripts/bwd_breakdown.py [workload=c2] [rays=1e8] [--profile]
"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import bench
import raytracetorch_b200 as rtt

wl = sys.argv[1] if len(sys.argv) > 1 else "c2"
n = int(float(sys.argv[2])) if len(sys.argv) > 2 else 10 ** 8
dev = torch.device("cuda", 0)
w = bench.build_workload(wl, dev)
scene = rtt.scene.SequentialScene(w["elements"])
scene.set_dispersion(w["dispersion"])
scene = scene.to(dev)
scene.record_hits = False
for p in scene.parameters():
    p.requires_grad_(False)
params = []
for el in w["elements"]:
    for s in getattr(el.shape, "surfaces", []):
        if hasattr(s, "c") and isinstance(s.c, torch.nn.Parameter):
            s.c.requires_grad_(True)
            params.append(s.c)
pos, dirs, inten, wav = bench.synth_bundle(w, n, dev, 1000)
tab = scene.table()
mode = rtt.ops.get_default_mode()
fwd = torch.ops.rtt_b200.trace_seq_fwd(pos, dirs, inten, wav, tab.f.detach(), tab.i, tab.lut, tab.lut_wavelengths,
                                       [], False, mode)
opos, odir, oint, hitmask = fwd[:4]
g_pos = torch.zeros_like(opos)
g_pos[:, :2] = 2.0 * oint[:, None] * opos[:, :2]
g_int = (opos[:, :2] ** 2).sum(1)


def bwd():
    return torch.ops.rtt_b200.trace_seq_bwd(pos, dirs, inten, wav, hitmask, g_pos, None, g_int, None, tab.f.detach(),
                                            tab.i, tab.lut, tab.lut_wavelengths, False, True,
                                            mode | rtt.ops.adjoint_hint(tab))


ts = []
for k in range(8):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); out = bwd(); b.record()
    torch.cuda.synchronize()
    ts.append(a.elapsed_time(b))
gt = out[3]
print(f"{wl} n={n} adjoint kernel ms: median {np.median(ts[2:]):.3f} min {min(ts):.3f}  |g_table| {float(gt.abs().sum()):.6e}")

if "--goal" in sys.argv:
    ids = torch.zeros(n, dtype=torch.int8, device=dev)
    wav_r = wav if wav is not None else torch.zeros(n, device=dev)

    class ResidentBundle(rtt.rays.Bundle):
        def sample(self, N):
            return rtt.rays.Rays._wrap(pos=pos, dir=dirs, intensity=inten, id=ids, wavelength=wav_r)

    goal = rtt.optim.SpotSizeLoss(w["sensor"], [ResidentBundle(0, device=dev)], N_rays=n, target_xy=torch.zeros(2))

    def gstep():
        for p in params:
            p.grad = None
        goal(scene).backward()
    for _ in range(3):
        gstep()
    torch.cuda.synchronize()
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(3):
            gstep()
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=22, max_name_column_width=90))

if "--profile" in sys.argv:
    def step():
        for p in params:
            p.grad = None
        t = scene.table()
        o = rtt.ops.trace_sequential(t, pos, dirs, inten, wav, want_record=False, sensor_cfg=[])
        xy = o["pos"][:, :2]
        loss = torch.dot(o["intensity"], (xy * xy).sum(1))
        loss.backward()
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        for _ in range(3):
            step()
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=80))
