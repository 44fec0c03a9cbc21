"""Where does a C3 optimisation step spend its time?  (run on the GPU box; prints a kernel table)"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import bench
import raytracetorch_b200 as rtt

dev = torch.device("cuda", 0)
w = bench.build_workload("c3", dev)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10 ** 7
scene = rtt.scene.SequentialScene(w["elements"]).to(dev)
bundle = rtt.rays.CollimatedDisk(5.0, 0, device=dev, transform=rtt.geom.RayTransformBundle(
    translation=[0.0, 0.0, -10.0]).to(dev))
goal = rtt.optim.SpotSizeLoss(w["sensor"], [bundle], N_rays=n)
params = [p for p in scene.parameters() if p.requires_grad]
opt = torch.optim.Adam(params, lr=1e-5)


def step():
    opt.zero_grad(set_to_none=True)
    loss = goal(scene)
    loss.backward()
    opt.step()
    return loss


for _ in range(3):
    step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5):
    step()
torch.cuda.synchronize()
print("ms/step", 1e3 * (time.perf_counter() - t0) / 5)
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=40, max_name_column_width=70))
