import os, sys, time, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gi-gs_b200"))
from gigs import scene, shade, step as gstep
dev = torch.device("cuda:0")
raw = scene.make_scene(300000, seed=0, regime="trained")
params = gstep.GaussianParams(raw, dev, light=scene.make_light(0))
light = params.light(); lut = shade.make_brdf_lut().to(dev)
cam = scene.orbit_camera(0, 8, 800, 800).to(dev)
rays = scene.canonical_rays(cam, dev)
gt = torch.rand(3, 800, 800, device=dev); bg = torch.zeros(3, device=dev)
gi = dict(radius=0.8, bias=0.01, thick=0.05, delta=0.0625, step=16, start=64)
for i in range(5):
    params.zero_grad(); gstep.training_step(params, cam, light, lut, rays, gt, bg, gi)
torch.cuda.synchronize()
for rep in range(3):
    t0 = time.perf_counter()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(10):
        params.zero_grad(); gstep.training_step(params, cam, light, lut, rays, gt, bg, gi)
    e1.record()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"cpu issue {1e2*(t1-t0):.3f} ms/step, wall {1e2*(t2-t0):.3f} ms/step, gpu events {e0.elapsed_time(e1)/10:.3f} ms/step")
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for i in range(3):
        params.zero_grad(); gstep.training_step(params, cam, light, lut, rays, gt, bg, gi)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=40, max_name_column_width=60))
