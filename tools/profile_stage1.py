"""A few first-stage iterations (fused first-stage frame + one-launch Adam) at the headline scene size, for ncu:
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python tools/profile_stage1.py
  ncu --set full --clock-control none --import-source on -k regex:'adam|ssim|hybrid|gaussian_backward|normal_loss|stage1_normals' \
      -s 28 -c 14 -o gpurun_out/stage1 python tools/profile_stage1.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gi-gs_b200"))
import torch

from gigs import optim as gopt, scene, step as gstep

GI = dict(radius=0.8, bias=0.01, thick=0.05, delta=0.0625, step=16, start=64)


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 5
    dev = torch.device("cuda:0")
    raw = scene.make_scene(300000, seed=0, regime="trained")
    cams = [scene.orbit_camera(k, 8, 800, 800).to(dev) for k in range(8)]
    g = torch.Generator().manual_seed(7)
    gts = [torch.rand(3, 800, 800, generator=g).to(dev) for _ in range(8)]
    bg = torch.zeros(3, device=dev)
    params = gstep.GaussianParams(raw, dev)
    opt = gopt.GaussianOptimizer(params)
    for i in range(n):
        loss, _ = gstep.first_stage_step(params, cams[i % 8], gts[i % 8], bg, GI, fused=True)
        opt.step(light=False)
    torch.cuda.synchronize()
    print("first-stage loss", float(loss))


if __name__ == "__main__":
    main()
