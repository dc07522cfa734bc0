#!/usr/bin/env python
"""Timings of the BASELINE.json configs that bench.py does not time (they are parity-test cases; these numbers are
for DESIGN.md): C1 100k forward + GI, C3 bicycle-shaped 6M fwd+bwd, C4 K-view step, C5 relight/eval sweep.
One process per GPU under torch.distributed.run for C4/C5 with N > 1. CUDA-event timing, max over ranks.
usage: python tools/bench_configs.py [--only c1,c3,c4,c5] > gpurun_out/configs.json"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gi-gs_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from gigs import frame as gframe, renderer, scene, shade, step as gstep  # noqa: E402

GI = dict(radius=0.8, bias=0.01, thick=0.05, delta=0.0625, step=16, start=64)


def ev_ms(fn, reps, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="c1,c3,c4,c5")
    args = ap.parse_args()
    only = set(args.only.split(","))
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    out = {"n_gpus": world}

    def maxr(x):
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    lut = shade.make_brdf_lut().to(dev)
    bg = torch.zeros(3, device=dev)

    if "c1" in only and world == 1:
        # C1: 100k Gaussians, one 800x800 camera, G-buffer forward + screen-space indirect (step 16, start 64)
        raw = scene.make_scene(100000, seed=0); g = scene.activate(raw, dev)
        lh = scene.make_light(0); light = shade.Light([s.to(dev) for s in lh["specular"]], lh["diffuse"].to(dev))
        cam = scene.orbit_camera(0, 8, 800, 800).to(dev); rays = scene.canonical_rays(cam, dev)
        for start in (64, 8):
            gi = dict(GI, start=start)
            ms_f = ev_ms(lambda: gframe.pbr_frame_eval(g, cam, light, lut, rays, bg, gi, inference=False), 10)
            with torch.no_grad():
                ms_o = ev_ms(lambda: renderer.pbr_forward(cam, g, light, lut, rays, bg, gi=gi), 5)
            out[f"c1_forward_gi_start{start}"] = {"fused_frame_ms": ms_f, "operator_path_ms": ms_o}

    if "c3" in only and world == 1:
        # C3: bicycle-shaped 6M Gaussians, degree 3, --metallic, 1237x822, fwd+bwd (PBR-stage frame)
        P, W, H = 6_000_000, 1237, 822
        raw = scene.make_scene(P, seed=0, regime="trained", shape="bicycle")
        params = gstep.GaussianParams(raw, dev, light=scene.make_light(0))
        del raw
        cam = scene.look_at_camera([4.0, 0.0, 1.0], [0.0, 0.0, 0.0], W, H, fx=1040.0).to(dev)
        rays = scene.canonical_rays(cam, dev)
        gt = torch.rand(3, H, W, device=dev)

        def step():
            params.zero_grad(fused_only=True)
            gstep.training_step(params, cam, params.light(), lut, rays, gt, bg, GI)
        ms = ev_ms(step, 5)
        out["c3_bicycle_6M_fwd_bwd"] = {"ms_per_frame": ms, "frames_per_s": 1e3 / ms,
                                        "num_rendered": params.last_workspace.num_rendered}
        del params
        torch.cuda.empty_cache()
        # the same scene through the FIRST-stage frame: every parameter group receives gradients (general blend backward,
        # per-Gaussian backward through the getters), + the one-launch Adam over all 6M x 67 parameters
        from gigs import optim as gopt
        raw = scene.make_scene(P, seed=0, regime="trained", shape="bicycle")
        params = gstep.GaussianParams(raw, dev)
        del raw
        opt = gopt.GaussianOptimizer(params)

        def step1():
            gstep.first_stage_step(params, cam, gt, bg, GI, fused=True)
        params.zero_grad()
        ms1 = ev_ms(step1, 5)

        def step1o():
            gstep.first_stage_step(params, cam, gt, bg, GI, fused=True)
            opt.step(light=False)
        ms1o = ev_ms(step1o, 5)
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        params.mark_dirty(None)
        e0.record(); opt.step(light=False); e1.record()
        torch.cuda.synchronize()
        out["c3_bicycle_6M_first_stage"] = {"frame_fwd_bwd_ms": ms1, "iteration_with_adam_ms": ms1o,
                                            "adam_ms": e0.elapsed_time(e1),
                                            "adam_GBps": 32.0 * 67 * P / (e0.elapsed_time(e1) * 1e-3) / 1e9}
        del params, opt
        torch.cuda.empty_cache()

    if "c4" in only:
        # C4: K cameras per step sharded by view, replicated Gaussians, gradient all-reduce over NVLink
        raw = scene.make_scene(300000, seed=0)
        params = gstep.GaussianParams(raw, dev, light=scene.make_light(0))
        for K in (8, 32):
            cams = [scene.orbit_camera(k, K, 800, 800).to(dev) for k in range(K)]
            rays = scene.canonical_rays(cams[0], dev)
            gen = torch.Generator().manual_seed(1)
            gts = [torch.rand(3, 800, 800, generator=gen).to(dev) for _ in range(K)]

            def step():
                gstep.multi_view_step(params, cams, params.light(), lut, lambda c: rays, gts, bg, GI, rank=rank,
                                      world=world)
            ms = maxr(ev_ms(step, 3, warm=1))
            out[f"c4_multi_view_K{K}"] = {"ms_per_step": ms, "views_per_s": K * 1e3 / ms}
            del cams, gts

    if "c5" in only:
        # C5: relight / eval sweep, 200 views x one 2k environment map, forward only, camera sharded
        raw = scene.make_scene(300000, seed=0); g = scene.activate(raw, dev)
        lh = scene.make_light(1); light = shade.Light([s.to(dev) for s in lh["specular"]], lh["diffuse"].to(dev))
        for (W, H, tag) in ((800, 800, "800"), (3840, 2160, "4k")):
            V = 200
            cams = [scene.orbit_camera(k, V, W, H) for k in gstep.shard_views(V, rank, world)]
            rays = scene.canonical_rays(cams[0].to(dev), dev)
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            for c in cams:
                gframe.pbr_frame_eval(g, c.to(dev), light, lut, rays, bg, GI, inference=True)
            e1.record()
            torch.cuda.synchronize()
            ms = maxr(e0.elapsed_time(e1))
            out[f"c5_relight_sweep_{tag}"] = {"views": V, "total_ms": ms, "views_per_s": V * 1e3 / ms}
            gframe._workspaces.clear()
            torch.cuda.empty_cache()

    if rank == 0:
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
