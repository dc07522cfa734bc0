"""First-stage training iteration (render + L1/SSIM + normal losses + general backward + fused Adam) on one GPU at the
headline scene size, with the per-stage device times of our kernels: python tools/bench_stage1.py [P] [W] [H]"""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gi-gs_b200"))
sys.path.insert(0, ROOT)
import torch

from bench import STAGE_NAMES
from gigs import _lib, optim as gopt, scene, step as gstep

GI = dict(radius=0.8, bias=0.01, thick=0.05, delta=0.0625, step=16, start=64)


def main():
    P = int(sys.argv[1]) if len(sys.argv) > 1 else 300000
    W = int(sys.argv[2]) if len(sys.argv) > 2 else 800
    H = int(sys.argv[3]) if len(sys.argv) > 3 else 800
    dev = torch.device("cuda:0")
    L = _lib.load()
    raw = scene.make_scene(P, seed=0, regime="trained")
    cams = [scene.orbit_camera(k, 8, W, H).to(dev) for k in range(8)]
    g = torch.Generator().manual_seed(7)
    gts = [torch.rand(3, H, W, generator=g).to(dev) for _ in range(8)]
    bg = torch.zeros(3, device=dev)
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)
    out = {"P": P, "W": W, "H": H}
    for name, fused, with_opt, frame in (("operators_fused_losses", True, False, False),
                                         ("operators_framework_losses", False, False, False),
                                         ("operators_fused_losses_and_optimizer", True, True, False),
                                         ("fused_frame", True, False, True),
                                         ("fused_frame_and_optimizer", True, True, True)):
        params = gstep.GaussianParams(raw, dev)
        opt = gopt.GaussianOptimizer(params) if with_opt else None

        def one(i):
            if opt is None:
                params.zero_grad()
            gstep.first_stage_step(params, cams[i % 8], gts[i % 8], bg, GI, fused_losses=fused, fused=frame)
            if opt is not None:
                opt.step(light=False)
        for i in range(3):
            one(i)
        ts = []
        for i in range(8):
            flush.fill_(float(i))
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); one(i); e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts.sort()
        out[name] = {"ms": ts[len(ts) // 2], "iterations/s": 1e3 / ts[len(ts) // 2]}
        if with_opt:
            L.gigs_profile_enable(1)
            n = 4
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(n):
                one(i)
            e1.record()
            torch.cuda.synchronize()
            st = (C.c_int32 * 4096)(); ms = (C.c_float * 4096)()
            cnt = L.gigs_profile_read(st, ms, 4096)
            L.gigs_profile_enable(0)
            agg = {}
            for j in range(cnt):
                agg[STAGE_NAMES[st[j]]] = agg.get(STAGE_NAMES[st[j]], 0.0) + ms[j] / n
            agg.pop("radix_sort_pass", None)
            out[name]["our_kernels_ms"] = agg
            out[name]["our_kernels_total_ms"] = sum(agg.values())
            out[name]["ms_warm_l2"] = e0.elapsed_time(e1) / n
    print(json.dumps(out))


if __name__ == "__main__":
    main()
