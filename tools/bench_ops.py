"""Micro-timings of the training-loop kernels outside the frame (Adam step, image loss) on one GPU: CUDA events on the
launching stream, 256 MiB L2 flush between iterations. python tools/bench_ops.py [P]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gi-gs_b200"))
import torch

from gigs import losses, optim as gopt, scene, step as gstep


def ev_time(fn, flush, reps=10, warm=3):
    for _ in range(warm):
        fn()
    ts = []
    for i in range(reps):
        flush.fill_(float(i))
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    P = int(sys.argv[1]) if len(sys.argv) > 1 else 300000
    dev = torch.device("cuda:0")
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)
    out = {"P": P, "adam_vec": os.environ.get("GIGS_ADAM_VEC", "default")}
    raw = scene.make_scene(P, seed=0, regime="trained")
    base = torch.rand(6, 256, 256, 3) * 0.5 + 0.25
    p = gstep.GaussianParams(raw, dev)
    p.light_base = base.to(dev).requires_grad_(True)       # no prefilter needed for the optimiser timing
    p.light_base.grad = torch.zeros_like(p.light_base)
    p._span["light_base"] = (-1, -1)
    o = gopt.GaussianOptimizer(p)
    n_el = sum(t.numel() for t in p.leaves.values()) + p.light_base.numel()
    for nm, zero in (("adam_all_grads", []), ("adam_pbr_stage", [g["name"] for g in o.adam.param_groups
                                                                  if g["name"] not in ("albedo", "roughness", "metallic", "cubemap")])):
        ms = ev_time(lambda: o.adam.step(zero_grads=zero, clear_grad=True), flush)
        n_gr = sum(g["params"][0].numel() for g in o.adam.param_groups if g["name"] not in zero)
        by = 24 * n_el + 8 * n_gr
        out[nm] = {"ms": ms, "GB/s": by / ms / 1e6, "bytes": by}
    # torch's own Adam over the same tensors (foreach), for scale
    leaves = [t.detach().clone().requires_grad_(True) for t in p.leaves.values()]
    for t in leaves:
        t.grad = torch.zeros_like(t)
    topt = torch.optim.Adam(leaves, lr=1e-3, eps=1e-15)
    out["torch_adam_foreach"] = {"ms": ev_time(lambda: (topt.step(), topt.zero_grad(set_to_none=False)), flush)}
    g = torch.Generator().manual_seed(0)
    gt = torch.rand(3, 800, 800, generator=g).to(dev)
    img = (gt + 0.1 * torch.randn(3, 800, 800, generator=g).to(dev)).clamp(0, 1).requires_grad_(True)

    def ours():
        img.grad = None
        losses.l1_ssim_loss(img, gt, 0.2).backward()
    out["image_loss_fwd_bwd"] = {"ms": ev_time(ours, flush)}
    with torch.no_grad():
        out["image_loss_fwd_only"] = {"ms": ev_time(lambda: losses.l1_ssim_loss(img.detach(), gt, 0.2), flush)}
    # the reference's formulation in framework ops (utils/loss_utils.py:54-100 restated with conv2d), for scale
    import torch.nn.functional as F
    w1 = torch.tensor([__import__("math").exp(-((x - 5) ** 2) / 4.5) for x in range(11)])
    w1 = (w1 / w1.sum()).unsqueeze(1)
    win = w1.mm(w1.t()).float()[None, None].expand(3, 1, 11, 11).contiguous().to(dev)

    def ref_ops():
        img.grad = None
        a, b = img[None], gt[None]
        mu1, mu2 = F.conv2d(a, win, padding=5, groups=3), F.conv2d(b, win, padding=5, groups=3)
        s11 = F.conv2d(a * a, win, padding=5, groups=3) - mu1 * mu1
        s22 = F.conv2d(b * b, win, padding=5, groups=3) - mu2 * mu2
        s12 = F.conv2d(a * b, win, padding=5, groups=3) - mu1 * mu2
        m = ((2 * mu1 * mu2 + 1e-4) * (2 * s12 + 9e-4)) / ((mu1 * mu1 + mu2 * mu2 + 1e-4) * (s11 + s22 + 9e-4))
        (0.8 * (a - b).abs().mean() + 0.2 * (1 - m.mean())).backward()
    out["framework_ops_fwd_bwd"] = {"ms": ev_time(ref_ops, flush)}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
