"""Writes tests/golden/optim_ref.npz: (1) a 12-step trajectory of torch.optim.Adam — the optimiser the reference
itself calls (scene/gaussian_model.py:346, train.py:218) — on seeded parameters/gradients with the reference's group
settings, including all-zero-gradient steps (the PBR stage) and the clamped cubemap group; (2) values of the
reference's own utils/general_utils.get_expon_lr_func, imported from /root/reference. Run in the build container
(CPU): python tests/make_golden_optim.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    sys.path.insert(0, "/root/reference")
    from utils.general_utils import get_expon_lr_func
    out = {}
    steps = [0, 1, 7, 100, 2999, 15000, 30000, 40000, -5]
    f = get_expon_lr_func(lr_init=0.00016 * 5.2, lr_final=0.0000016 * 5.2, lr_delay_mult=0.01, max_steps=30000)
    out["lr_steps"] = np.array(steps)
    out["lr_xyz"] = np.array([f(s) for s in steps], dtype=np.float64)
    f2 = get_expon_lr_func(lr_init=0.05, lr_final=0.005, lr_delay_mult=0.01, max_steps=10000)
    out["lr_brdf"] = np.array([f2(s) for s in steps], dtype=np.float64)
    f3 = get_expon_lr_func(lr_init=0.01, lr_final=0.0001, lr_delay_steps=500, lr_delay_mult=0.01, max_steps=1000)
    out["lr_delay"] = np.array([f3(s) for s in steps], dtype=np.float64)

    g = torch.Generator().manual_seed(11)
    n = 1537
    p0 = torch.randn(n, generator=g)
    c0 = torch.rand(n, generator=g) * 0.01     # cubemap-like: near zero so that the clamp acts
    grads = torch.randn(12, n, generator=g) * torch.logspace(-6, 0, n)[None]
    grads[5:9] = 0.0                            # PBR-stage style: zero gradients, moments keep decaying
    p = p0.clone().requires_grad_(True)
    c = c0.clone().requires_grad_(True)
    opt = torch.optim.Adam([dict(params=[p], lr=0.0025, name="f_dc")], lr=0.0, eps=1e-15, foreach=False)
    lopt = torch.optim.Adam([dict(params=[c], lr=0.05, name="cubemap")], lr=0.05, foreach=False)
    traj_p, traj_c = [], []
    for t in range(12):
        p.grad = grads[t].clone()
        c.grad = grads[t].flip(0).clone()
        opt.step()
        lopt.step()
        with torch.no_grad():
            c.clamp_(min=0.0)
        traj_p.append(p.detach().clone())
        traj_c.append(c.detach().clone())
    out.update(p0=p0.numpy(), c0=c0.numpy(), grads=grads.numpy(), traj_p=torch.stack(traj_p).numpy(),
               traj_c=torch.stack(traj_c).numpy(),
               m_p=opt.state[p]["exp_avg"].numpy(), v_p=opt.state[p]["exp_avg_sq"].numpy())
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "optim_ref.npz"), **out)
    print("wrote optim_ref.npz", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
