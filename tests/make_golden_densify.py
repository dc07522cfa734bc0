"""Writes tests/golden/densify_ref.npz by running the reference's OWN GaussianModel methods (densify_and_prune,
densify_and_clone, densify_and_split, densification_postfix, cat_tensors_to_optimizer, prune_points, _prune_optimizer,
add_densification_stats, reset_opacity, replace_tensor_to_optimizer; scene/gaussian_model.py) on CPU: the class itself
cannot be imported here (simple_knn / pytorch3d / plyfile are absent and it hard-codes device="cuda"), so the methods are
compiled from the source file with `device="cuda"` rewritten to "cpu" and bound to a stub that holds the same attributes.
torch.normal is replaced by `noise * std` with recorded standard-normal noise so that the sampled positions can be
reproduced. Run in the build container: python tests/make_golden_densify.py"""
import ast
import os
import sys
import types

import numpy as np
import torch
from torch import nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gi-gs_b200"))
METHODS = ("densify_and_prune", "densify_and_clone", "densify_and_split", "densification_postfix",
           "cat_tensors_to_optimizer", "prune_points", "_prune_optimizer", "add_densification_stats", "reset_opacity",
           "replace_tensor_to_optimizer")
KEYS = ("xyz", "f_dc", "f_rest", "opacity", "normal", "albedo", "roughness", "metallic", "log_scale", "rot")
ATTR = {"xyz": "_xyz", "f_dc": "_features_dc", "f_rest": "_features_rest", "opacity": "_opacity", "normal": "_normal",
        "albedo": "_albedo", "roughness": "_roughness", "metallic": "_metallic", "log_scale": "_scaling", "rot": "_rotation"}
NAME = {"xyz": "xyz", "f_dc": "f_dc", "f_rest": "f_rest", "opacity": "opacity", "normal": "normal", "albedo": "albedo",
        "roughness": "roughness", "metallic": "metallic", "log_scale": "scaling", "rot": "rotation"}


class NoiseTorch:
    """`torch` for the compiled reference methods: everything falls through, except normal(), which consumes recorded
    standard-normal noise (torch.normal(mean, std) == randn * std + mean)."""

    def __init__(self, gen):
        self.gen = gen
        self.drawn = []

    def __getattr__(self, k):
        return getattr(torch, k)

    def normal(self, mean, std):
        n = torch.randn(std.shape, generator=self.gen)
        self.drawn.append(n)
        return n * std + mean


def reference_model(nt):
    src = open("/root/reference/scene/gaussian_model.py").read().replace('device="cuda"', 'device="cpu"')
    cls = next(n for n in ast.parse(src).body if isinstance(n, ast.ClassDef) and n.name == "GaussianModel")
    gsrc = open("/root/reference/utils/general_utils.py").read().replace('device="cuda"', 'device="cpu"')
    gfun = [n for n in ast.parse(gsrc).body if isinstance(n, ast.FunctionDef) and n.name in ("build_rotation", "inverse_sigmoid")]
    from typing import Dict, List, Optional, Tuple
    ns = {"torch": nt, "nn": nn, "Dict": Dict, "List": List, "Optional": Optional, "Tuple": Tuple}
    exec(compile(ast.Module(body=gfun, type_ignores=[]), "general_utils.py", "exec"), ns)
    body = [n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name in METHODS]
    stub = ast.ClassDef(name="RefModel", bases=[], keywords=[], body=body, decorator_list=[])
    mod = ast.Module(body=[stub], type_ignores=[])
    ast.fix_missing_locations(mod)
    exec(compile(mod, "/root/reference/scene/gaussian_model.py", "exec"), ns)
    R = ns["RefModel"]
    R.get_scaling = property(lambda self: torch.exp(self._scaling))
    R.get_opacity = property(lambda self: torch.sigmoid(self._opacity))
    R.get_xyz = property(lambda self: self._xyz)
    R.scaling_inverse_activation = staticmethod(torch.log)
    return R


def build(P, seed, extent):
    from gigs import scene
    raw = scene.make_scene(P, seed=seed, regime="trained")
    g = torch.Generator().manual_seed(seed + 100)
    # a spread of scales around percent_dense * extent so that both clone and split fire
    raw["log_scale"] = raw["log_scale"] + torch.randn(P, 1, generator=g) * 1.0 + 0.6
    raw["opacity"] = raw["opacity"] - 2.5 * (torch.rand(P, 1, generator=g) < 0.2)
    return raw, g


def run_case(P, seed, extent, max_grad, max_screen_size, views=3):
    raw, g = build(P, seed, extent)
    nt = NoiseTorch(g)
    R = reference_model(nt)
    m = R()
    m.percent_dense = 0.01
    for k in KEYS:
        setattr(m, ATTR[k], nn.Parameter(raw[k].clone().requires_grad_(True)))
    m.optimizer = torch.optim.Adam([dict(params=[getattr(m, ATTR[k])], lr=1e-3, name=NAME[k]) for k in KEYS], lr=0.0,
                                   eps=1e-15)
    # two optimiser steps so that the moments are non-trivial
    for _ in range(2):
        for k in KEYS:
            p = getattr(m, ATTR[k])
            p.grad = torch.randn(p.shape, generator=g) * 1e-3
        m.optimizer.step()
    out = {f"p_{k}": getattr(m, ATTR[k]).detach().clone().numpy() for k in KEYS}
    for k in KEYS:
        st = m.optimizer.state[getattr(m, ATTR[k])]
        out[f"m_{k}"], out[f"v_{k}"] = st["exp_avg"].clone().numpy(), st["exp_avg_sq"].clone().numpy()
    z = lambda *s: torch.zeros(*s)
    m.xyz_gradient_accum, m.xyz_gradient_accum_abs, m.xyz_gradient_accum_abs_max, m.denom = z(P, 1), z(P, 1), z(P, 1), z(P, 1)
    m.max_radii2D = z(P)
    radii_all, grad_all = [], []
    for v in range(views):
        radii = torch.randint(-3, 30, (P,), generator=g).int()
        grad = torch.randn(P, 3, generator=g) * 3e-4
        vis = radii > 0
        m.max_radii2D[vis] = torch.max(m.max_radii2D[vis], radii[vis].float())       # train.py:491-493
        vp = types.SimpleNamespace(grad=grad)
        m.add_densification_stats(vp, vis)
        radii_all.append(radii)
        grad_all.append(grad)
    out.update(radii=torch.stack(radii_all).numpy(), grad2D=torch.stack(grad_all).numpy(),
               accum=m.xyz_gradient_accum.clone().numpy(), accum_abs=m.xyz_gradient_accum_abs.clone().numpy(),
               accum_abs_max=m.xyz_gradient_accum_abs_max.clone().numpy(), denom=m.denom.clone().numpy(),
               max_radii2D=m.max_radii2D.clone().numpy())
    m.densify_and_prune(max_grad, 0.05, extent, max_screen_size)
    n_clone, n_child = (nt.drawn[0].shape[0], nt.drawn[1].shape[0])
    out["noise_clone"], out["noise_split"] = nt.drawn[0].numpy(), nt.drawn[1].numpy()
    for k in KEYS:
        p = getattr(m, ATTR[k])
        out[f"q_{k}"] = p.detach().numpy()
        st = m.optimizer.state[p]
        out[f"qm_{k}"], out[f"qv_{k}"] = st["exp_avg"].numpy(), st["exp_avg_sq"].numpy()
        out[f"step_{k}"] = np.array(float(st["step"]))
    out["after_accum"] = m.xyz_gradient_accum.numpy()
    out["after_max_radii2D"] = m.max_radii2D.numpy()
    m.reset_opacity()
    out["reset_opacity"] = m._opacity.detach().numpy()
    out["reset_m"] = m.optimizer.state[m._opacity]["exp_avg"].numpy()
    out["meta"] = np.array([P, seed, extent, max_grad, max_screen_size or 0, n_clone, n_child], dtype=np.float64)
    print(f"P {P} -> {m._xyz.shape[0]}  clones {n_clone}  split children {n_child}")
    return out


def main():
    cases = [run_case(1000, 1, 5.0, 0.0002, 20), run_case(500, 2, 3.0, 0.0003, None)]
    flat = {}
    for i, c in enumerate(cases):
        for k, v in c.items():
            flat[f"c{i}_{k}"] = v
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "densify_ref.npz"), **flat)
    print("wrote densify_ref.npz")


if __name__ == "__main__":
    main()
