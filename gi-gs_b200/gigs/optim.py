"""The optimiser step of the training loop as one launch of `gigs_adam_step`.

Mirrors, name for name, what /root/reference does around `loss.backward()`:
  * GaussianModel.training_setup (scene/gaussian_model.py:318-359): torch.optim.Adam(lr=0, eps=1e-15) over the groups
    xyz, f_dc, f_rest, opacity, normal, albedo, roughness, metallic, scaling, rotation with their learning rates;
  * GaussianModel.update_learning_rate (:386-395) with its early `return` (see update_learning_rate below);
  * the light optimiser (train.py:215-218: Adam over the base cubemap, lr = opacity_lr, default eps 1e-8) and
    `cubemap.clamp_(min=0.0)` (train.py:523);
  * `zero_grad` of both optimisers (train.py:518,522), fused into the same pass.
There is no framework fallback: without libgigs_b200.so the step raises.
"""
import ctypes as C
import math
from dataclasses import dataclass
from typing import Callable, Dict, Iterable, List, Optional

import torch

from . import _lib

# GaussianParams key -> the reference's param-group name (scene/gaussian_model.py:325-345)
REFERENCE_GROUP_NAME = {"xyz": "xyz", "f_dc": "f_dc", "f_rest": "f_rest", "opacity": "opacity", "normal": "normal",
                        "albedo": "albedo", "roughness": "roughness", "metallic": "metallic", "log_scale": "scaling",
                        "rot": "rotation"}


@dataclass
class OptimizationParams:
    """arguments/__init__.py:78-98 (defaults)."""
    iterations: int = 30_000
    position_lr_init: float = 0.00016
    position_lr_final: float = 0.0000016
    position_lr_delay_mult: float = 0.01
    position_lr_max_steps: int = 30_000
    feature_lr: float = 0.0025
    opacity_lr: float = 0.05
    BRDF_lr: float = 0.005
    scaling_lr: float = 0.005
    rotation_lr: float = 0.001
    percent_dense: float = 0.01
    lambda_dssim: float = 0.2
    densification_interval: int = 100
    opacity_reset_interval: int = 3000
    densify_from_iter: int = 500
    densify_until_iter: int = 15_000
    densify_grad_threshold: float = 0.0002


def get_expon_lr_func(lr_init: float, lr_final: float, lr_delay_steps: int = 0, lr_delay_mult: float = 1.0,
                      max_steps: int = 1000000) -> Callable[[int], float]:
    """utils/general_utils.py:33-70: log-linear interpolation lr_init -> lr_final over max_steps, optionally eased in."""
    def helper(step: int) -> float:
        if step < 0 or (lr_init == 0.0 and lr_final == 0.0):
            return 0.0
        if lr_delay_steps > 0:
            delay_rate = lr_delay_mult + (1 - lr_delay_mult) * math.sin(
                0.5 * math.pi * min(max(step / lr_delay_steps, 0.0), 1.0))
        else:
            delay_rate = 1.0
        t = min(max(step / max_steps, 0.0), 1.0)
        return delay_rate * math.exp(math.log(lr_init) * (1 - t) + math.log(lr_final) * t)
    return helper


class FusedAdam:
    """torch.optim.Adam semantics (amsgrad off, no weight decay, not maximize) over named groups of ONE tensor each;
    `step()` is one kernel launch for all groups. State layout follows torch: per parameter `step`, `exp_avg`,
    `exp_avg_sq`, created on the first step that sees the parameter."""

    def __init__(self, groups: Iterable[Dict], lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8):
        self.param_groups: List[Dict] = []
        self.state: Dict[str, Dict] = {}
        for g in groups:
            p = g["params"][0] if isinstance(g["params"], (list, tuple)) else g["params"]
            if p.dtype != torch.float32 or not p.is_contiguous():
                raise ValueError("FusedAdam: parameters must be contiguous float32 tensors")
            self.param_groups.append(dict(name=g["name"], params=[p], lr=float(g.get("lr", lr)),
                                          betas=tuple(g.get("betas", betas)), eps=float(g.get("eps", eps)),
                                          clamp_min0=bool(g.get("clamp_min0", False))))

    def group(self, name: str) -> Dict:
        for g in self.param_groups:
            if g["name"] == name:
                return g
        raise KeyError(name)

    def _state_of(self, g: Dict) -> Dict:
        st = self.state.get(g["name"])
        p = g["params"][0]
        if st is None or st["exp_avg"].shape != p.shape:
            st = dict(step=0, exp_avg=torch.zeros_like(p), exp_avg_sq=torch.zeros_like(p))
            self.state[g["name"]] = st
        return st

    def step(self, zero_grads: Iterable[str] = (), clear_grad: bool = True, grads: Optional[Dict] = None) -> None:
        """One Adam update of every group.
        zero_grads: names of groups whose gradient is known to be all zero (not read; torch's Adam would read zeros).
        clear_grad: zero the gradients that were read, in the same pass (`zero_grad`).
        grads: optional name -> gradient tensor; default is the parameter's .grad. A parameter whose .grad is None
        and that is not in zero_grads is skipped, as torch does."""
        L = _lib.load()
        zero = set(zero_grads)
        if any(not g["params"][0].is_cuda for g in self.param_groups):
            raise RuntimeError("FusedAdam.step: parameters must be CUDA tensors (there is no CPU fallback)")
        arr = (_lib.GigsAdamGroup * len(self.param_groups))()
        n = 0
        keep = []
        for g in self.param_groups:
            p = g["params"][0]
            gr = None
            if g["name"] not in zero:
                gr = grads.get(g["name"]) if grads is not None else p.grad
                if gr is None:
                    continue
                if gr.dtype != torch.float32 or not gr.is_contiguous() or gr.numel() != p.numel():
                    raise ValueError(f"FusedAdam: gradient of '{g['name']}' must be contiguous float32 of the parameter's size")
            st = self._state_of(g)
            st["step"] += 1
            a = arr[n]
            n += 1
            a.param, a.grad = p.data_ptr(), (gr.data_ptr() if gr is not None else None)
            a.exp_avg, a.exp_avg_sq = st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr()
            a.count = p.numel()
            a.lr, a.beta1, a.beta2, a.eps = g["lr"], g["betas"][0], g["betas"][1], g["eps"]
            a.step, a.clamp_min0, a.clear_grad = st["step"], int(g["clamp_min0"]), int(clear_grad)
            keep.append((p, gr))
        if n == 0:
            return
        dev = self.param_groups[0]["params"][0].device
        with torch.cuda.device(dev):
            _lib.check(L.gigs_adam_step(n, arr, torch.cuda.current_stream().cuda_stream), "gigs_adam_step")

    # torch.optim.Optimizer-shaped state for GaussianModel.capture() / restore() (scene/gaussian_model.py:82-131)
    def state_dict(self) -> Dict:
        state = {}
        for i, g in enumerate(self.param_groups):
            st = self.state.get(g["name"])
            if st is not None:
                state[i] = dict(step=torch.tensor(float(st["step"])), exp_avg=st["exp_avg"], exp_avg_sq=st["exp_avg_sq"])
        groups = [dict(name=g["name"], lr=g["lr"], betas=g["betas"], eps=g["eps"], weight_decay=0, amsgrad=False,
                       maximize=False, foreach=None, capturable=False, differentiable=False, fused=None, params=[i])
                  for i, g in enumerate(self.param_groups)]
        return dict(state=state, param_groups=groups)

    def load_state_dict(self, sd: Dict) -> None:
        for i, g in enumerate(self.param_groups):
            src = sd["param_groups"][i]
            g["lr"], g["betas"], g["eps"] = float(src["lr"]), tuple(src["betas"]), float(src["eps"])
            st = sd["state"].get(i)
            if st is not None:
                p = g["params"][0]
                self.state[g["name"]] = dict(step=int(float(st["step"])),
                                             exp_avg=st["exp_avg"].to(p.device, torch.float32).contiguous().clone(),
                                             exp_avg_sq=st["exp_avg_sq"].to(p.device, torch.float32).contiguous().clone())


class GaussianOptimizer:
    """Both optimisers of the reference's loop over a gigs.step.GaussianParams: the Gaussians' (eps 1e-15) and, when
    the light is trainable, the light's (eps 1e-8, clamped at 0 after the step) — stepped by one launch."""

    def __init__(self, params, opt: Optional[OptimizationParams] = None, spatial_lr_scale: float = 1.0,
                 train_light: bool = True):
        self.params = params
        self.opt = opt = opt or OptimizationParams()
        L = params.leaves
        lrs = {"xyz": opt.position_lr_init * spatial_lr_scale, "f_dc": opt.feature_lr, "f_rest": opt.feature_lr / 20.0,
               "opacity": opt.opacity_lr, "normal": opt.opacity_lr, "albedo": opt.opacity_lr,
               "roughness": opt.opacity_lr, "metallic": opt.opacity_lr, "log_scale": opt.scaling_lr,
               "rot": opt.rotation_lr}
        groups = [dict(params=[L[k]], lr=lrs[k], name=REFERENCE_GROUP_NAME[k], eps=1e-15) for k in lrs]
        self._key_of = {REFERENCE_GROUP_NAME[k]: k for k in lrs}
        if train_light:
            if params.light_base is not None:
                groups.append(dict(params=[params.light_base], lr=opt.opacity_lr, name="cubemap", eps=1e-8,
                                   clamp_min0=True))
                self._key_of["cubemap"] = "light_base"
            else:
                for i, t in enumerate(params.light_leaves):   # ready-made textures trained directly (not a reference mode)
                    groups.append(dict(params=[t], lr=opt.opacity_lr, name=f"light{i}", eps=1e-8, clamp_min0=True))
                    self._key_of[f"light{i}"] = f"light{i}"
        self.adam = FusedAdam(groups)
        self.xyz_scheduler_args = get_expon_lr_func(opt.position_lr_init * spatial_lr_scale,
                                                    opt.position_lr_final * spatial_lr_scale,
                                                    lr_delay_mult=opt.position_lr_delay_mult,
                                                    max_steps=opt.position_lr_max_steps)
        self.BRDF_scheduler_args = get_expon_lr_func(opt.opacity_lr, opt.BRDF_lr,
                                                     lr_delay_mult=opt.position_lr_delay_mult, max_steps=10000)

    def _name_of(self, key: str) -> str:
        return REFERENCE_GROUP_NAME.get(key, key)

    def rebind(self, new_state: Optional[Dict[str, Dict]] = None) -> None:
        """After GaussianParams.rebuild: point the Gaussian groups at the new leaves; new_state maps a GaussianParams key
        to {'exp_avg', 'exp_avg_sq'} of the new size (step counts are kept, as the reference keeps state['step'] through
        cat_tensors_to_optimizer / _prune_optimizer). A group without an entry restarts from empty state."""
        for g in self.adam.param_groups:
            key = self._key_of[g["name"]]
            if key not in self.params.leaves:
                continue
            g["params"] = [self.params.leaves[key]]
            st = self.adam.state.get(g["name"])
            ns = (new_state or {}).get(key)
            if ns is None:
                self.adam.state.pop(g["name"], None)
            else:
                self.adam.state[g["name"]] = dict(step=st["step"] if st is not None else 0, exp_avg=ns["exp_avg"],
                                                  exp_avg_sq=ns["exp_avg_sq"])

    def update_learning_rate(self, iteration: int) -> Optional[float]:
        """scene/gaussian_model.py:386-395, including its control flow: groups are visited in order, `xyz` gets its
        scheduled rate, and the first BRDF group met (`albedo`) gets BRDF_scheduler(iteration - 30000) and RETURNS, so
        `roughness` / `metallic` keep opacity_lr and albedo's rate is 0 until iteration 30000."""
        for g in self.adam.param_groups:
            if g["name"] in ("albedo", "roughness", "metallic"):
                lr = self.BRDF_scheduler_args(iteration - 30000)
                g["lr"] = lr
                return lr
            if g["name"] == "xyz":
                g["lr"] = self.xyz_scheduler_args(iteration)
        return None

    def step(self, light: bool = True, skip_groups=()) -> None:
        """optimizer.step() + zero_grad() (+ light_optimizer.step() + zero_grad() + clamp_ when `light`,
        train.py:516-523). Groups the fused frame is known not to have written since the last clear pass a NULL
        gradient: same update as torch's (their gradients ARE zero), 4 B per element less traffic and no clear.
        skip_groups: reference group names whose tensors torch's Adam would skip this iteration because their .grad is
        None (a parameter replaced after backward, e.g. `opacity` by reset_opacity): not stepped, step count kept."""
        p = self.params
        zero = []
        if p._dirty is not None:
            dirty = {s for s in p._dirty}
            for g in self.adam.param_groups:
                if p._span[self._key_of[g["name"]]] not in dirty:
                    zero.append(g["name"])
        skip = [] if light else [g["name"] for g in self.adam.param_groups
                                 if g["name"] == "cubemap" or g["name"].startswith("light")]
        skip += [n for n in skip_groups if n not in skip]
        if skip:
            saved = self.adam.param_groups
            self.adam.param_groups = [g for g in saved if g["name"] not in skip]
            try:
                self.adam.step(zero_grads=zero, clear_grad=True)
            finally:
                self.adam.param_groups = saved
            # a skipped Gaussian group's gradient is dropped (the reference's new Parameter has none), the light's stays
            for n in skip_groups:
                t = self.params.leaves.get(self._key_of.get(n, n))
                if t is not None and t.grad is not None:
                    t.grad.zero_()
            kept = [n for n in skip if n not in skip_groups]
            if p._dirty is not None:
                p._dirty = [s for s in p._dirty if any(s == p._span[self._key_of[n]] for n in kept)]
        else:
            self.adam.step(zero_grads=zero, clear_grad=True)
            p._dirty = []
