"""One iteration of the reference's training loop (/root/reference/train.py:246-523) assembled from the pieces of this
package: which stage runs, when the model is densified / pruned / its opacity reset, the learning-rate schedule and the
optimiser step. The driver around it (argument parsing, dataset readers, logging, TensorBoard, evaluation, file layout)
is outside the hot path and not provided.

    iteration <= pbr_iteration : first stage  (render + L1/SSIM + normal losses; all 10 groups train; densification)
    iteration >  pbr_iteration : PBR stage    (build_mips, G-buffer + SSAO + split-sum shading + SSR, L1 + priors;
                                               materials and the light train; fused frame)
"""
from dataclasses import dataclass, field
from typing import Callable, Dict, List, Optional

import torch

from . import densify as _densify
from .optim import GaussianOptimizer, OptimizationParams
from .step import GaussianParams, first_stage_step, training_step


@dataclass
class TrainConfig:
    """train.py:171-192 keyword defaults + the GI arguments."""
    pbr_iteration: int = 30_000
    metallic: bool = True
    tone: bool = False
    gamma: bool = True
    indirect: bool = True
    normal_tv_weight: float = 1.0
    brdf_tv_weight: float = 1.0
    env_tv_weight: float = 0.01
    white_background: bool = False
    gi: Dict = field(default_factory=lambda: dict(radius=0.8, bias=0.01, thick=0.05, delta=0.0625, step=16, start=64))
    opt: OptimizationParams = field(default_factory=OptimizationParams)


class Trainer:
    def __init__(self, params: GaussianParams, brdf_lut: torch.Tensor, cameras_extent: float, cfg: Optional[TrainConfig] = None,
                 spatial_lr_scale: float = 1.0, rays_of: Optional[Callable] = None, fused: bool = True):
        self.fused = fused          # False: operator path + autograd for both stages (the parity reference)
        self.params, self.lut, self.extent = params, brdf_lut, float(cameras_extent)
        self.cfg = cfg or TrainConfig()
        self.optimizer = GaussianOptimizer(params, self.cfg.opt, spatial_lr_scale)
        self.stats = _densify.DensifyState(params.P, params.flat_grad.device)
        self.spatial_lr_scale = spatial_lr_scale
        self.rays_of = rays_of
        dev = params.flat_grad.device
        self.bg = torch.tensor([1.0, 1.0, 1.0] if self.cfg.white_background else [0.0, 0.0, 0.0], device=dev)
        self.log: List[Dict] = []

    def iteration(self, it: int, cam, gt_image: torch.Tensor, rays: Optional[torch.Tensor] = None) -> torch.Tensor:
        """train.py:246-523 for iteration `it` (1-based) on one view. Returns the detached loss."""
        cfg, opt, p = self.cfg, self.cfg.opt, self.params
        # "Every 1000 its we increase the levels of SH up to a maximum degree" (train.py:243-244, oneupSHdegree)
        if it % 1000 == 0 and p.sh_degree < p.max_sh_degree:
            p.sh_degree += 1
        first = it <= cfg.pbr_iteration
        if first:
            loss, _ = first_stage_step(p, cam, gt_image, self.bg, cfg.gi, lambda_dssim=opt.lambda_dssim,
                                       normal_tv_weight=cfg.normal_tv_weight, fused=self.fused,
                                       stats=self.stats if it < opt.densify_until_iter else None)
        else:
            if rays is None:
                rays = self.rays_of(cam)
            loss = training_step(p, cam, p.light(), self.lut, rays, gt_image, torch.zeros_like(self.bg), cfg.gi,
                                 metallic=cfg.metallic, gamma=cfg.gamma, tone=cfg.tone, indirect=cfg.indirect,
                                 fused=self.fused, brdf_tv_weight=cfg.brdf_tv_weight,
                                 env_tv_weight=cfg.env_tv_weight if p.prefiltered is not None else 0.0)
        event = None
        with torch.no_grad():
            if it < opt.densify_until_iter:                                        # train.py:488-511
                if not first:        # the PBR stage's fused frame carries no screen-space gradient; the reference's
                    pass             # schedule ends densification (15k) long before the PBR stage starts (30k)
                if it > opt.densify_from_iter and it % opt.densification_interval == 0:
                    size_threshold = 20 if it > opt.opacity_reset_interval else None
                    event = _densify.densify_and_prune(p, self.optimizer, self.stats, opt.densify_grad_threshold, 0.05,
                                                       self.extent, size_threshold, percent_dense=opt.percent_dense)
                if it % opt.opacity_reset_interval == 0 or (cfg.white_background and it == opt.densify_from_iter):
                    _densify.reset_opacity(p, self.optimizer)
                    event = dict(event or {}, reset_opacity=True)
            if it < opt.iterations:                                                # train.py:515-523
                # a rebuilt model has fresh zero gradients (the reference also steps on the new, gradient-less tensors:
                # torch's Adam skips parameters whose .grad is None)
                if event is None or "P_after" not in event:
                    # torch's Adam skips a parameter whose .grad is None: `opacity` in the iteration that reset it
                    # (reset_opacity swaps in a new Parameter after backward), and the light before the PBR stage has
                    # produced its first gradient (at it == pbr_iteration the loop calls light_optimizer.step() on a
                    # light that was never rendered: no update, no step count)
                    reset = event is not None and event.get("reset_opacity")
                    self.optimizer.step(light=(it >= cfg.pbr_iteration and not first),
                                        skip_groups=("opacity",) if reset else ())
                self.optimizer.update_learning_rate(it)
        self.log.append(dict(iteration=it, stage=1 if first else 2, loss=float(loss), P=p.P, event=event))
        return loss
