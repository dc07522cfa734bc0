"""The PBR-stage training step on one view (the unit BASELINE.json's metric counts: one "frame" =
G-buffer forward + SSAO + split-sum shading + SSR + loss + full backward), and its view-sharded multi-GPU
form. Mirrors /root/reference/train.py:240-422 for the parts on the hot path; the optimiser, densification and
TV losses are outside it. The light can be given as ready-made textures (leaves: `light=`) or as the
trainable base cubemap (`light_base=`), in which case every step rebuilds the mips from it like train.py:340 does
(gigs.light.PrefilteredLight: CubemapLight.build_mips and its backward).
"""
import math
from typing import Dict, List, Optional

import torch
import torch.nn.functional as F

from . import scene as _scene
from .renderer import pbr_forward, pbr_loss
from .shade import Light

PARAM_KEYS = ("xyz", "f_dc", "f_rest", "opacity", "normal", "albedo", "roughness", "metallic", "log_scale", "rot")
PARAM_WIDTH = {"xyz": 3, "f_dc": 3, "f_rest": 45, "opacity": 1, "normal": 3, "albedo": 3, "roughness": 1,
               "metallic": 1, "log_scale": 3, "rot": 4}  # 67 floats = 268 B per Gaussian


class GaussianParams:
    """Raw (pre-activation) parameters as leaf tensors whose .grad tensors are VIEWS into one flat buffer,
    so the per-Gaussian gradient all-reduce of the view-sharded step is a single collective with no packing
    copy (autograd accumulates into an existing .grad in place)."""

    def __init__(self, raw: Dict, device, light: Optional[Dict] = None, light_base: Optional[torch.Tensor] = None,
                 cutoff: float = 0.99, peer: bool = False):
        if light is not None and light_base is not None:
            raise ValueError("give the light either as textures (light=) or as the base cubemap (light_base=)")
        P = raw["xyz"].shape[0]
        self.P = P
        self.sh_degree = raw["sh_degree"]              # active degree (GaussianModel.active_sh_degree)
        self.max_sh_degree = int(raw.get("max_sh_degree", raw["sh_degree"]))
        self.leaves: Dict[str, torch.Tensor] = {}
        for k in PARAM_KEYS:
            self.leaves[k] = raw[k].to(device).float().contiguous().requires_grad_(True)
        self.light_leaves: List[torch.Tensor] = []
        if light is not None:
            self.light_leaves = [t.to(device).float().contiguous().requires_grad_(True)
                                 for t in (light["diffuse"], *light["specular"])]
        self.light_base: Optional[torch.Tensor] = None
        extra = [(f"light{i}", t) for i, t in enumerate(self.light_leaves)]
        if light_base is not None:
            self.light_base = light_base.to(device).float().contiguous().requires_grad_(True)
            extra = [("light_base", self.light_base)]
        n = sum(t.numel() for t in self.leaves.values()) + sum(t.numel() for _, t in extra)
        # peer=True (view-sharded multi-GPU training): the gradient buffer is a symmetric allocation every rank of the
        # node can address, and all_reduce_grads is ONE kernel of ours over NVLink (gigs.peer / csrc/peer_reduce.cu)
        # instead of NCCL collectives. Falls back to a private buffer + NCCL when peer mapping is not available.
        self._peer = None
        if peer:
            from . import peer as _peer
            if _peer.available():
                try:
                    self._peer = _peer.PeerBuffer(n, device)
                except Exception as ex:      # no P2P / IPC on this box: NCCL remains the exchange
                    import warnings
                    warnings.warn(f"gigs: peer-mapped gradient buffer unavailable ({ex}); using NCCL all-reduce")
        self.flat_grad = self._peer.buf if self._peer is not None else torch.zeros(n, dtype=torch.float32, device=device)
        o = 0
        self._span = {}
        for k, t in list(self.leaves.items()) + extra:
            t.grad = self.flat_grad[o:o + t.numel()].view_as(t)
            self._span[k] = (o, o + t.numel())
            o += t.numel()
        self._dirty = None   # spans the fused path has written since the last zero_grad; None = unknown / anything
        self.prefiltered = None
        if self.light_base is not None:
            from .light import PrefilteredLight
            self.prefiltered = PrefilteredLight(self.light_base, cutoff=cutoff)

    def rebuild(self, new_leaves: Dict[str, torch.Tensor]) -> None:
        """Swap in a new set of Gaussians (densification / pruning / checkpoint restore change P): new leaf tensors, a
        new flat gradient buffer of the new size with the light's gradient carried over, all gradients zero."""
        dev = self.flat_grad.device
        P = new_leaves["xyz"].shape[0]
        self.P = P
        self.leaves = {k: new_leaves[k].detach().to(dev).float().contiguous().requires_grad_(True) for k in PARAM_KEYS}
        extra = [(f"light{i}", t) for i, t in enumerate(self.light_leaves)]
        if self.light_base is not None:
            extra = [("light_base", self.light_base)]
        old = {k: t.grad.clone() for k, t in extra if t.grad is not None}
        n = sum(t.numel() for t in self.leaves.values()) + sum(t.numel() for _, t in extra)
        if self._peer is not None:       # collective: every rank rebuilds the same (replicated) model
            from . import peer as _peer
            self._peer = _peer.PeerBuffer(n, dev)
            self.flat_grad = self._peer.buf
        else:
            self.flat_grad = torch.zeros(n, dtype=torch.float32, device=dev)
        o = 0
        self._span = {}
        for k, t in list(self.leaves.items()) + extra:
            t.grad = self.flat_grad[o:o + t.numel()].view_as(t)
            if k in old:
                t.grad.copy_(old[k])
            self._span[k] = (o, o + t.numel())
            o += t.numel()
        self._dirty = None

    def mark_dirty(self, keys=None):
        if keys is None:
            self._dirty = None
        elif self._dirty is not None:
            self._dirty.extend(self._span[k] for k in keys)

    def zero_grad(self, fused_only: bool = False):
        """Zero the flat gradient buffer. fused_only=True is the caller's statement that nothing but the fused PBR
        frame path (training_step(fused=True)) has written gradients since the last zero_grad: that path only
        produces material and light gradients (train.py:343-351 detaches everything else), so 12.6 MB are cleared
        instead of the whole 87 MB buffer (300k Gaussians). Falls back to a full clear whenever that is not known."""
        if not fused_only or self._dirty is None:
            spans = [(0, self.flat_grad.numel())]
        else:
            spans = self._merged_dirty()
        if self.flat_grad.is_cuda and 0 < len(spans) <= 8:
            # one launch of ours (chains with the frame's kernels) instead of one framework fill per span
            import ctypes as C
            from . import _lib
            _L = _lib.load()
            n = len(spans)
            lo = (C.c_uint64 * n)(*[int(a) for a, _ in spans])
            hi = (C.c_uint64 * n)(*[int(b) for _, b in spans])
            with torch.cuda.device(self.flat_grad.device):
                _lib.check(_L.gigs_clear_spans(self.flat_grad.data_ptr(), n, lo, hi,
                                               torch.cuda.current_stream(self.flat_grad.device).cuda_stream), "gigs_clear_spans")
        else:
            for lo, hi in spans:
                self.flat_grad[lo:hi].zero_()
        self._dirty = []

    def _merged_dirty(self):
        merged = []
        for lo, hi in sorted(set(self._dirty)):
            if merged and lo <= merged[-1][1]:
                merged[-1][1] = max(merged[-1][1], hi)
            else:
                merged.append([lo, hi])
        return merged

    def begin_light_all_reduce(self, light_ready) -> None:
        """Start the all-reduce of the light-texture gradients on a side stream as soon as `light_ready` (an event the
        fused backward records before its blend backward) fires, so the exchange overlaps the rest of the backward.
        all_reduce_grads(fused_only=True) then exchanges the remaining spans and waits for this one."""
        import torch.distributed as dist
        if not self.light_leaves and self.prefiltered is None:
            return
        if self._peer is not None:
            # peer mode: all spans go in ONE kernel at the end of the step (all_reduce_grads); only the mips' backward
            # (texture gradients -> base cubemap gradient) still runs under the blend backward
            if self.prefiltered is not None:
                self.light_backward_overlapped(light_ready)
            return
        if getattr(self, "_comm", None) is None:
            self._comm = torch.cuda.Stream(device=self.flat_grad.device)
        if self.prefiltered is not None:
            lo, hi = self._span["light_base"]
        else:
            lo = self._span["light0"][0]
            hi = self._span[f"light{len(self.light_leaves) - 1}"][1]
        with torch.cuda.stream(self._comm):
            self._comm.wait_event(light_ready)
            if self.prefiltered is not None:     # texture gradients -> d loss / d base, then exchange the base's
                self.prefiltered.backward(self.light_base.grad, accumulate=True)
            work = dist.all_reduce(self.flat_grad[lo:hi], op=dist.ReduceOp.SUM, async_op=True)
        self._pending_light = (work, lo, hi)

    def all_reduce_grads(self, fused_only: bool = False):
        """Sum the gradient buffer over the ranks of the default process group (view-sharded training).
        With fused_only (same contract as zero_grad) only the spans the fused PBR path writes are exchanged:
        12.6 MB instead of 87 MB at 300k Gaussians; everything else is zero on every rank."""
        import torch.distributed as dist
        if self._peer is not None:
            self._peer.all_reduce(self._merged_dirty() if (fused_only and self._dirty is not None) else None)
            return
        if not fused_only or self._dirty is None:
            pending = getattr(self, "_pending_light", None)
            if pending is not None:
                pending[0].wait()   # cannot double-count: finish it, then subtract nothing — exchange the rest only
                self._pending_light = None
                lo, hi = pending[1], pending[2]
                dist.all_reduce(self.flat_grad[:lo], op=dist.ReduceOp.SUM)
                if hi < self.flat_grad.numel():
                    dist.all_reduce(self.flat_grad[hi:], op=dist.ReduceOp.SUM)
                return
            dist.all_reduce(self.flat_grad, op=dist.ReduceOp.SUM)
            return
        pending = getattr(self, "_pending_light", None)
        for lo, hi in self._merged_dirty():
            if pending is not None and lo >= pending[1] and hi <= pending[2]:
                continue   # already on its way
            dist.all_reduce(self.flat_grad[lo:hi], op=dist.ReduceOp.SUM)
        if pending is not None:
            pending[0].wait()
            self._pending_light = None

    def activated(self) -> Dict:
        """scene/gaussian_model.py:178-266 getters, autograd-tracked."""
        L = self.leaves
        return dict(means3D=L["xyz"], opacity=torch.sigmoid(L["opacity"]), normal=F.normalize(L["normal"], dim=-1),
                    albedo=torch.sigmoid(L["albedo"]), roughness=torch.sigmoid(L["roughness"]),
                    metallic=torch.sigmoid(L["metallic"]), scales=torch.exp(L["log_scale"]),
                    rotations=F.normalize(L["rot"], dim=-1), shs=torch.cat((L["f_dc"], L["f_rest"]), dim=1),
                    sh_degree=self.sh_degree)

    def light(self):
        """The light the shading reads: ready-made leaf textures, or the PrefilteredLight built from `light_base`."""
        if self.prefiltered is not None:
            return self.prefiltered
        if not self.light_leaves:
            return None
        return Light(specular=list(self.light_leaves[1:]), diffuse=self.light_leaves[0])

    def env_dirs(self) -> torch.Tensor:
        """The fixed lat-long directions of the env-map TV prior (train.py:209 computes them once as well)."""
        if getattr(self, "_env_dirs", None) is None:
            from .light import envmap_dirs
            self._env_dirs = envmap_dirs(device=self.flat_grad.device)
        return self._env_dirs

    def light_keys(self) -> List[str]:
        return ["light_base"] if self.prefiltered is not None else [f"light{i}" for i in range(len(self.light_leaves))]

    def light_backward_overlapped(self, light_ready) -> None:
        """Single-rank form of begin_light_all_reduce: run the mips' backward on a side stream as soon as the texture
        gradients are final (`light_ready`, recorded before the blend backward), then make the main stream wait for
        it — the 0.23 ms of filter backward run under the blend backward instead of after it."""
        if getattr(self, "_comm", None) is None:
            self._comm = torch.cuda.Stream(device=self.flat_grad.device)
        with torch.cuda.stream(self._comm):
            self._comm.wait_event(light_ready)
            self.prefiltered.backward(self.light_base.grad, accumulate=True)
            done = torch.cuda.Event()
            done.record()
        torch.cuda.current_stream().wait_event(done)


def training_step(params: GaussianParams, cam, light, brdf_lut, rays, gt_image, background, gi: Dict,
                  metallic=True, gamma=True, tone=False, indirect=True, loss_scale: float = 1.0,
                  fused: bool = True, gt_ready=None, light_ready=None, build_light: bool = True,
                  finish_light: bool = True, brdf_tv_weight: float = 0.0, env_tv_weight: float = 0.0,
                  radiance: bool = False) -> torch.Tensor:
    """forward + loss + backward for ONE view; gradients accumulate into params.flat_grad.

    When `light` is params.prefiltered (the light given as its trainable base cubemap), build_light rebuilds the mips
    first (train.py:340) and finish_light back-propagates the texture gradients into the base afterwards, on a side
    stream under the blend backward; a multi-view step builds once, finishes once (multi_view_step).
    brdf_tv_weight / env_tv_weight: the two smoothness priors of the reference's PBR-stage loss (train.py:388-420,
    defaults there 1.0 and 0.01); the env-map one needs the base cubemap (light_base=).
    radiance: also produce the SH radiance image and the blended position in the fused frame (the PBR stage reads
    neither, so by default they are skipped: gigs.frame.pbr_frame_step).

    fused=True (default) runs the frame as two C-ABI calls (gigs.frame: activations, rasterizer, deferred shading /
    SSR / loss kernels and the material-only backward, ~20 kernel launches). fused=False runs the same frame
    operator by operator through the drop-in modules and autograd (~200 launches): same result to float rounding,
    kept as the reference-shaped path and as the parity check of the fused one."""
    pre = params.prefiltered if (params.prefiltered is not None and light is params.prefiltered) else None
    if pre is not None and build_light:
        pre.build()
    if fused:
        from .frame import pbr_frame_step
        params.mark_dirty(["albedo", "roughness", "metallic"] + params.light_keys())
        own_event = pre is not None and finish_light and light_ready is None
        if own_event:
            light_ready = torch.cuda.Event()
        env_loss = None
        if pre is not None and env_tv_weight:
            # the env-map prior only needs the base cubemap: it goes FIRST, so that its accumulation into base.grad is
            # ordered before the mips' backward / the all-reduce that later run on the side stream
            from .light import env_tv_fused
            env_loss = torch.empty(1, dtype=torch.float32, device=params.flat_grad.device)
            env_tv_fused(params.light_base.detach(), params.env_dirs(), env_tv_weight * loss_scale,
                         grad_base=params.light_base.grad, loss_out=env_loss)
        loss = pbr_frame_step(params, cam, light, brdf_lut, rays, gt_image, background, gi, metallic=metallic,
                              gamma=gamma, tone=tone, indirect=indirect, loss_scale=loss_scale, gt_ready=gt_ready,
                              light_ready=light_ready, brdf_tv_weight=brdf_tv_weight, radiance=radiance)
        if own_event:
            params.light_backward_overlapped(light_ready)
        return loss if env_loss is None else loss + env_loss[0]
    if gt_ready is not None:
        torch.cuda.current_stream().wait_event(gt_ready)
    params.mark_dirty(None)
    g = params.activated()
    res = pbr_forward(cam, g, light, brdf_lut, rays, background, indirect=indirect, metallic=metallic, tone=tone,
                      gamma=gamma, gi=gi)
    loss = pbr_loss(res, gt_image, brdf_tv_weight=brdf_tv_weight) * loss_scale
    if pre is not None and env_tv_weight:
        from .light import env_tv_loss
        loss = loss + (env_tv_weight * loss_scale) * env_tv_loss(params.light_base, params.env_dirs())
    loss.backward()
    if pre is not None and finish_light:
        pre.backward(params.light_base.grad, accumulate=True)
    return loss.detach()


def _framework_first_stage_loss(res: Dict, gt_image: torch.Tensor, lambda_dssim: float, normal_weight: float,
                                normal_tv_weight: float) -> torch.Tensor:
    """The first-stage loss in framework ops exactly as train.py:318-328 + utils/loss_utils.py:54-100 write it: the
    parity check of the fused loss kernels (fused_losses=False), not a fallback."""
    image, nm, nd = res["render"], res["normal_map"], res["normal_map_from_depth"]
    ch = image.shape[0]
    w1 = torch.tensor([math.exp(-((x - 5) ** 2) / float(2 * 1.5 ** 2)) for x in range(11)])
    w1 = (w1 / w1.sum()).unsqueeze(1)
    win = w1.mm(w1.t()).float()[None, None].expand(ch, 1, 11, 11).contiguous().to(image.device)
    a, b = image[None], gt_image[None]
    mu1, mu2 = F.conv2d(a, win, padding=5, groups=ch), F.conv2d(b, win, padding=5, groups=ch)
    mu1_sq, mu2_sq, mu12 = mu1.pow(2), mu2.pow(2), mu1 * mu2
    s11 = F.conv2d(a * a, win, padding=5, groups=ch) - mu1_sq
    s22 = F.conv2d(b * b, win, padding=5, groups=ch) - mu2_sq
    s12 = F.conv2d(a * b, win, padding=5, groups=ch) - mu12
    ssim = (((2 * mu12 + 0.01 ** 2) * (2 * s12 + 0.03 ** 2)) / ((mu1_sq + mu2_sq + 0.01 ** 2) * (s11 + s22 + 0.03 ** 2))).mean()
    loss = (1.0 - lambda_dssim) * torch.abs(image - gt_image).mean() + lambda_dssim * (1.0 - ssim)
    mask = res["normal_from_depth_mask"]
    loss = loss + normal_weight * F.l1_loss(nm[:, mask], nd[:, mask])
    wh = torch.exp(-(gt_image[:, 1:, :] - gt_image[:, :-1, :]).abs().mean(dim=0, keepdim=True))
    ww = torch.exp(-(gt_image[:, :, 1:] - gt_image[:, :, :-1]).abs().mean(dim=0, keepdim=True))
    tv = ((nm[:, 1:, :] - nm[:, :-1, :]).pow(2) * wh).mean() + ((nm[:, :, 1:] - nm[:, :, :-1]).pow(2) * ww).mean()
    return loss + normal_tv_weight * tv


def first_stage_step(params: GaussianParams, cam, gt_image, background, gi: Dict, lambda_dssim: float = 0.2,
                     normal_weight: float = 1.0, normal_tv_weight: float = 1.0, loss_scale: float = 1.0,
                     fused_losses: bool = True, stats=None, fused: bool = False, gt_ready=None):
    """One view of the FIRST training stage (iteration <= pbr_iteration, train.py:266-328): render() with
    derive_normal, loss = (1 - lambda) L1 + lambda (1 - SSIM) + normal L1 inside the normal-from-depth mask + edge-aware
    TV of the normal map, backward through the rasterizer's general backward (all 10 parameter groups receive
    gradients). The image and normal losses run as the fused kernels of csrc/loss.cu (gigs.losses).
    `stats` (gigs.densify.DensifyState) receives this view's densification statistics (train.py:489-495).
    fused=True runs the whole view as two C-ABI calls (gigs.frame.stage1_frame_step, csrc/stage1.cu: getters inside
    preprocess, post-processing / loss / backward kernels, gradients written straight into the leaves' gradient
    tensors); the result dict then holds only "radii" and "viewspace_grad".
    Returns (loss, render result)."""
    from .renderer import render
    from . import losses
    params.mark_dirty(None)
    if fused:
        from .frame import stage1_frame_step
        loss, g2d, radii = stage1_frame_step(params, cam, gt_image, background, lambda_dssim, normal_weight,
                                             normal_tv_weight, loss_scale, gt_ready=gt_ready)
        if stats is not None:
            # the statistics are per-view norms of the UNSCALED loss's screen-space gradient (train.py:489-495): a K-view
            # step renders each view with loss_scale = 1/K, which must not shrink them against the fixed threshold
            stats.add_view(g2d, radii, 1.0 / loss_scale)
        return loss.clone(), {"radii": radii, "viewspace_grad": g2d}
    g = params.activated()
    res = render(cam, g, background, derive_normal=True, **gi)
    if fused_losses:
        loss = losses.l1_ssim_loss(res["render"], gt_image, lambda_dssim)
        loss = loss + losses.normal_loss(res["normal_map"], res["normal_map_from_depth"], res["normal_from_depth_mask"],
                                         gt_image, normal_weight, normal_tv_weight)
    else:
        loss = _framework_first_stage_loss(res, gt_image, lambda_dssim, normal_weight, normal_tv_weight)
    (loss * loss_scale).backward()
    if stats is not None:
        stats.add_view(res["viewspace_points"].grad, res["radii"], 1.0 / loss_scale)
    return loss.detach() * loss_scale, res


def multi_view_step(params: GaussianParams, cams: List, light: Light, brdf_lut, rays_of, gts: List, background,
                    gi: Dict, rank: int = 0, world: int = 1, **kw) -> torch.Tensor:
    """K-view step, views sharded round-robin over `world` ranks (BASELINE C4): loss = mean over
    the K views; each rank back-propagates its own views, then ONE all-reduce(sum) of the flat gradient buffer
    (268 B per Gaussian + the light texels). With world == 1 this is plain gradient accumulation."""
    K = len(cams)
    params.zero_grad(fused_only=bool(kw.get("fused", True)))
    total = torch.zeros((), device=params.flat_grad.device)
    mine = list(range(rank, K, world))
    fused = bool(kw.get("fused", True))
    pre = params.prefiltered if (params.prefiltered is not None and light is params.prefiltered) else None
    for k in mine:
        ev = None
        last = k == mine[-1]
        if world > 1 and fused and last:
            ev = torch.cuda.Event()   # light gradients are final after the last local view's deferred backward
        # the mips are built once per step and back-propagated once, after the last local view (both are linear)
        total = total + training_step(params, cams[k], light, brdf_lut, rays_of(cams[k]), gts[k], background, gi,
                                      loss_scale=1.0 / K, light_ready=ev, build_light=(k == mine[0]),
                                      finish_light=(last and ev is None), **kw)
        if ev is not None:
            params.begin_light_all_reduce(ev)
    if world > 1:
        import torch.distributed as dist
        params.all_reduce_grads(fused_only=bool(kw.get("fused", True)))
        dist.all_reduce(total, op=dist.ReduceOp.SUM)
    return total


def multi_view_first_stage_step(params: GaussianParams, cams: List, gts: List, background, gi: Dict, rank: int = 0,
                                world: int = 1, stats=None, **kw) -> torch.Tensor:
    """K-view step of the FIRST training stage, views sharded round-robin over `world` ranks: every parameter group
    receives gradients there, so the exchange is the whole flat buffer — 268 B per Gaussian — in one
    gigs_peer_allreduce launch (GaussianParams(peer=True)) or one NCCL all-reduce. loss = mean over the K views.
    `stats` accumulates the densification statistics of the LOCAL views (per-view norms: they are summed over ranks with
    reduce_densify_stats, not derived from the reduced gradient — SURVEY §8e)."""
    K = len(cams)
    params.zero_grad()
    total = torch.zeros((), device=params.flat_grad.device)
    for k in range(rank, K, world):
        loss, _ = first_stage_step(params, cams[k], gts[k], background, gi, loss_scale=1.0 / K, stats=stats,
                                   fused=bool(kw.get("fused", True)),
                                   **{a: b for a, b in kw.items() if a != "fused"})
        total = total + loss
    if world > 1:
        import torch.distributed as dist
        params.mark_dirty(None)
        params.all_reduce_grads(fused_only=False)
        dist.all_reduce(total, op=dist.ReduceOp.SUM)
    return total


def reduce_densify_stats(stats, world: int) -> None:
    """Sum the per-view densification statistics over the ranks (xyz_gradient_accum*, denom) and take the maximum of
    the per-view maxima (xyz_gradient_accum_abs_max, max_radii2D): afterwards every rank holds what a single process
    that had rendered all K views would hold (scene/gaussian_model.py:933-945)."""
    if world <= 1:
        return
    import torch.distributed as dist
    for t in (stats.xyz_gradient_accum, stats.xyz_gradient_accum_abs, stats.denom):
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    for t in (stats.xyz_gradient_accum_abs_max, stats.max_radii2D):
        dist.all_reduce(t, op=dist.ReduceOp.MAX)


def shard_views(n_views: int, rank: int, world: int) -> List[int]:
    """Camera-sharded eval / relight sweep (BASELINE C5): independent views, round-robin, no data-path collective."""
    return list(range(rank, n_views, world))
