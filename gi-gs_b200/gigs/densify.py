"""Densification, pruning and opacity reset of the first training stage, behind the names of
/root/reference/scene/gaussian_model.py: add_densification_stats (:933-945), densify_and_prune (:905-931),
reset_opacity (:467-472).

B200-first shape of the work: the reference rebuilds the model in four rounds of torch.cat / boolean-mask copies over
10 parameter tensors and 20 Adam moment tensors. The result is a pure function of three per-point decisions (clone?
split? prune?), so the decisions are taken on [P]-sized vectors and ONE launch of gigs_densify_gather writes the new
parameters and moments (csrc/densify.cu). Row order, values and optimiser state equal the reference's outcome
(tests/golden/densify_ref.npz is produced by the reference's own methods).
"""
import ctypes as C
from typing import Dict, Optional

import torch

from . import _lib
from .step import PARAM_KEYS, PARAM_WIDTH, GaussianParams


class DensifyState:
    """xyz_gradient_accum, xyz_gradient_accum_abs, xyz_gradient_accum_abs_max, denom [P,1] and max_radii2D [P]
    (GaussianModel.training_setup, scene/gaussian_model.py:320-323,:53)."""

    def __init__(self, P: int, device):
        self.device = device
        self.reset(P)

    def reset(self, P: int) -> None:
        z = lambda *s: torch.zeros(*s, dtype=torch.float32, device=self.device)
        self.xyz_gradient_accum, self.xyz_gradient_accum_abs = z(P, 1), z(P, 1)
        self.xyz_gradient_accum_abs_max, self.denom, self.max_radii2D = z(P, 1), z(P, 1), z(P)

    def add_view(self, grad2D: torch.Tensor, radii: torch.Tensor, grad_scale: float = 1.0) -> None:
        """train.py:489-495: max_radii2D[vis] = max(., radii[vis]); add_densification_stats(viewspace_points, vis).
        grad_scale undoes a loss scale the view was rendered with (multi-view steps use 1/K)."""
        L = _lib.load()
        P = radii.shape[0]
        if grad2D is None:
            raise RuntimeError("DensifyState.add_view: the screen-space points carry no gradient (run backward first)")
        g = grad2D.detach().float().contiguous()
        if grad_scale != 1.0:
            g = g * float(grad_scale)
        r = radii.to(torch.int32).contiguous()
        with torch.cuda.device(g.device):
            _lib.check(L.gigs_densify_stats(P, r.data_ptr(), g.data_ptr(), g.shape[1], self.xyz_gradient_accum.data_ptr(),
                                            self.xyz_gradient_accum_abs.data_ptr(),
                                            self.xyz_gradient_accum_abs_max.data_ptr(), self.denom.data_ptr(),
                                            self.max_radii2D.data_ptr(), torch.cuda.current_stream().cuda_stream),
                       "gigs_densify_stats")


def plan(params: GaussianParams, state: DensifyState, max_grad: float, min_opacity: float, extent: float,
         max_screen_size: Optional[int], percent_dense: float = 0.01, N: int = 2) -> Dict:
    """The three per-point decisions of densify_and_prune and the source map of the model they lead to.
    Row order of the reference's outcome: surviving original points (not split), clones, split children (the N copies
    block-repeated: all first children, then all second children), each filtered by the final prune."""
    L = params.leaves
    P = params.P
    dev = L["xyz"].device
    grads = state.xyz_gradient_accum / state.denom
    grads[grads.isnan()] = 0.0
    grads_abs = state.xyz_gradient_accum_abs / state.denom
    grads_abs[grads_abs.isnan()] = 0.0
    ratio = (torch.norm(grads, dim=-1) >= max_grad).float().mean()
    Q = torch.quantile(grads_abs.reshape(-1), 1 - ratio)
    scaling_max = torch.exp(L["log_scale"].detach()).max(dim=1).values
    hot = torch.logical_or(torch.norm(grads, dim=-1) >= max_grad, torch.norm(grads_abs, dim=-1) >= Q)
    small = scaling_max <= percent_dense * extent
    clone = torch.logical_and(hot, small)                                   # densify_and_clone :787-791
    # densify_and_split :743-751 looks at the model AFTER the clones were appended, with zero-padded gradients
    split = torch.logical_and(hot, ~small)
    pad_hot = torch.logical_or(torch.zeros((), device=dev) >= max_grad, torch.zeros((), device=dev) >= Q)
    clone_idx = clone.nonzero().squeeze(1)
    # a clone has its source's scaling (<= threshold), so it can never be split even when the zero-padded gradient passes
    assert not bool(pad_hot) or bool((scaling_max[clone_idx] <= percent_dense * extent).all())
    split_idx = split.nonzero().squeeze(1)
    keep_idx = (~split).nonzero().squeeze(1)
    src = torch.cat([keep_idx, clone_idx, split_idx.repeat(N)])
    kind = torch.cat([torch.zeros_like(keep_idx), torch.ones_like(clone_idx), torch.full_like(split_idx.repeat(N), 2)])
    # final prune (:920-926) on the candidates' own values: opacity is copied; a split child's scaling is s / (0.8 N);
    # max_radii2D was zeroed by densification_postfix (:706), so the screen-size test can never fire — kept literally
    opacity = torch.sigmoid(L["opacity"].detach())[src, 0]
    prune = opacity < min_opacity
    if max_screen_size:
        cand_scale = torch.where(kind == 2, scaling_max[src] / (0.8 * N), scaling_max[src])
        big_vs = torch.zeros_like(prune)                                    # max_radii2D == 0 > max_screen_size: never
        big_ws = cand_scale > 0.1 * extent
        prune = torch.logical_or(torch.logical_or(prune, big_vs), big_ws)
    sel = ~prune
    return dict(src=src[sel].to(torch.int32).contiguous(), kind=kind[sel].to(torch.int8).contiguous(),
                n_clone=int(clone_idx.numel()), n_split=int(split_idx.numel()), n_candidates=int(src.numel()),
                cand_kind=kind, cand_keep=sel, P_before=P)


def densify_and_prune(params: GaussianParams, optimizer, state: DensifyState, max_grad: float, min_opacity: float,
                      extent: float, max_screen_size: Optional[int], percent_dense: float = 0.01, N: int = 2,
                      noise: Optional[torch.Tensor] = None, generator: Optional[torch.Generator] = None) -> Dict:
    """GaussianModel.densify_and_prune. `noise` ([n_candidates, 3] standard normals in candidate order: one row per
    original point (unused), per clone, per split child) makes the sampled positions reproducible; by default it is
    drawn on the device. Rebuilds `params` in place, moves the optimiser state (moments of kept points kept, of new
    points zero, step counts unchanged) and resets the densification statistics, as the reference does."""
    Lb = _lib.load()
    pl = plan(params, state, max_grad, min_opacity, extent, max_screen_size, percent_dense, N)
    dev = params.flat_grad.device
    n_out = int(pl["src"].numel())
    if noise is None:
        noise_out = torch.randn(n_out, 3, device=dev, generator=generator)
    else:
        noise_out = noise.to(dev).float()[pl["cand_keep"]].contiguous()
    old = params.leaves
    new_leaves, new_state, keep = {}, {}, []
    arr = (_lib.GigsDensifyGroup * len(PARAM_KEYS))()
    for i, k in enumerate(PARAM_KEYS):
        w = PARAM_WIDTH[k]
        t = old[k].detach()
        dst = torch.empty((n_out,) + tuple(t.shape[1:]), dtype=torch.float32, device=dev)
        name = optimizer._name_of(k) if optimizer is not None else None
        st = optimizer.adam.state.get(name) if optimizer is not None else None
        a = arr[i]
        a.src, a.dst, a.width = t.data_ptr(), dst.data_ptr(), w
        a.role = 1 if k == "xyz" else (2 if k == "log_scale" else 0)
        if st is not None:
            dm, dv = torch.empty_like(dst), torch.empty_like(dst)
            a.src_exp_avg, a.src_exp_avg_sq = st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr()
            a.dst_exp_avg, a.dst_exp_avg_sq = dm.data_ptr(), dv.data_ptr()
            new_state[k] = dict(exp_avg=dm, exp_avg_sq=dv)
        new_leaves[k] = dst
        keep.append((t, dst))
    if n_out:
        with torch.cuda.device(dev):
            _lib.check(Lb.gigs_densify_gather(n_out, pl["src"].data_ptr(), pl["kind"].data_ptr(), noise_out.data_ptr(),
                                              old["log_scale"].data_ptr(), old["rot"].data_ptr(), float(0.8 * N),
                                              len(PARAM_KEYS), arr, torch.cuda.current_stream().cuda_stream),
                       "gigs_densify_gather")
    params.rebuild(new_leaves)
    if optimizer is not None:
        optimizer.rebind(new_state)
    state.reset(n_out)
    return dict(P_before=pl["P_before"], P_after=n_out, n_clone=pl["n_clone"], n_split=pl["n_split"])


def reset_opacity(params: GaussianParams, optimizer) -> None:
    """GaussianModel.reset_opacity (:467-472): opacity = inverse_sigmoid(min(sigmoid(opacity), 0.01)), its Adam moments
    zeroed (replace_tensor_to_optimizer :580-592)."""
    with torch.no_grad():
        o = params.leaves["opacity"]
        x = torch.min(torch.sigmoid(o), torch.ones_like(o) * 0.01)
        o.copy_(torch.log(x / (1 - x)))
    if optimizer is not None:
        st = optimizer.adam.state.get(optimizer._name_of("opacity"))
        if st is not None:
            st["exp_avg"].zero_()
            st["exp_avg_sq"].zero_()
