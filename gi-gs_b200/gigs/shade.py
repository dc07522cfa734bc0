"""Split-sum deferred shading: the host-side mirror of the reference's `pbr_shading`
(/root/reference/pbr/shade.py:104-237) backed by ONE fused CUDA kernel forward and one backward
(gigs_shade_forward / gigs_shade_backward, include/gigs_b200.h) instead of ~25 elementwise launches and
three third-party texture kernels. Same argument names / meanings / result keys as the reference.
"""
import ctypes as C
import math
from typing import Dict, List, Optional

import torch

from . import _lib
from ._lib import GigsShade, check, ptr

_L = _lib.load()

MIN_ROUGHNESS = 0.08  # pbr/light.py:88-89
MAX_ROUGHNESS = 0.5


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _chw(t: Optional[torch.Tensor], c: int) -> Optional[torch.Tensor]:
    """[H,W,c] (usually a permuted view of a CHW map) -> contiguous [c,H,W] without a copy when possible."""
    if t is None:
        return None
    if t.dim() == 4:
        t = t[0]
    return t.permute(2, 0, 1).contiguous()


class _Shade(torch.autograd.Function):
    @staticmethod
    def forward(ctx, albedo, roughness, metallic, diffuse_tex, normals, view_dirs, occlusion, mask, background, lut,
                tone, gamma, rmin, rmax, *spec):
        dev = albedo.device
        _, H, W = albedo.shape
        f32 = dict(dtype=torch.float32, device=dev)
        render = torch.empty((3, H, W), **f32)
        diffuse_rgb = torch.empty((3, H, W), **f32)
        specular_rgb = torch.empty((3, H, W), **f32)
        diffuse_light = torch.empty((3, H, W), **f32)
        a = _Shade._fill(albedo, roughness, metallic, diffuse_tex, normals, view_dirs, occlusion, mask, background,
                         lut, tone, gamma, rmin, rmax, spec)
        a.render_rgb = render.data_ptr(); a.diffuse_rgb = diffuse_rgb.data_ptr()
        a.specular_rgb = specular_rgb.data_ptr(); a.diffuse_light = diffuse_light.data_ptr()
        with torch.cuda.device(dev):
            check(_L.gigs_shade_forward(C.byref(a)), "gigs_shade_forward")
        ctx.save_for_backward(albedo, roughness, metallic, diffuse_tex, normals, view_dirs, occlusion, mask,
                              background, lut, *spec)
        ctx.cfg = (tone, gamma, rmin, rmax)
        ctx.set_materialize_grads(False)
        ctx.mark_non_differentiable(diffuse_light)
        return render, diffuse_rgb, specular_rgb, diffuse_light

    @staticmethod
    def _fill(albedo, roughness, metallic, diffuse_tex, normals, view_dirs, occlusion, mask, background, lut, tone,
              gamma, rmin, rmax, spec) -> GigsShade:
        _, H, W = albedo.shape
        a = GigsShade()
        a.W = W; a.H = H
        a.n_spec_levels = len(spec)
        for i, s in enumerate(spec):
            a.spec_res[i] = s.shape[1]
            a.spec[i] = s.data_ptr()
        a.diffuse_res = diffuse_tex.shape[1]
        a.diffuse = diffuse_tex.data_ptr()
        a.brdf_lut = lut.data_ptr()
        a.lut_res = lut.shape[-2]
        a.tone = int(bool(tone)); a.gamma = int(bool(gamma))
        a.has_metallic = int(metallic is not None); a.has_occlusion = int(occlusion is not None)
        a.min_roughness = float(rmin); a.max_roughness = float(rmax)
        a.normals = ptr(normals); a.view_dirs = ptr(view_dirs); a.albedo = ptr(albedo); a.roughness = ptr(roughness)
        a.metallic = ptr(metallic); a.occlusion = ptr(occlusion); a.mask = ptr(mask); a.background = ptr(background)
        a.stream = _stream()
        return a

    @staticmethod
    def backward(ctx, g_render, g_diffuse, g_specular, _g_light=None):
        (albedo, roughness, metallic, diffuse_tex, normals, view_dirs, occlusion, mask, background, lut,
         *spec) = ctx.saved_tensors
        tone, gamma, rmin, rmax = ctx.cfg
        dev = albedo.device
        a = _Shade._fill(albedo, roughness, metallic, diffuse_tex, normals, view_dirs, occlusion, mask, background,
                         lut, tone, gamma, rmin, rmax, spec)
        gs = [None if g is None else g.contiguous() for g in (g_render, g_diffuse, g_specular)]
        a.g_render_rgb = ptr(gs[0]); a.g_diffuse_rgb = ptr(gs[1]); a.g_specular_rgb = ptr(gs[2])
        g_albedo = torch.empty_like(albedo)
        g_rough = torch.empty_like(roughness)
        g_metal = torch.empty_like(metallic) if metallic is not None else None
        need = ctx.needs_input_grad
        g_dtex = torch.zeros_like(diffuse_tex) if need[3] else None
        g_spec = [torch.zeros_like(s) if need[14 + i] else None for i, s in enumerate(spec)]
        a.g_albedo = g_albedo.data_ptr(); a.g_roughness = g_rough.data_ptr(); a.g_metallic = ptr(g_metal)
        a.g_diffuse_tex = ptr(g_dtex)
        for i, t in enumerate(g_spec):
            a.g_spec[i] = ptr(t)
        with torch.cuda.device(dev):
            check(_L.gigs_shade_backward(C.byref(a)), "gigs_shade_backward")
        return (g_albedo, g_rough, g_metal, g_dtex, None, None, None, None, None, None, None, None, None, None,
                *g_spec)


def pbr_shading(light, normals: torch.Tensor, view_dirs: torch.Tensor, albedo: torch.Tensor, roughness: torch.Tensor,
                mask: torch.Tensor, tone: bool = False, gamma: bool = False, occlusion: Optional[torch.Tensor] = None,
                metallic: Optional[torch.Tensor] = None, brdf_lut: Optional[torch.Tensor] = None,
                background: Optional[torch.Tensor] = None) -> Dict:
    """Same contract as /root/reference/pbr/shade.py:104-237: HWC inputs ([H,W,3] / [H,W,1]), result dict with
    render_rgb / diffuse_rgb / specular_rgb / diffuse_light as [H,W,3]. `light` needs `.diffuse` [6,r,r,3] and
    `.specular` (list of [6,r_i,r_i,3]); gradients flow to albedo, roughness, metallic and the light textures
    (normals / view_dirs / occlusion are detached by the reference's caller, train.py:343-351)."""
    if brdf_lut is None:
        raise RuntimeError("brdf_lut is required")
    spec: List[torch.Tensor] = [s.contiguous() for s in light.specular]
    diffuse_tex = light.diffuse.contiguous()
    rmin = getattr(light, "MIN_ROUGHNESS", MIN_ROUGHNESS)
    rmax = getattr(light, "MAX_ROUGHNESS", MAX_ROUGHNESS)
    n = _chw(normals.detach(), 3); v = _chw(view_dirs.detach(), 3)
    alb = _chw(albedo, 3); rg = _chw(roughness, 1)
    met = _chw(metallic, 1) if metallic is not None else None
    occ = _chw(occlusion.detach(), 1) if occlusion is not None else None
    m = mask
    if m.dim() == 3:
        m = m[..., 0]
    m = m.to(torch.uint8).contiguous()
    bgt = _chw(background, 3) if background is not None else None
    lut = brdf_lut.contiguous()
    render, d_rgb, s_rgb, d_light = _Shade.apply(alb, rg, met, diffuse_tex, n, v, occ, m, bgt, lut, tone, gamma,
                                                 rmin, rmax, *spec)
    return {"diffuse_light": d_light.permute(1, 2, 0), "render_rgb": render.permute(1, 2, 0),
            "diffuse_rgb": d_rgb.permute(1, 2, 0), "specular_rgb": s_rgb.permute(1, 2, 0)}


class Light:
    """Minimal stand-in for CubemapLight's shading-side interface (pbr/light.py:84-152): the textures the
    fused shade pass consumes. build_mips (GGX prefilter) is the adjacent §8f-1 row and is not reproduced."""
    MIN_ROUGHNESS = MIN_ROUGHNESS
    MAX_ROUGHNESS = MAX_ROUGHNESS

    def __init__(self, specular: List[torch.Tensor], diffuse: torch.Tensor):
        self.specular = specular
        self.diffuse = diffuse

    def get_mip(self, roughness: torch.Tensor) -> torch.Tensor:
        n = len(self.specular)
        return torch.where(
            roughness < self.MAX_ROUGHNESS,
            (torch.clamp(roughness, self.MIN_ROUGHNESS, self.MAX_ROUGHNESS) - self.MIN_ROUGHNESS)
            / (self.MAX_ROUGHNESS - self.MIN_ROUGHNESS) * (n - 2),
            (torch.clamp(roughness, self.MAX_ROUGHNESS, 1.0) - self.MAX_ROUGHNESS) / (1.0 - self.MAX_ROUGHNESS) + n - 2)


from .scene import make_brdf_lut  # noqa: E402,F401  (synthetic-input generator: lives with the other generators, no CUDA library needed)
