"""Per-view render glue: mirror of the reference's `gaussian_renderer.render`
(/root/reference/gaussian_renderer/__init__.py:30-220) and of the PBR-stage part of its training step
(/root/reference/train.py:266-422), written against the drop-in `diff_gaussian_rasterization` package.

The reference's render() takes a Camera / GaussianModel pair; here the same steps take the activated tensors
(scene.activate = the GaussianModel getters) so the file has no dependency on the trainer's classes. The
post-processing (masks, normalisation, 3x3 medians, rotation to view space) is kept operation for operation.
"""
from typing import Dict, Optional

import torch
import torch.nn.functional as F

from diff_gaussian_rasterization import (GaussianRasterizationSettings, GaussianRasterizer, Gaussian_SSR,
                                         median_blur3x3)

from .shade import pbr_shading


def render(cam, g: Dict, bg_color: torch.Tensor, scaling_modifier: float = 1.0,
           override_color: Optional[torch.Tensor] = None, inference: bool = False, pad_normal: bool = False,
           derive_normal: bool = False, radius: float = 0.8, bias: float = 0.01, thick: float = 0.05,
           delta: float = 0.0625, step: int = 16, start: int = 8, debug: bool = False) -> Dict:
    """gaussian_renderer/__init__.py:30-220. `g` holds means3D, opacity, normal, albedo, roughness, metallic,
    scales, rotations, shs, sh_degree (already activated)."""
    means3D = g["means3D"]
    screenspace_points = torch.zeros_like(means3D, dtype=means3D.dtype, requires_grad=True, device=means3D.device) + 0
    try:
        screenspace_points.retain_grad()
    except Exception:
        pass
    raster_settings = GaussianRasterizationSettings(
        image_height=int(cam.image_height), image_width=int(cam.image_width), tanfovx=cam.tanfovx,
        tanfovy=cam.tanfovy, radius=radius, bias=bias, thick=thick, delta=delta, step=step, start=start, bg=bg_color,
        scale_modifier=scaling_modifier, viewmatrix=cam.world_view_transform, projmatrix=cam.full_proj_transform,
        sh_degree=g["sh_degree"], campos=cam.camera_center, prefiltered=False, debug=debug, inference=inference,
        argmax_depth=False)
    rasterizer = GaussianRasterizer(raster_settings=raster_settings)
    shs, colors_precomp = (g["shs"], None) if override_color is None else (None, override_color)
    (rendered_image, radii, opacity_map, depth_map, normal_map_from_depth, normal_map, occlusion_map, albedo_map,
     roughness_map, metallic_map, out_normal_view, depth_pos) = rasterizer(
        means3D=means3D, means2D=screenspace_points, opacities=g["opacity"], normal=g["normal"], shs=shs,
        colors_precomp=colors_precomp, albedo=g["albedo"], roughness=g["roughness"], metallic=g["metallic"],
        scales=g["scales"], rotations=g["rotations"], cov3D_precomp=None, derive_normal=derive_normal)

    normal_from_depth_mask = (normal_map_from_depth != 0).all(0)
    normal_mask = (normal_map != 0).all(0, keepdim=True)
    if pad_normal:
        opacity_map = torch.where(opacity_map < 0.004, torch.zeros_like(opacity_map), opacity_map)
        opacity_map = torch.where(opacity_map > 1.0 - 0.004, torch.ones_like(opacity_map), opacity_map)
        normal_bg = torch.tensor([0.0, 0.0, 1.0], device=normal_map.device)
        normal_map = normal_map * opacity_map + (1.0 - opacity_map) * normal_bg[:, None, None]
        mask_from_depth = (normal_map_from_depth == 0.0).all(0, keepdim=True).float()
        normal_map_from_depth = normal_map_from_depth * (1.0 - mask_from_depth) + mask_from_depth * normal_bg[:, None, None]

    normal_map_from_depth = torch.where(torch.norm(normal_map_from_depth, dim=0, keepdim=True) > 0,
                                        F.normalize(normal_map_from_depth, dim=0, p=2), normal_map_from_depth)
    normal_map = torch.where(torch.norm(normal_map, dim=0, keepdim=True) > 0, F.normalize(normal_map, dim=0, p=2),
                             normal_map)
    normal_map = median_blur3x3(normal_map)

    R = cam.world_view_transform[:3, :3]  # rotation only
    normals_view = (normal_map.permute(1, 2, 0) @ R).permute(2, 0, 1)
    normals_view = -normals_view

    out_normal_view = torch.where(torch.norm(out_normal_view, dim=0, keepdim=True) > 0,
                                  F.normalize(out_normal_view, dim=0, p=2), out_normal_view)
    out_normal_view = median_blur3x3(out_normal_view)

    return {
        "render": rendered_image, "viewspace_points": screenspace_points, "visibility_filter": radii > 0,
        "radii": radii, "opacity_map": opacity_map, "depth_map": depth_map,
        "normal_map_from_depth": normal_map_from_depth, "normal_from_depth_mask": normal_from_depth_mask,
        "normal_map": normals_view, "normal_mask": normal_mask, "albedo_map": albedo_map,
        "roughness_map": roughness_map, "metallic_map": metallic_map, "occlusion_map": occlusion_map,
        "out_normal_view": out_normal_view, "depth_pos": depth_pos,
    }


def srgb_to_linear(srgb: torch.Tensor) -> torch.Tensor:
    # train.py:70-75
    linear0 = 25 / 323 * srgb
    linear1 = ((srgb + 0.055) / 1.055) ** 2.4
    return torch.where(srgb <= 0.04045, linear0, linear1)


def linear_to_srgb(linear: torch.Tensor) -> torch.Tensor:
    # train.py:54-61
    eps = torch.finfo(torch.float32).eps
    srgb0 = 323 / 25 * linear
    srgb1 = (211 * torch.clamp(linear, min=eps) ** (5 / 12) - 11) / 200
    return torch.where(linear <= 0.0031308, srgb0, srgb1)


def pbr_forward(cam, g: Dict, light, brdf_lut, canonical_rays, background, indirect=True, metallic=True, tone=False,
                gamma=True, gi=None, inference=False) -> Dict:
    """G-buffer + SSAO + split-sum shading + SSR for one view: train.py:266-384 (training) and
    render.py:202-333 (eval) share this sequence."""
    gi = gi or {}
    rr = render(cam, g, background, pad_normal=False, derive_normal=True, inference=inference, **gi)
    H, W = cam.image_height, cam.image_width
    c2w = torch.inverse(cam.world_view_transform.T)
    albedo_map, metallic_map = rr["albedo_map"], rr["metallic_map"]
    roughness_map = rr["roughness_map"] * (1.0 - 0.04) + 0.04
    view_dirs = -((F.normalize(canonical_rays[:, None, :], p=2, dim=-1) * c2w[None, :3, :3]).sum(dim=-1)
                  .reshape(H, W, 3))
    if indirect:
        occlusion = rr["occlusion_map"].permute(1, 2, 0)
    else:
        occlusion = torch.ones_like(roughness_map).permute(1, 2, 0)
    normal_mask = rr["normal_mask"]
    pbr = pbr_shading(light=light, normals=rr["normal_map"].permute(1, 2, 0).detach(), view_dirs=view_dirs,
                      mask=normal_mask.permute(1, 2, 0), albedo=albedo_map.permute(1, 2, 0),
                      roughness=roughness_map.permute(1, 2, 0),
                      metallic=metallic_map.permute(1, 2, 0) if metallic else None, tone=tone, gamma=gamma,
                      occlusion=occlusion.detach(), brdf_lut=brdf_lut)
    render_direct = pbr["render_rgb"].permute(2, 0, 1)
    render_direct = torch.where(normal_mask, render_direct, background[:, None, None])
    tanfovx, tanfovy = cam.tanfovx, cam.tanfovy
    ssr_mod = Gaussian_SSR(tanfovx, tanfovy, W, H, gi.get("radius", 0.8), gi.get("bias", 0.01), gi.get("thick", 0.05),
                           gi.get("delta", 0.0625), gi.get("step", 16), gi.get("start", 8))
    if metallic:
        F0 = (1.0 - metallic_map) * 0.04 + albedo_map * metallic_map
    else:
        F0 = torch.ones_like(albedo_map) * 0.04
        metallic_map = torch.zeros_like(roughness_map)
    linear_rgb = srgb_to_linear(render_direct)
    (IRR, _) = ssr_mod(rr["out_normal_view"].detach(), rr["depth_pos"].detach(), linear_rgb.detach(), albedo_map,
                       roughness_map, metallic_map, F0)
    IRR = linear_to_srgb(IRR)
    IRR = median_blur3x3(IRR)
    render_rgb = render_direct + IRR
    rr.update(render_rgb=render_rgb, render_direct=render_direct, indirect=IRR, roughness_remap=roughness_map,
              metallic_used=metallic_map)
    return rr


def masked_tv_loss(mask: torch.Tensor, gt_image: torch.Tensor, prediction: torch.Tensor) -> torch.Tensor:
    """train.py:118-142 get_masked_tv_loss (erosion=False); with an all-true mask it equals get_tv_loss(pad=1, step=1)
    (:83-115), which is the branch train.py:389-401 takes in that case."""
    rgb_grad_h = torch.exp(-(gt_image[:, 1:, :] - gt_image[:, :-1, :]).abs().mean(dim=0, keepdim=True))
    rgb_grad_w = torch.exp(-(gt_image[:, :, 1:] - gt_image[:, :, :-1]).abs().mean(dim=0, keepdim=True))
    tv_h = torch.pow(prediction[:, 1:, :] - prediction[:, :-1, :], 2)
    tv_w = torch.pow(prediction[:, :, 1:] - prediction[:, :, :-1], 2)
    mask = mask.float()
    mask_h = mask[:, 1:, :] * mask[:, :-1, :]
    mask_w = mask[:, :, 1:] * mask[:, :, :-1]
    return (tv_h * rgb_grad_h * mask_h).mean() + (tv_w * rgb_grad_w * mask_w).mean()


def pbr_loss(res: Dict, gt_image: torch.Tensor, lamb_weight: float = 0.001, brdf_tv_weight: float = 0.0) -> torch.Tensor:
    """The PBR-stage loss of one view (train.py:385-404): L1 on render_direct + IRR, the BRDF smoothness prior
    (brdf_tv_weight, reference default 1.0) and the 'lamb' prior. The env-map TV term (train.py:406-420) does not
    depend on the view: gigs.light.env_tv_loss."""
    loss = torch.abs(res["render_rgb"] - gt_image).mean()
    nm = res["normal_mask"].float()
    rough, metal = res["roughness_remap"], res["metallic_used"]
    if brdf_tv_weight:
        loss = loss + brdf_tv_weight * masked_tv_loss(res["normal_mask"], gt_image,
                                                      torch.cat([res["albedo_map"], rough, metal], dim=0))
    cnt = nm.sum().clamp_min(1.0)  # masked means without a host sync (the reference indexes with the mask)
    loss = loss + lamb_weight * (((1.0 - rough) * nm).sum() / cnt + (metal * nm).sum() / cnt)
    return loss
