"""Drop-in for the reference's `diff_gaussian_rasterization` Python package
(/root/reference/submodules/diff-gaussian-rasterization/diff_gaussian_rasterization/__init__.py).

Same public names, argument order, defaults, return order, dtypes and CHW layouts:
    GaussianRasterizationSettings (:31-51)   GaussianRasterizer (:375-537, 12-tuple)
    Gaussian_SSR (:696-743)                  _C.{rasterize_gaussians, lite_rasterize_gaussians,
                                                 rasterize_gaussians_backward, mark_visible,
                                                 depth_to_normal, SSAO, SSR, SSR_BACKWARD}  (ext.cpp:16-24)
backed by hand-written sm_100a kernels behind the C-ABI of include/gigs_b200.h (ctypes, raw device
pointers, the caller's current CUDA stream). There is no CPU / PyTorch fallback: a missing
libgigs_b200.so is an ImportError.

Deliberate differences from the reference module (none changes a returned value):
  * CUDA_LAUNCH_BLOCKING is NOT forced to 1 (:20) and kernels run on torch's current stream rather
    than the legacy default stream;
  * the kornia median / bilateral filters the reference calls (:478,491,504) are our own kernels,
    fused with depth_to_normal into one pass;
  * upstream gradients that autograd did not produce are passed as NULL instead of materialised
    zero tensors (identical result, less work).
"""
from typing import NamedTuple, Optional, Tuple

import ctypes as C
import os
import sys

import torch
import torch.nn as nn

_PKG_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _PKG_ROOT not in sys.path:
    sys.path.insert(0, _PKG_ROOT)

from gigs import _lib  # noqa: E402
from gigs._lib import GigsCamera, GigsLayout, GigsRasterBwd, GigsRasterFwd, GigsSizes, check, ptr  # noqa: E402

_L = _lib.load()


def cpu_deep_copy_tuple(input_tuple: Tuple) -> Tuple:
    copied_tensors = [item.cpu().clone() if isinstance(item, torch.Tensor) else item for item in input_tuple]
    return tuple(copied_tensors)


class GaussianRasterizationSettings(NamedTuple):
    image_height: int
    image_width: int
    tanfovx: float
    tanfovy: float
    radius: float
    bias: float
    thick: float
    delta: float
    step: int
    start: int
    bg: torch.Tensor
    scale_modifier: float
    viewmatrix: torch.Tensor
    projmatrix: torch.Tensor
    sh_degree: int
    campos: torch.Tensor
    prefiltered: bool
    debug: bool
    inference: bool
    argmax_depth: bool


# ------------------------------------------------------------------------------------------------
# helpers
# ------------------------------------------------------------------------------------------------
_scratch = {}   # device index -> cached transient sort scratch (grow-only)
_pinned = {}    # device index -> pinned int32[1] for num_rendered


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _f32(t: Optional[torch.Tensor], name: str) -> Optional[torch.Tensor]:
    """contiguous float32 CUDA tensor, or None for None/empty (the reference calls .contiguous() too)."""
    if t is None or t.numel() == 0:
        return None
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor")
    if t.dtype != torch.float32:
        raise RuntimeError(f"{name} must be float32")
    return t.contiguous()


def _camera(bg, viewmatrix, projmatrix, campos, scale_modifier, tanfovx, tanfovy, H, W, degree, M, prefiltered,
            argmax_depth, inference, debug, keep):
    bg = _f32(bg, "bg"); viewmatrix = _f32(viewmatrix, "viewmatrix")
    projmatrix = _f32(projmatrix, "projmatrix"); campos = _f32(campos, "campos")
    keep.extend([bg, viewmatrix, projmatrix, campos])
    return GigsCamera(int(W), int(H), float(tanfovx), float(tanfovy), float(scale_modifier), int(degree), int(M),
                      int(bool(prefiltered)), int(bool(debug)), int(bool(inference)), int(bool(argmax_depth)),
                      ptr(viewmatrix), ptr(projmatrix), ptr(campos), ptr(bg))


def _sizes(P, W, H, R) -> GigsSizes:
    s = GigsSizes()
    check(_L.gigs_raster_sizes(P, W, H, R, C.byref(s)), "gigs_raster_sizes")
    return s


def raster_layout(P, W, H, R) -> GigsLayout:
    lay = GigsLayout()
    check(_L.gigs_raster_layout(P, W, H, R, C.byref(lay)), "gigs_raster_layout")
    return lay


def sort_scratch(device) -> Optional[torch.Tensor]:
    """The transient sort scratch of the last forward on `device` (tests decode sorted keys from it)."""
    return _scratch.get(torch.device(device).index or 0)


def _rasterize(lite, background, means3D, colors, opacity, normal, albedo, roughness, metallic, scales, rotations,
               cov3D_precomp, sh, campos, viewmatrix, projmatrix, scale_modifier, tan_fovx, tan_fovy, image_height,
               image_width, degree, prefiltered, argmax_depth, inference, debug):
    if means3D.ndimension() != 2 or means3D.size(1) != 3:
        raise RuntimeError("means3D must have dimensions (num_points, 3)")
    if not means3D.is_cuda:
        raise RuntimeError("means3D must be a CUDA tensor")
    dev = means3D.device
    P, H, W = means3D.size(0), int(image_height), int(image_width)
    keep = []
    means3D_c = _f32(means3D, "means3D"); sh_c = _f32(sh, "sh"); colors_c = _f32(colors, "colors_precomp")
    opac_c = _f32(opacity, "opacities"); scales_c = _f32(scales, "scales"); rot_c = _f32(rotations, "rotations")
    cov_c = _f32(cov3D_precomp, "cov3D_precomp")
    normal_c = albedo_c = rough_c = metal_c = None
    if not lite:
        normal_c = _f32(normal, "normal"); albedo_c = _f32(albedo, "albedo")
        rough_c = _f32(roughness, "roughness"); metal_c = _f32(metallic, "metallic")
    M = sh.size(1) if (sh is not None and sh.numel() != 0) else 0
    cam = _camera(background, viewmatrix, projmatrix, campos, scale_modifier, tan_fovx, tan_fovy, H, W, degree, M,
                  prefiltered, argmax_depth, inference, debug, keep)

    f32 = dict(dtype=torch.float32, device=dev)
    out_color = torch.empty((3, H, W), **f32)
    out_opacity = torch.empty((1, H, W), **f32)
    out_depth = torch.empty((1, H, W), **f32)
    radii = torch.empty((P,), dtype=torch.int32, device=dev)
    if lite:
        out_normal = out_normal_view = out_pos = out_albedo = out_roughness = out_metallic = None
    else:
        out_normal = torch.empty((3, H, W), **f32); out_normal_view = torch.empty((3, H, W), **f32)
        out_pos = torch.empty((3, H, W), **f32); out_albedo = torch.empty((3, H, W), **f32)
        out_roughness = torch.empty((1, H, W), **f32); out_metallic = torch.empty((1, H, W), **f32)

    sz0 = _sizes(P, W, H, 0)
    geom = torch.empty((sz0.geom_bytes,), dtype=torch.uint8, device=dev)
    img = torch.empty((sz0.img_bytes,), dtype=torch.uint8, device=dev)
    di = dev.index or 0
    if di not in _pinned:
        _pinned[di] = torch.zeros(1, dtype=torch.int32).pin_memory()
    a = GigsRasterFwd()
    a.P = P; a.material_only = 0; a.cam = cam
    a.means3D = ptr(means3D_c); a.shs = ptr(sh_c); a.colors_precomp = ptr(colors_c); a.opacities = ptr(opac_c)
    a.normal = ptr(normal_c); a.albedo = ptr(albedo_c); a.roughness = ptr(rough_c); a.metallic = ptr(metal_c)
    a.scales = ptr(scales_c); a.rotations = ptr(rot_c); a.cov3D_precomp = ptr(cov_c)
    a.out_color = ptr(out_color); a.out_opacity = ptr(out_opacity); a.out_depth = ptr(out_depth)
    a.out_normal = ptr(out_normal); a.out_normal_view = ptr(out_normal_view); a.out_pos = ptr(out_pos)
    a.out_albedo = ptr(out_albedo); a.out_roughness = ptr(out_roughness); a.out_metallic = ptr(out_metallic)
    a.radii = ptr(radii)
    a.geom = geom.data_ptr(); a.geom_bytes = geom.numel()
    a.img = img.data_ptr(); a.img_bytes = img.numel()
    a.pinned_num_rendered = _pinned[di].data_ptr()
    a.stream = _stream()
    with torch.cuda.device(dev):
        check(_L.gigs_raster_forward_begin(C.byref(a)), "gigs_raster_forward_begin")
        R = int(a.num_rendered)
        sz = _sizes(P, W, H, R)
        binning = torch.empty((sz.binning_bytes,), dtype=torch.uint8, device=dev)
        scratch = _scratch.get(di)
        if scratch is None or scratch.numel() < sz.sort_bytes or scratch.device != dev:
            scratch = torch.empty((int(sz.sort_bytes * 1.25) + 1024,), dtype=torch.uint8, device=dev)
            _scratch[di] = scratch
        a.binning = binning.data_ptr(); a.binning_bytes = binning.numel()
        a.sort = scratch.data_ptr(); a.sort_bytes = scratch.numel()
        fn = _L.gigs_lite_forward_finish if lite else _L.gigs_raster_forward_finish
        check(fn(C.byref(a)), "gigs_raster_forward_finish")
    if lite:
        return R, out_color, out_opacity, radii, out_depth
    return (R, out_color, radii, geom, binning, img, out_opacity, out_depth, out_normal, out_normal_view, out_pos,
            out_albedo, out_roughness, out_metallic)


class _CShim:
    """The eight native entry points the reference binds with pybind (ext.cpp:16-24), same signatures."""

    @staticmethod
    def rasterize_gaussians(background, means3D, colors, opacity, normal, albedo, roughness, metallic, scales,
                            rotations, cov3D_precomp, sh, campos, viewmatrix, projmatrix, scale_modifier, tan_fovx,
                            tan_fovy, image_height, image_width, degree, prefiltered, argmax_depth, inference, debug):
        return _rasterize(False, background, means3D, colors, opacity, normal, albedo, roughness, metallic, scales,
                          rotations, cov3D_precomp, sh, campos, viewmatrix, projmatrix, scale_modifier, tan_fovx,
                          tan_fovy, image_height, image_width, degree, prefiltered, argmax_depth, inference, debug)

    @staticmethod
    def lite_rasterize_gaussians(background, means3D, colors, opacity, scales, rotations, cov3D_precomp, sh, campos,
                                 viewmatrix, projmatrix, scale_modifier, tan_fovx, tan_fovy, image_height, image_width,
                                 degree, prefiltered, argmax_depth):
        return _rasterize(True, background, means3D, colors, opacity, None, None, None, None, scales, rotations,
                          cov3D_precomp, sh, campos, viewmatrix, projmatrix, scale_modifier, tan_fovx, tan_fovy,
                          image_height, image_width, degree, prefiltered, argmax_depth, False, False)

    @staticmethod
    def rasterize_gaussians_backward(background, means3D, radii, colors, normal, albedo, roughness, metallic, scales,
                                     rotations, cov3D_precomp, sh, campos, viewmatrix, projmatrix, scale_modifier,
                                     tan_fovx, tan_fovy, degree, dL_dout_depth, dL_dout_color, dL_dout_opacity,
                                     dL_dout_normal, dL_dout_albedo, dL_dout_roughness, dL_dout_metallic, geomBuffer,
                                     binningBuffer, imageBuffer, R, debug, image_height=None, image_width=None):
        dev = means3D.device
        P = means3D.size(0)
        M = sh.size(1) if (sh is not None and sh.numel() != 0) else 0
        if image_height is None:
            ref_map = next(g for g in (dL_dout_color, dL_dout_albedo, dL_dout_normal, dL_dout_opacity, dL_dout_depth,
                                       dL_dout_roughness, dL_dout_metallic) if g is not None)
            image_height, image_width = ref_map.size(1), ref_map.size(2)
        H, W = int(image_height), int(image_width)
        keep = []
        cam = _camera(background, viewmatrix, projmatrix, campos, scale_modifier, tan_fovx, tan_fovy, H, W, degree, M,
                      False, False, False, debug, keep)
        f32 = dict(dtype=torch.float32, device=dev)
        # every element is written by the kernel: no zero fill (the reference does 14 torch::zeros)
        dL_dmeans3D = torch.empty((P, 3), **f32); dL_dmeans2D = torch.empty((P, 3), **f32)
        dL_dcolors = torch.empty((P, 3), **f32); dL_dopacity = torch.empty((P, 1), **f32)
        dL_dnormal = torch.empty((P, 3), **f32); dL_dalbedo = torch.empty((P, 3), **f32)
        dL_droughness = torch.empty((P, 1), **f32); dL_dmetallic = torch.empty((P, 1), **f32)
        dL_dcov3D = torch.empty((P, 6), **f32)
        has_scales = scales is not None and scales.numel() != 0
        dL_dsh = torch.empty((P, M, 3), **f32) if M > 0 else torch.zeros((P, M, 3), **f32)
        dL_dscales = torch.empty((P, 3), **f32) if has_scales else torch.zeros((P, 3), **f32)
        dL_drotations = torch.empty((P, 4), **f32) if has_scales else torch.zeros((P, 4), **f32)
        if P == 0:
            return (dL_dmeans2D, dL_dcolors, dL_dopacity, dL_dnormal, dL_dalbedo, dL_droughness, dL_dmetallic,
                    dL_dmeans3D, dL_dcov3D, dL_dsh, dL_dscales, dL_drotations)
        accum = torch.empty((P, 20), **f32)
        tens = [_f32(t, n) for t, n in ((means3D, "means3D"), (sh, "sh"), (colors, "colors_precomp"),
                                        (normal, "normal"), (albedo, "albedo"), (roughness, "roughness"),
                                        (metallic, "metallic"), (scales, "scales"), (rotations, "rotations"),
                                        (cov3D_precomp, "cov3D_precomp"))]
        grads = [_f32(g, "upstream gradient") for g in (dL_dout_depth, dL_dout_color, dL_dout_opacity, dL_dout_normal,
                                                        dL_dout_albedo, dL_dout_roughness, dL_dout_metallic)]
        a = GigsRasterBwd()
        a.P = P; a.num_rendered = int(R); a.cam = cam
        (a.means3D, a.shs, a.colors_precomp, a.normal, a.albedo, a.roughness, a.metallic, a.scales, a.rotations,
         a.cov3D_precomp) = [ptr(t) for t in tens]
        a.radii = ptr(radii.contiguous())
        a.geom = geomBuffer.data_ptr(); a.binning = binningBuffer.data_ptr(); a.img = imageBuffer.data_ptr()
        (a.dL_dpix_depth, a.dL_dpix, a.dL_dpix_opacity, a.dL_dpix_normal, a.dL_dpix_albedo, a.dL_dpix_roughness,
         a.dL_dpix_metallic) = [ptr(g) for g in grads]
        a.accum = accum.data_ptr()
        a.dL_dmean2D = dL_dmeans2D.data_ptr(); a.dL_dconic = None; a.dL_dopacity = dL_dopacity.data_ptr()
        a.dL_dcolor = dL_dcolors.data_ptr(); a.dL_dnormal = dL_dnormal.data_ptr(); a.dL_dalbedo = dL_dalbedo.data_ptr()
        a.dL_droughness = dL_droughness.data_ptr(); a.dL_dmetallic = dL_dmetallic.data_ptr()
        a.dL_dmean3D = dL_dmeans3D.data_ptr(); a.dL_dcov3D = dL_dcov3D.data_ptr()
        a.dL_dsh = dL_dsh.data_ptr() if M > 0 else None
        a.dL_dscale = dL_dscales.data_ptr() if has_scales else None
        a.dL_drot = dL_drotations.data_ptr() if has_scales else None
        a.stream = _stream()
        with torch.cuda.device(dev):
            check(_L.gigs_raster_backward(C.byref(a)), "gigs_raster_backward")
        return (dL_dmeans2D, dL_dcolors, dL_dopacity, dL_dnormal, dL_dalbedo, dL_droughness, dL_dmetallic,
                dL_dmeans3D, dL_dcov3D, dL_dsh, dL_dscales, dL_drotations)

    @staticmethod
    def mark_visible(means3D, viewmatrix, projmatrix):
        P = means3D.size(0)
        present = torch.zeros((P,), dtype=torch.bool, device=means3D.device)
        if P != 0:
            m, v = _f32(means3D, "means3D"), _f32(viewmatrix, "viewmatrix")
            with torch.cuda.device(means3D.device):
                check(_L.gigs_mark_visible(P, ptr(m), ptr(v), present.data_ptr(), _stream()), "gigs_mark_visible")
        return present

    @staticmethod
    def depth_to_normal(width, height, focal_x, focal_y, viewmatrix, depthMap):
        d, v = _f32(depthMap, "depthMap"), _f32(viewmatrix, "viewmatrix")
        normalMap = torch.empty((3, height, width), dtype=torch.float32, device=d.device)
        depth_pos = torch.empty((3, height, width), dtype=torch.float32, device=d.device)
        with torch.cuda.device(d.device):
            check(_L.gigs_depth_to_normal(int(width), int(height), float(focal_x), float(focal_y), ptr(v), ptr(d),
                                          normalMap.data_ptr(), depth_pos.data_ptr(), _stream()), "gigs_depth_to_normal")
        return normalMap, depth_pos

    @staticmethod
    def SSAO(width, height, focal_x, focal_y, radius, bias, thick, delta, step, start, out_normal, out_pos):
        n, p = _f32(out_normal, "out_normal"), _f32(out_pos, "out_pos")
        occlusion = torch.empty((1, height, width), dtype=torch.float32, device=n.device)
        scratch = torch.empty(int(_L.gigs_gi_scratch_bytes(int(width), int(height))), dtype=torch.uint8, device=n.device)
        with torch.cuda.device(n.device):
            check(_L.gigs_ssao(int(width), int(height), float(focal_x), float(focal_y), float(radius), float(bias),
                               float(thick), float(delta), int(step), int(start), ptr(n), ptr(p),
                               occlusion.data_ptr(), scratch.data_ptr(), scratch.numel(), _stream()), "gigs_ssao")
        return occlusion

    @staticmethod
    def SSR(width, height, focal_x, focal_y, radius, bias, thick, delta, step, start, out_normal, out_pos, out_rgb,
            out_albedo, out_roughness, out_metallic, out_F0):
        ts = [_f32(t, n) for t, n in ((out_normal, "normal"), (out_pos, "pos"), (out_rgb, "rgb"),
                                      (out_albedo, "albedo"), (out_roughness, "roughness"),
                                      (out_metallic, "metallic"), (out_F0, "F0"))]
        dev = ts[0].device
        color = torch.empty((3, height, width), dtype=torch.float32, device=dev)
        abd = torch.empty((3, height, width), dtype=torch.float32, device=dev)
        scratch = torch.empty(int(_L.gigs_gi_scratch_bytes(int(width), int(height))), dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            check(_L.gigs_ssr(int(width), int(height), float(focal_x), float(focal_y), float(radius), float(bias),
                              float(thick), float(delta), int(step), int(start), *[ptr(t) for t in ts],
                              color.data_ptr(), abd.data_ptr(), scratch.data_ptr(), scratch.numel(), _stream()),
                  "gigs_ssr")
        return color, abd

    @staticmethod
    def SSR_BACKWARD(width, height, focal_x, focal_y, out_normal, out_pos, out_rgb, out_albedo, out_roughness,
                     out_metallic, out_F0, dL_dpixels, abd=None):
        """The reference never calls its native SSR_BACKWARD kernel (its Python computes
        grad_albedo = grad * abd, :666-673). This entry point implements those live semantics and
        therefore needs `abd` (the second output of SSR); shapes follow rasterize_points.cu:492-494."""
        if abd is None:
            raise RuntimeError("SSR_BACKWARD needs abd (the reference's native kernel is dead code; see docstring)")
        g, ab = _f32(dL_dpixels, "dL_dpixels"), _f32(abd, "abd")
        dl_albedo = torch.empty((3, height, width), dtype=torch.float32, device=g.device)
        dl_roughness = torch.zeros((3, height, width), dtype=torch.float32, device=g.device)
        dl_metallic = torch.zeros((3, height, width), dtype=torch.float32, device=g.device)
        with torch.cuda.device(g.device):
            check(_L.gigs_ssr_backward(int(width), int(height), ptr(g), ptr(ab), dl_albedo.data_ptr(), None, None,
                                       _stream()), "gigs_ssr_backward")
        return dl_albedo, dl_roughness, dl_metallic


_C = _CShim()


def geometry_chain(width, height, focal_x, focal_y, viewmatrix, depth, derive_normal=True):
    """median3x3(depth) -> depth_to_normal -> (bilateral3x3(normal), median3x3(depth_pos)) in one kernel
    (reference :475-504)."""
    d, v = _f32(depth, "depth"), _f32(viewmatrix, "viewmatrix")
    normal = torch.empty((3, height, width), dtype=torch.float32, device=d.device)
    pos = torch.empty((3, height, width), dtype=torch.float32, device=d.device)
    with torch.cuda.device(d.device):
        check(_L.gigs_geometry_chain(int(width), int(height), float(focal_x), float(focal_y), ptr(v), ptr(d),
                                     int(bool(derive_normal)), normal.data_ptr(), pos.data_ptr(), _stream()),
              "gigs_geometry_chain")
    return normal, pos


class _Median3x3(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        xc = _f32(x, "median input")
        Cn, H, W = xc.shape
        out = torch.empty_like(xc)
        with torch.cuda.device(xc.device):
            check(_L.gigs_median3x3(Cn, W, H, ptr(xc), out.data_ptr(), _stream()), "gigs_median3x3")
        ctx.save_for_backward(xc)
        return out

    @staticmethod
    def backward(ctx, g):
        (xc,) = ctx.saved_tensors
        Cn, H, W = xc.shape
        gi = torch.empty_like(xc)
        gc = g.contiguous()
        with torch.cuda.device(xc.device):
            check(_L.gigs_median3x3_backward(Cn, W, H, ptr(xc), ptr(gc), gi.data_ptr(), _stream()),
                  "gigs_median3x3_backward")
        return gi


def median_blur3x3(x: torch.Tensor) -> torch.Tensor:
    """[C,H,W] -> [C,H,W]; kornia.filters.median_blur(x[None], (3,3))[0] semantics, differentiable."""
    return _Median3x3.apply(x)


def bilateral_blur3x3(x: torch.Tensor, sigma_color: float = 1.0, sigma_space: float = 3.0) -> torch.Tensor:
    """kornia.filters.bilateral_blur(x[None], (3,3), sigma_color, (sigma_space,)*2)[0] semantics (no autograd)."""
    xc = _f32(x, "bilateral input")
    Cn, H, W = xc.shape
    out = torch.empty_like(xc)
    with torch.cuda.device(xc.device):
        check(_L.gigs_bilateral3x3(Cn, W, H, float(sigma_color), float(sigma_space), ptr(xc), out.data_ptr(),
                                   _stream()), "gigs_bilateral3x3")
    return out


# ------------------------------------------------------------------------------------------------
# autograd wrappers (same structure as the reference's)
# ------------------------------------------------------------------------------------------------
class _RasterizeGaussians(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means3D, means2D, opacities, normal, albedo, roughness, metallic, sh, colors_precomp, scales,
                rotations, cov3Ds_precomp, raster_settings: GaussianRasterizationSettings):
        # Restructure arguments the way that the native lib expects them (reference :84-110)
        args = (
            raster_settings.bg, means3D, colors_precomp, opacities, normal, albedo, roughness, metallic, scales,
            rotations, cov3Ds_precomp, sh, raster_settings.campos, raster_settings.viewmatrix,
            raster_settings.projmatrix, raster_settings.scale_modifier, raster_settings.tanfovx,
            raster_settings.tanfovy, raster_settings.image_height, raster_settings.image_width,
            raster_settings.sh_degree, raster_settings.prefiltered, raster_settings.argmax_depth,
            raster_settings.inference, raster_settings.debug,
        )
        if raster_settings.debug:
            cpu_args = cpu_deep_copy_tuple(args)  # Copy them before they can be corrupted
            try:
                res = _C.rasterize_gaussians(*args)
            except Exception as ex:
                torch.save(cpu_args, "snapshot_fw.dump")
                print("\nAn error occured in forward. Please forward snapshot_fw.dump for debugging.")
                raise ex
        else:
            res = _C.rasterize_gaussians(*args)
        (num_rendered, color, radii, geomBuffer, binningBuffer, imgBuffer, opacity_map, depth, out_normal,
         out_normal_view, out_pos, albedo_map, roughness_map, metallic_map) = res

        ctx.raster_settings = raster_settings
        ctx.num_rendered = num_rendered
        ctx.set_materialize_grads(False)
        ctx.mark_non_differentiable(radii)
        ctx.save_for_backward(colors_precomp, normal, albedo, roughness, metallic, means3D, scales, rotations,
                              cov3Ds_precomp, radii, sh, geomBuffer, binningBuffer, imgBuffer)
        return (color, radii, opacity_map, depth, out_normal, albedo_map, roughness_map, metallic_map,
                out_normal_view, out_pos)

    @staticmethod
    def backward(ctx, grad_out_color, gard_radii=None, grad_out_opacity=None, grad_depth=None, grad_out_normal=None,
                 grad_out_albedo=None, grad_out_roughness=None, grad_out_metallic=None, grad_out_normal_view=None,
                 grad_out_pos=None):
        num_rendered = ctx.num_rendered
        rs = ctx.raster_settings
        (colors_precomp, normal, albedo, roughness, metallic, means3D, scales, rotations, cov3Ds_precomp, radii, sh,
         geomBuffer, binningBuffer, imgBuffer) = ctx.saved_tensors
        # grad_out_normal_view / grad_out_pos are ignored, as in the reference (:320-334)
        args = (
            rs.bg, means3D, radii, colors_precomp, normal, albedo, roughness, metallic, scales, rotations,
            cov3Ds_precomp, sh, rs.campos, rs.viewmatrix, rs.projmatrix, rs.scale_modifier, rs.tanfovx, rs.tanfovy,
            rs.sh_degree, grad_depth, grad_out_color, grad_out_opacity, grad_out_normal, grad_out_albedo,
            grad_out_roughness, grad_out_metallic, geomBuffer, binningBuffer, imgBuffer, num_rendered, rs.debug,
        )
        kw = dict(image_height=rs.image_height, image_width=rs.image_width)
        if rs.debug:
            cpu_args = cpu_deep_copy_tuple(args)
            try:
                res = _C.rasterize_gaussians_backward(*args, **kw)
            except Exception as ex:
                torch.save(cpu_args, "snapshot_bw.dump")
                print("\nAn error occured in backward. Writing snapshot_bw.dump for debugging.\n")
                raise ex
        else:
            res = _C.rasterize_gaussians_backward(*args, **kw)
        (grad_means2D, grad_colors_precomp, grad_opacities, grad_normal, grad_albedo, grad_roughness, grad_metallic,
         grad_means3D, grad_cov3Ds_precomp, grad_sh, grad_scales, grad_rotations) = res

        def _or_none(g, inp):
            return g if (inp is not None and inp.numel() != 0) else None

        return (grad_means3D, grad_means2D, grad_opacities, grad_normal, grad_albedo, grad_roughness, grad_metallic, _or_none(grad_sh, sh),
                _or_none(grad_colors_precomp, colors_precomp), _or_none(grad_scales, scales),
                _or_none(grad_rotations, rotations), _or_none(grad_cov3Ds_precomp, cov3Ds_precomp), None)



class GaussianRasterizer(nn.Module):
    def __init__(self, raster_settings: GaussianRasterizationSettings):
        super().__init__()
        self.raster_settings = raster_settings

    def markVisible(self, positions: torch.Tensor) -> torch.Tensor:
        # Mark visible points (based on frustum culling for camera) with a boolean
        with torch.no_grad():
            raster_settings = self.raster_settings
            visible = _C.mark_visible(positions, raster_settings.viewmatrix, raster_settings.projmatrix)
        return visible

    def forward(self, means3D, means2D, opacities, normal, albedo, roughness, metallic, shs=None, colors_precomp=None,
                scales=None, rotations=None, cov3D_precomp=None, derive_normal: bool = True):
        raster_settings = self.raster_settings

        if (shs is None and colors_precomp is None) or (shs is not None and colors_precomp is not None):
            raise Exception("Please provide excatly one of either SHs or precomputed colors!")

        if ((scales is None or rotations is None) and cov3D_precomp is None) or (
            (scales is not None or rotations is not None) and cov3D_precomp is not None
        ):
            raise Exception("Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!")

        if shs is None:
            shs = torch.Tensor([])
        if colors_precomp is None:
            colors_precomp = torch.Tensor([])
        if scales is None:
            scales = torch.Tensor([])
        if rotations is None:
            rotations = torch.Tensor([])
        if cov3D_precomp is None:
            cov3D_precomp = torch.Tensor([])

        (color, radii, opacity_map, depth, out_normal, albedo_map, roughness_map, metallic_map, out_normal_view,
         _) = _RasterizeGaussians.apply(means3D, means2D, opacities, normal, albedo, roughness, metallic, shs,
                                        colors_precomp, scales, rotations, cov3D_precomp, raster_settings)

        W, H = raster_settings.image_width, raster_settings.image_height
        focal_x = W / (2.0 * raster_settings.tanfovx)
        focal_y = H / (2.0 * raster_settings.tanfovy)
        with torch.no_grad():
            # reference :475-504 (median -> depth_to_normal -> bilateral; median(depth_pos)), fused
            normal_from_depth, depth_pos_filter = geometry_chain(W, H, focal_x, focal_y, raster_settings.viewmatrix,
                                                                 depth, derive_normal)
            occlusion = _C.SSAO(W, H, focal_x, focal_y, raster_settings.radius, raster_settings.bias,
                                raster_settings.thick, raster_settings.delta, raster_settings.step,
                                raster_settings.start, out_normal_view, depth_pos_filter)

        return (color, radii, opacity_map, depth, normal_from_depth, out_normal, occlusion, albedo_map, roughness_map,
                metallic_map, out_normal_view, depth_pos_filter)


class _SSR(torch.autograd.Function):
    @staticmethod
    def forward(ctx, image_width, image_height, focal_x, focal_y, radius, bias, thick, delta, step, start, normal,
                pos, rgb, albedo, roughness, metallic, F0):
        (color, abd) = _C.SSR(image_width, image_height, focal_x, focal_y, radius, bias, thick, delta, step, start,
                              normal, pos, rgb, albedo, roughness, metallic, F0)
        ctx.image_width = image_width
        ctx.image_height = image_height
        ctx.save_for_backward(roughness, metallic, abd)
        return (color, abd)

    @staticmethod
    def backward(ctx, grad_out_color, grad_abd=None):
        roughness, metallic, abd = ctx.saved_tensors
        W, H = ctx.image_width, ctx.image_height
        # live semantics of the reference (:671-673): grad_albedo = g * abd, zeros for roughness / metallic
        g = _f32(grad_out_color, "grad_out_color")
        grad_albedo = torch.empty_like(abd)
        grad_roughness = torch.empty_like(roughness)
        grad_metallic = torch.empty_like(metallic)
        with torch.cuda.device(abd.device):
            check(_L.gigs_ssr_backward(int(W), int(H), ptr(g), ptr(abd), grad_albedo.data_ptr(),
                                       grad_roughness.data_ptr(), grad_metallic.data_ptr(), _stream()),
                  "gigs_ssr_backward")
        return (None,) * 13 + (grad_albedo, grad_roughness, grad_metallic, None)


class Gaussian_SSR(nn.Module):
    def __init__(self, tanfovx, tanfovy, image_width, image_height, radius, bias, thick, delta, step, start):
        super().__init__()
        self.tanfovx = tanfovx
        self.tanfovy = tanfovy
        self.image_width = image_width
        self.image_height = image_height
        self.radius = radius
        self.bias = bias
        self.thick = thick
        self.delta = delta
        self.step = step
        self.start = start

    def forward(self, normal, pos, rgb, albedo, roughness, metallic, F0):
        focal_x = self.image_width / (2.0 * self.tanfovx)
        focal_y = self.image_height / (2.0 * self.tanfovy)
        (color, abd) = _SSR.apply(self.image_width, self.image_height, focal_x, focal_y, self.radius, self.bias,
                                  self.thick, self.delta, self.step, self.start, normal, pos, rgb, albedo, roughness,
                                  metallic, F0)
        return (color, abd)
