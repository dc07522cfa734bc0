"""Fused PBR-stage frame: host side of gigs_frame_forward / gigs_frame_backward (include/gigs_b200.h).

One view of /root/reference/train.py:266-404 — getters, rasterize, geometry chain + SSAO, render() post-processing,
pbr_shading, SSR, loss, and the whole backward — as two C-ABI calls on the caller's current stream, with every
workspace and intermediate map in buffers that persist across frames (no per-frame allocation). The unfused
operator path (gigs.renderer / diff_gaussian_rasterization) computes the same thing op by op and is what the parity
tests compare this against.
"""
import ctypes as C
from typing import Dict, Optional

import torch

from . import _lib
from ._lib import GIGS_E_GROW, GigsCamera, GigsFrame, GigsFrameLayout, GigsSizes, check

_L = _lib.load()

_FLOAT_PLANES = dict(color=3, opacity=1, depth=1, normal=3, normal_view=3, pos=3, albedo=3, roughness=1, metallic=1,
                     normal_from_depth=3, depth_pos=3, occlusion=1, shade_normal=3, ssr_normal=3, render_direct=3,
                     linear_rgb=3, F0=3, rough_remap=1, metal_used=1, ssr_color=3, ssr_abd=3, render_rgb=3, g_rgb=3,
                     g_albedo=3, g_roughness=1, g_metallic=1)


class FrameWorkspace:
    """Device buffers of one (P, W, H) frame shape. geom / img / maps / radii / accum are sized by shape; binning and
    the sort scratch grow (x1.25) when a frame needs more than any before it."""

    def __init__(self, P: int, W: int, H: int, device):
        self.P, self.W, self.H, self.device = P, W, H, torch.device(device)
        sz = GigsSizes()
        check(_L.gigs_raster_sizes(P, W, H, 0, C.byref(sz)), "gigs_raster_sizes")
        self.layout = GigsFrameLayout()
        check(_L.gigs_frame_layout(W, H, C.byref(self.layout)), "gigs_frame_layout")
        u8 = dict(dtype=torch.uint8, device=self.device)
        self.geom = torch.empty(sz.geom_bytes, **u8)
        self.img = torch.empty(sz.img_bytes, **u8)
        self.maps = torch.empty(self.layout.total_bytes, **u8)
        self.radii = torch.empty(P, dtype=torch.int32, device=self.device)
        self.accum = torch.empty((P, 20), dtype=torch.float32, device=self.device)
        self.binning: Optional[torch.Tensor] = None
        self.sort: Optional[torch.Tensor] = None
        self.pinned = torch.zeros(1, dtype=torch.int32).pin_memory()
        self.num_rendered = 0

    def grow(self, binning_bytes: int, sort_bytes: int):
        u8 = dict(dtype=torch.uint8, device=self.device)
        if self.binning is None or self.binning.numel() < binning_bytes:
            self.binning = torch.empty(int(binning_bytes * 1.25) + 1024, **u8)
        if self.sort is None or self.sort.numel() < sort_bytes:
            self.sort = torch.empty(int(sort_bytes * 1.25) + 1024, **u8)

    def map(self, name: str) -> torch.Tensor:
        """View of one intermediate map of the last frame ([c,H,W] float32; 'mask' [H,W] / 'median_sel' [3,H,W] uint8;
        'stats' float32[8] = loss, l1_mean, mask_count, sum((1-rough)*mask), sum(metal*mask), brdf_tv)."""
        off = getattr(self.layout, name)
        N = self.W * self.H
        if name in _FLOAT_PLANES:
            c = _FLOAT_PLANES[name]
            return self.maps[off:off + 4 * c * N].view(torch.float32).view(c, self.H, self.W)
        if name == "mask":
            return self.maps[off:off + N].view(self.H, self.W)
        if name == "median_sel":
            return self.maps[off:off + 3 * N].view(3, self.H, self.W)
        if name == "stats":
            return self.maps[off:off + 32].view(torch.float32)
        raise KeyError(name)


_workspaces: Dict = {}


def workspace(P: int, W: int, H: int, device) -> FrameWorkspace:
    key = (P, W, H, torch.device(device).index or 0)
    ws = _workspaces.get(key)
    if ws is None:
        # densification changes P every 100 iterations: the buffers of the previous model size are dropped, not kept
        for old in [k for k in _workspaces if k[1:] == key[1:] and k[0] != P]:
            del _workspaces[old]
        ws = _workspaces[key] = FrameWorkspace(P, W, H, device)
    return ws


def _p(t):
    return None if t is None else t.data_ptr()


def make_frame(ws: FrameWorkspace, cam, bg, params: Dict, raw: bool, sh_degree: int, light, brdf_lut, rays, gt,
               gi: Dict, indirect: bool, metallic: bool, tone: bool, gamma: bool, loss_scale: float, lamb_weight: float,
               keep: list, gt_ready: Optional[torch.cuda.Event] = None, inference: bool = False,
               brdf_tv_weight: float = 0.0, material_only: bool = False, skip_geometry: bool = False) -> GigsFrame:
    """Fill the C-ABI argument struct of one frame. `keep` receives every tensor whose pointer went into the struct
    (hold it until the calls are done). gt_ready: an event recorded on the stream that copies `gt` to the device; the
    frame's stream waits on it only right before the loss kernel, so the copy overlaps the rasterizer.
    params: raw leaves (xyz, f_dc, f_rest, opacity, normal, albedo, roughness, metallic, log_scale, rot) when `raw`,
    else activated tensors (means3D, shs, opacity, normal, albedo, roughness, metallic, scales, rotations)."""
    f = GigsFrame()
    f.P = ws.P
    f.raw_params = int(raw)

    def c32(t):
        t = t.detach()
        if t.dtype != torch.float32 or not t.is_contiguous():
            t = t.float().contiguous()
        keep.append(t)
        return t

    vm, pm, cp, bgc = c32(cam.world_view_transform), c32(cam.full_proj_transform), c32(cam.camera_center), c32(bg)
    if raw:
        M = 1 + params["f_rest"].shape[1]
        f.means3D = _p(c32(params["xyz"])); f.sh_dc = _p(c32(params["f_dc"])); f.sh_rest = _p(c32(params["f_rest"]))
        f.opacities = _p(c32(params["opacity"])); f.scales = _p(c32(params["log_scale"]))
        f.rotations = _p(c32(params["rot"]))
    else:
        M = params["shs"].shape[1]
        f.means3D = _p(c32(params["means3D"])); f.sh_dc = _p(c32(params["shs"])); f.sh_rest = None
        f.opacities = _p(c32(params["opacity"])); f.scales = _p(c32(params["scales"]))
        f.rotations = _p(c32(params["rotations"]))
    f.normal = _p(c32(params["normal"])); f.albedo = _p(c32(params["albedo"]))
    f.roughness = _p(c32(params["roughness"])); f.metallic = _p(c32(params["metallic"]))
    f.cam = GigsCamera(int(cam.image_width), int(cam.image_height), float(cam.tanfovx), float(cam.tanfovy), 1.0,
                       int(sh_degree), int(M), 0, 0, int(bool(inference)), 0, _p(vm), _p(pm), _p(cp), _p(bgc))
    f.radius = float(gi.get("radius", 0.8)); f.bias = float(gi.get("bias", 0.01)); f.thick = float(gi.get("thick", 0.05))
    f.delta = float(gi.get("delta", 0.0625)); f.step = int(gi.get("step", 16)); f.start = int(gi.get("start", 8))
    f.indirect = int(bool(indirect)); f.use_metallic = int(bool(metallic)); f.tone = int(bool(tone))
    f.gamma = int(bool(gamma))
    spec = [c32(s) for s in light.specular]
    f.n_spec_levels = len(spec)
    for i, s in enumerate(spec):
        f.spec_res[i] = s.shape[1]
        f.spec[i] = s.data_ptr()
    dtex = c32(light.diffuse)
    f.diffuse_res = dtex.shape[1]; f.diffuse = dtex.data_ptr()
    lut = c32(brdf_lut)
    f.brdf_lut = lut.data_ptr(); f.lut_res = lut.shape[-2]
    f.min_roughness = float(getattr(light, "MIN_ROUGHNESS", 0.08))
    f.max_roughness = float(getattr(light, "MAX_ROUGHNESS", 0.5))
    f.canonical_rays = _p(c32(rays))
    f.gt_image = _p(c32(gt)) if gt is not None else None
    f.loss_scale = float(loss_scale); f.lamb_weight = float(lamb_weight); f.brdf_tv_weight = float(brdf_tv_weight)
    f.material_only = int(bool(material_only))
    f.skip_geometry = int(bool(skip_geometry))
    f.geom = ws.geom.data_ptr(); f.geom_bytes = ws.geom.numel()
    f.img = ws.img.data_ptr(); f.img_bytes = ws.img.numel()
    f.maps = ws.maps.data_ptr(); f.maps_bytes = ws.maps.numel()
    f.radii = ws.radii.data_ptr(); f.accum = ws.accum.data_ptr()
    f.pinned_num_rendered = ws.pinned.data_ptr()
    if gt_ready is not None:
        keep.append(gt_ready)
        f.gt_ready_event = gt_ready.cuda_event
    f.stream = torch.cuda.current_stream().cuda_stream
    return f


_fill = make_frame


def _set_sort(ws: FrameWorkspace, f: GigsFrame):
    f.binning = _p(ws.binning); f.binning_bytes = 0 if ws.binning is None else ws.binning.numel()
    f.sort = _p(ws.sort); f.sort_bytes = 0 if ws.sort is None else ws.sort.numel()


def frame_forward(ws: FrameWorkspace, f: GigsFrame) -> torch.Tensor:
    """Runs the forward; returns the device scalar holding the loss (a view into the maps blob)."""
    with torch.cuda.device(ws.device):
        _set_sort(ws, f)
        f.resume = 0
        st = _L.gigs_frame_forward(C.byref(f))
        if st == GIGS_E_GROW:
            ws.grow(f.need_binning_bytes, f.need_sort_bytes)
            _set_sort(ws, f)
            f.resume = 1
            st = _L.gigs_frame_forward(C.byref(f))
        check(st, "gigs_frame_forward")
    ws.num_rendered = int(f.num_rendered)
    return ws.map("stats")[0]


def frame_backward(ws: FrameWorkspace, f: GigsFrame, g_albedo, g_roughness, g_metallic, g_diffuse_tex, g_spec,
                   light_ready: Optional[torch.cuda.Event] = None):
    """Accumulates (+=) into the given gradient tensors (contiguous float32; None = not wanted). light_ready is
    recorded on the stream as soon as the light-texture gradients are final (before the blend backward)."""
    if light_ready is not None:
        # torch creates the cudaEvent lazily on first record: materialise the handle (the C side re-records it at the
        # right point; a wait issued after this call returns sees that later record)
        light_ready.record()
        f.light_ready_event = light_ready.cuda_event
    else:
        f.light_ready_event = None
    f.g_albedo = _p(g_albedo); f.g_roughness = _p(g_roughness); f.g_metallic = _p(g_metallic)
    f.g_diffuse_tex = _p(g_diffuse_tex)
    for i in range(8):
        f.g_spec[i] = _p(g_spec[i]) if (g_spec is not None and i < len(g_spec)) else None
    with torch.cuda.device(ws.device):
        check(_L.gigs_frame_backward(C.byref(f)), "gigs_frame_backward")


def pbr_frame_eval(g: Dict, cam, light, brdf_lut, rays, background, gi: Dict, metallic=True, gamma=True, tone=False,
                   indirect=True, inference=True, raw: bool = False, sh_degree: Optional[int] = None) -> FrameWorkspace:
    """Forward-only frame for the evaluation / relighting sweeps (render.py:202-333, relight.py:114-251): G-buffer,
    SSAO, shading, SSR, final image, no loss. `g` holds activated tensors (scene.activate / GaussianParams.activated)
    or raw leaves (raw=True). Returns the workspace; ws.map("render_rgb") etc. are views valid until the next frame."""
    P = (g["xyz"] if raw else g["means3D"]).shape[0]
    dev = (g["xyz"] if raw else g["means3D"]).device
    ws = workspace(P, int(cam.image_width), int(cam.image_height), dev)
    keep: list = []
    deg = g["sh_degree"] if sh_degree is None else sh_degree
    f = make_frame(ws, cam, background, g, raw, deg, light, brdf_lut, rays, None, gi, indirect, metallic, tone, gamma,
                   1.0, 0.0, keep, inference=inference)
    frame_forward(ws, f)
    return ws


def _is_f32c(*ts) -> bool:
    return all(t is not None and t.dtype == torch.float32 and t.is_contiguous() and t.is_cuda for t in ts)


def _grad_of(t: torch.Tensor) -> Optional[torch.Tensor]:
    """The tensor autograd would accumulate into (allocated zero-filled on first use), None for non-leaves."""
    if not t.requires_grad:
        return None
    if t.grad is None:
        t.grad = torch.zeros_like(t, memory_format=torch.contiguous_format)
    if t.grad.dtype != torch.float32 or not t.grad.is_contiguous():
        raise RuntimeError("fused frame: parameter gradients must be contiguous float32")
    return t.grad


def pbr_frame_step(params, cam, light, brdf_lut, rays, gt_image, background, gi: Dict, metallic=True, gamma=True,
                   tone=False, indirect=True, loss_scale: float = 1.0, lamb_weight: float = 0.001,
                   backward: bool = True, gt_ready: Optional[torch.cuda.Event] = None,
                   light_ready: Optional[torch.cuda.Event] = None, brdf_tv_weight: float = 0.0,
                   radiance: bool = False) -> torch.Tensor:
    """forward (+ backward) of one PBR-stage view for a gigs.step.GaussianParams: gradients accumulate into
    params.flat_grad exactly as autograd would through the unfused path. Returns the (detached) loss scalar.
    radiance=False: only what the PBR-stage loss reads is produced. The SH radiance image (`color`) and the blended
    position (`pos`) are not (train.py:302-420 uses `render` only in its first stage and for logging): preprocess skips
    the SH evaluation and the blend 6 of its 17 channels. Without a GI march (start >= step) the depth -> normal /
    position chain is left out as well (`normal_from_depth` is a first-stage loss term, `depth_pos` only feeds the
    march, train.py:290-381): those two maps are then not written. Loss, gradients and every other map are unchanged."""
    L = params.leaves
    dev = L["xyz"].device
    ws = workspace(params.P, int(cam.image_width), int(cam.image_height), dev)
    # The argument struct is ~60 fields; filling it costs ~85 us of Python per frame, which is GPU idle time in an
    # end-to-end loop that reads the loss back every step. Everything that does not change from frame to frame
    # (parameter / light / LUT / ray pointers, flags) is cached per workspace; a hit only refreshes the per-view fields.
    spec = list(light.specular)
    key = (tuple(t.data_ptr() for t in L.values()), tuple(t.data_ptr() for t in spec), light.diffuse.data_ptr(),
           brdf_lut.data_ptr(), rays.data_ptr(), params.sh_degree, bool(indirect), bool(metallic), bool(tone),
           bool(gamma), tuple(sorted(gi.items())), float(lamb_weight), float(brdf_tv_weight), bool(radiance))
    cached = getattr(ws, "_frame_cache", None)
    if cached is not None and cached[0] == key and _is_f32c(cam.world_view_transform, cam.full_proj_transform,
                                                             cam.camera_center, background, gt_image):
        f, keep = cached[1], cached[2]
        f.cam.tan_fovx = float(cam.tanfovx); f.cam.tan_fovy = float(cam.tanfovy)
        f.cam.viewmatrix = cam.world_view_transform.data_ptr(); f.cam.projmatrix = cam.full_proj_transform.data_ptr()
        f.cam.campos = cam.camera_center.data_ptr(); f.cam.bg = background.data_ptr()
        f.gt_image = gt_image.data_ptr()
        f.loss_scale = float(loss_scale)
        f.gt_ready_event = gt_ready.cuda_event if gt_ready is not None else None
        f.stream = torch.cuda.current_stream().cuda_stream
        keep[-6:] = [cam.world_view_transform, cam.full_proj_transform, cam.camera_center, background, gt_image,
                     gt_ready]
    else:
        keep: list = []
        f = make_frame(ws, cam, background, L, True, params.sh_degree, light, brdf_lut, rays, gt_image, gi, indirect,
                       metallic, tone, gamma, loss_scale, lamb_weight, keep, gt_ready=gt_ready,
                       brdf_tv_weight=brdf_tv_weight, material_only=not radiance, skip_geometry=not radiance)
        keep.extend([None] * 6)
        # cache only pointer-stable frames: had make_frame needed a contiguous / float32 COPY of a parameter, the copy
        # would go stale as soon as the optimiser updates the original in place
        ws._frame_cache = (key, f, keep) if _is_f32c(*L.values(), *spec, light.diffuse, brdf_lut, rays) else None
    loss = frame_forward(ws, f)
    if backward:
        frame_backward(ws, f, _grad_of(L["albedo"]), _grad_of(L["roughness"]),
                       _grad_of(L["metallic"]) if metallic else None, _grad_of(light.diffuse),
                       [_grad_of(t) for t in light.specular], light_ready=light_ready)
    params.last_workspace = ws
    return loss.clone()


# ------------------------------------------------------------------------------------------------------------------
# The fused FIRST-STAGE frame (gigs_stage1_forward / gigs_stage1_backward, csrc/stage1.cu)
# ------------------------------------------------------------------------------------------------------------------
_S1_PLANES = dict(color=3, opacity=1, depth=1, normal=3, normal_view=3, pos=3, albedo=3, roughness=1, metallic=1,
                  normal_from_depth=3, depth_pos=3, normals_view=3, nfd_unit=3, g_color=3, g_normals_view=3, g_normal=3)


class Stage1Maps:
    """The maps blob of the first-stage frame for one (W, H), carved by gigs_stage1_layout."""

    def __init__(self, W: int, H: int, device):
        self.W, self.H = W, H
        self.layout = _lib.GigsStage1Layout()
        check(_L.gigs_stage1_layout(W, H, C.byref(self.layout)), "gigs_stage1_layout")
        self.blob = torch.empty(self.layout.total_bytes, dtype=torch.uint8, device=device)

    def map(self, name: str) -> torch.Tensor:
        off = getattr(self.layout, name)
        N = self.W * self.H
        if name in _S1_PLANES:
            c = _S1_PLANES[name]
            return self.blob[off:off + 4 * c * N].view(torch.float32).view(c, self.H, self.W)
        if name == "mask":
            return self.blob[off:off + N].view(self.H, self.W)
        if name == "median_sel":
            return self.blob[off:off + 3 * N].view(3, self.H, self.W)
        if name == "stats":
            return self.blob[off:off + 32].view(torch.float32)
        raise KeyError(name)


def stage1_frame_step(params, cam, gt_image, background, lambda_dssim: float = 0.2, normal_weight: float = 1.0,
                      normal_tv_weight: float = 1.0, loss_scale: float = 1.0, backward: bool = True,
                      gt_ready: Optional[torch.cuda.Event] = None, want_means2D: bool = True):
    """forward (+ backward) of one FIRST-stage view for a gigs.step.GaussianParams: gradients of all ten parameter
    groups accumulate into params.flat_grad exactly as autograd would through the operator path. Returns
    (loss, means2D_grad [P,3] or None, radii [P] int32)."""
    L = params.leaves
    dev = L["xyz"].device
    W, H = int(cam.image_width), int(cam.image_height)
    ws = workspace(params.P, W, H, dev)
    if getattr(ws, "s1", None) is None:
        ws.s1 = Stage1Maps(W, H, dev)
        ws.s1_means2D = torch.zeros((params.P, 3), dtype=torch.float32, device=dev)
    keep = []

    def c32(t):
        t = t.detach()
        if t.dtype != torch.float32 or not t.is_contiguous():
            t = t.float().contiguous()
        keep.append(t)
        return t

    f = _lib.GigsStage1()
    f.P = params.P
    vm, pm, cp, bgc = c32(cam.world_view_transform), c32(cam.full_proj_transform), c32(cam.camera_center), c32(background)
    M = 1 + L["f_rest"].shape[1]
    f.cam = GigsCamera(W, H, float(cam.tanfovx), float(cam.tanfovy), 1.0, int(params.sh_degree), int(M), 0, 0, 0, 0,
                       _p(vm), _p(pm), _p(cp), _p(bgc))
    for k in ("xyz", "f_dc", "f_rest", "opacity", "normal", "albedo", "roughness", "metallic", "log_scale", "rot"):
        if not _is_f32c(L[k]):
            raise RuntimeError("first-stage frame: parameters must be contiguous float32 CUDA tensors")
        setattr(f, k, L[k].data_ptr())
    f.gt_image = _p(c32(gt_image)) if gt_image is not None else None
    f.lambda_dssim, f.normal_weight = float(lambda_dssim), float(normal_weight)
    f.normal_tv_weight, f.loss_scale = float(normal_tv_weight), float(loss_scale)
    f.geom = ws.geom.data_ptr(); f.geom_bytes = ws.geom.numel()
    f.img = ws.img.data_ptr(); f.img_bytes = ws.img.numel()
    f.maps = ws.s1.blob.data_ptr(); f.maps_bytes = ws.s1.blob.numel()
    f.radii = ws.radii.data_ptr(); f.accum = ws.accum.data_ptr()
    f.pinned_num_rendered = ws.pinned.data_ptr()
    if gt_ready is not None:
        keep.append(gt_ready)
        f.gt_ready_event = gt_ready.cuda_event
    f.stream = torch.cuda.current_stream().cuda_stream
    with torch.cuda.device(dev):
        f.binning = _p(ws.binning); f.binning_bytes = 0 if ws.binning is None else ws.binning.numel()
        f.sort = _p(ws.sort); f.sort_bytes = 0 if ws.sort is None else ws.sort.numel()
        f.resume = 0
        st = _L.gigs_stage1_forward(C.byref(f))
        if st == GIGS_E_GROW:
            ws.grow(f.need_binning_bytes, f.need_sort_bytes)
            f.binning = _p(ws.binning); f.binning_bytes = ws.binning.numel()
            f.sort = _p(ws.sort); f.sort_bytes = ws.sort.numel()
            f.resume = 1
            st = _L.gigs_stage1_forward(C.byref(f))
        check(st, "gigs_stage1_forward")
        ws.num_rendered = int(f.num_rendered)
        stats = ws.s1.map("stats")
        loss = (stats[0] + stats[4]) if gt_image is not None else None
        if backward and gt_image is not None:
            for k in ("xyz", "f_dc", "f_rest", "opacity", "normal", "albedo", "roughness", "metallic", "log_scale", "rot"):
                setattr(f, "g_" + k, _grad_of(L[k]).data_ptr())
            f.g_means2D = ws.s1_means2D.data_ptr() if want_means2D else None
            check(_L.gigs_stage1_backward(C.byref(f)), "gigs_stage1_backward")
    params.last_workspace = ws
    return loss, (ws.s1_means2D if (backward and want_means2D) else None), ws.radii
