"""Environment light: host-side mirror of the reference's `CubemapLight` (/root/reference/pbr/light.py:81-170) and of
the nvdiffrec `renderutils` cubemap ops it calls (/root/reference/pbr/renderutils/ops.py:391-459), backed by the
C-ABI entry points gigs_cubemap_* / gigs_diffuse_cubemap_* / gigs_specular_cubemap_* (csrc/cubemap.cu). SURVEY §8f-1.

Same names and meanings: `CubemapLight(base_res, scale, bias)`, `.base` (the trainable [6,res,res,3] texture),
`.build_mips(cutoff)`, `.specular` (list, fine to coarse), `.diffuse`, `.get_mip`, `MIN_ROUGHNESS` / `MAX_ROUGHNESS`,
`.clamp_`, plus the three ops `cubemap_mip`, `diffuse_cubemap`, `specular_cubemap` as differentiable functions.

`PrefilteredLight` is the same light for the fused training frame (gigs.frame): build_mips as 2 launches, its backward
as 6, every level in one workspace blob, no autograd graph.
"""
import ctypes as C
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch
import torch.nn as nn

from . import _lib
from ._lib import GigsLightLayout, check

_L = _lib.load()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _cube(t: torch.Tensor, name: str) -> torch.Tensor:
    if t.dim() != 4 or t.shape[0] != 6 or t.shape[1] != t.shape[2] or t.shape[3] != 3:
        raise RuntimeError(f"Bad shape for {name}: {tuple(t.shape)} (expected [6,res,res,3])")
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor")
    return t.float().contiguous()


# ---- per-resolution constants: texel directions + solid-angle weights, and the GGX cone bounds ------------------
_tables: Dict[Tuple[int, int], torch.Tensor] = {}
_bounds: Dict[Tuple[int, float, float, int], Tuple[float, torch.Tensor]] = {}


def texel_table(res: int, device) -> torch.Tensor:
    """[6,res,res,4]: unit direction of each texel centre and its pixel_area (c_src/cubemap.cu:17-47)."""
    key = (res, torch.device(device).index or 0)
    t = _tables.get(key)
    if t is None:
        t = torch.empty((6, res, res, 4), dtype=torch.float32, device=device)
        with torch.cuda.device(device):
            check(_L.gigs_cubemap_table(res, t.data_ptr(), _stream()), "gigs_cubemap_table")
        _tables[key] = t
    return t


def ndf_cutoff(roughness: float, cutoff: float) -> float:
    """cos(theta) that keeps `cutoff` of the GGX NDF's energy (renderutils/ops.py:430-446, __ndfBounds; numpy, same
    1e6-sample cumulative sum)."""
    def ndf_ggx(alpha_sqr, costheta):
        costheta = np.clip(costheta, 0.0, 1.0)
        d = (costheta * alpha_sqr - costheta) * costheta + 1.0
        return alpha_sqr / (d * d * np.pi)
    n = 1000000
    costheta = np.cos(np.linspace(0, np.pi / 2.0, n))
    D = np.cumsum(ndf_ggx(roughness ** 4, costheta))
    idx = np.argmax(D >= D[..., -1] * cutoff)
    return float(costheta[idx])


def specular_bounds(res: int, roughness: float, cutoff: float, device) -> Tuple[float, torch.Tensor]:
    """(costheta_cutoff, bounds int16 [6,res,res,6,4]); cached per (res, roughness, cutoff) like __ndfBoundsDict."""
    key = (res, float(roughness), float(cutoff), torch.device(device).index or 0)
    hit = _bounds.get(key)
    if hit is None:
        c = ndf_cutoff(roughness, cutoff)
        b = torch.empty((6, res, res, 6, 4), dtype=torch.int16, device=device)
        with torch.cuda.device(device):
            check(_L.gigs_specular_bounds(res, c, texel_table(res, device).data_ptr(), b.data_ptr(), _stream()),
                  "gigs_specular_bounds")
        hit = _bounds[key] = (c, b)
    return hit


# ---- the three ops ------------------------------------------------------------------------------------------------
class _CubemapMip(torch.autograd.Function):
    """pbr/light.py:54-79: forward 2x2 average pool; backward = seamless bilinear lookup of 0.25*dout at the fine
    texel-centre directions (the reference's choice, not the adjoint of the pool)."""

    @staticmethod
    def forward(ctx, cubemap):
        x = _cube(cubemap, "cubemap")
        r = x.shape[1] // 2
        out = torch.empty((6, r, r, 3), dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            check(_L.gigs_cubemap_mip_forward(r, x.data_ptr(), out.data_ptr(), _stream()), "gigs_cubemap_mip_forward")
        return out

    @staticmethod
    def backward(ctx, dout):
        g = _cube(dout, "dout")
        r = g.shape[1]
        out = torch.empty((6, 2 * r, 2 * r, 3), dtype=torch.float32, device=g.device)
        with torch.cuda.device(g.device):
            check(_L.gigs_cubemap_mip_backward(r, g.data_ptr(), out.data_ptr(), 0, _stream()),
                  "gigs_cubemap_mip_backward")
        return out


class _DiffuseCubemap(torch.autograd.Function):
    @staticmethod
    def forward(ctx, cubemap):
        x = _cube(cubemap, "cubemap")
        r = x.shape[1]
        out = torch.empty_like(x)
        with torch.cuda.device(x.device):
            check(_L.gigs_diffuse_cubemap_forward(r, texel_table(r, x.device).data_ptr(), x.data_ptr(), out.data_ptr(),
                                                  _stream()), "gigs_diffuse_cubemap_forward")
        return out

    @staticmethod
    def backward(ctx, dout):
        g = _cube(dout, "dout")
        r = g.shape[1]
        gin = torch.empty_like(g)
        with torch.cuda.device(g.device):
            check(_L.gigs_diffuse_cubemap_backward(r, texel_table(r, g.device).data_ptr(), g.data_ptr(), gin.data_ptr(),
                                                   _stream()), "gigs_diffuse_cubemap_backward")
        return gin


class _SpecularCubemap(torch.autograd.Function):
    @staticmethod
    def forward(ctx, cubemap, roughness, costheta_cutoff, bounds):
        x = _cube(cubemap, "cubemap")
        r = x.shape[1]
        out = torch.empty_like(x)
        wsum = torch.empty((6, r, r), dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            check(_L.gigs_specular_cubemap_forward(r, texel_table(r, x.device).data_ptr(), bounds.data_ptr(),
                                                   float(roughness), float(costheta_cutoff), x.data_ptr(),
                                                   out.data_ptr(), wsum.data_ptr(), _stream()),
                  "gigs_specular_cubemap_forward")
        ctx.save_for_backward(wsum, bounds)
        ctx.cfg = (float(roughness), float(costheta_cutoff))
        return out

    @staticmethod
    def backward(ctx, dout):
        wsum, bounds = ctx.saved_tensors
        g = _cube(dout, "dout")
        r = g.shape[1]
        gin = torch.empty_like(g)
        with torch.cuda.device(g.device):
            check(_L.gigs_specular_cubemap_backward(r, texel_table(r, g.device).data_ptr(), bounds.data_ptr(),
                                                    ctx.cfg[0], ctx.cfg[1], g.data_ptr(), wsum.data_ptr(),
                                                    gin.data_ptr(), _stream()), "gigs_specular_cubemap_backward")
        return gin, None, None, None


def cubemap_mip(cubemap: torch.Tensor) -> torch.Tensor:
    return _CubemapMip.apply(cubemap)


def diffuse_cubemap(cubemap: torch.Tensor) -> torch.Tensor:
    """renderutils.diffuse_cubemap (ops.py:405-413)."""
    return _DiffuseCubemap.apply(cubemap)


def specular_cubemap(cubemap: torch.Tensor, roughness: float, cutoff: float = 0.99) -> torch.Tensor:
    """renderutils.specular_cubemap (ops.py:446-459): the GGX-filtered cubemap, already divided by the weight sum."""
    if cubemap.shape[0] != 6 or cubemap.shape[1] != cubemap.shape[2]:
        raise AssertionError("Bad shape for cubemap tensor: %s" % str(cubemap.shape))
    c, b = specular_bounds(cubemap.shape[1], roughness, cutoff, cubemap.device)
    return _SpecularCubemap.apply(cubemap, roughness, c, b)


class CubemapLight(nn.Module):
    """pbr/light.py:81-170. `base` is the only parameter; build_mips() derives `specular` (len = log2(base_res/16)+1
    levels, each GGX-filtered for its roughness) and `diffuse` from it, differentiably."""
    LIGHT_MIN_RES = 16
    MIN_ROUGHNESS = 0.08
    MAX_ROUGHNESS = 0.5

    def __init__(self, base_res: int = 16, scale: float = 0.5, bias: float = 0.25, device="cuda",
                 base: Optional[torch.Tensor] = None) -> None:
        super().__init__()
        self.mtx = None
        if base is None:
            base = torch.rand(6, base_res, base_res, 3, dtype=torch.float32, device=device) * scale + bias
        self.base = nn.Parameter(base.to(device).float().contiguous())
        self.register_parameter("env_base", self.base)
        self.specular: List[torch.Tensor] = []
        self.diffuse: Optional[torch.Tensor] = None

    def xfm(self, mtx) -> None:
        self.mtx = mtx

    def clamp_(self, min: Optional[float] = None, max: Optional[float] = None) -> None:
        self.base.data.clamp_(min, max)

    def get_mip(self, roughness: torch.Tensor) -> torch.Tensor:
        n = len(self.specular)
        return torch.where(
            roughness < self.MAX_ROUGHNESS,
            (torch.clamp(roughness, self.MIN_ROUGHNESS, self.MAX_ROUGHNESS) - self.MIN_ROUGHNESS)
            / (self.MAX_ROUGHNESS - self.MIN_ROUGHNESS) * (n - 2),
            (torch.clamp(roughness, self.MAX_ROUGHNESS, 1.0) - self.MAX_ROUGHNESS) / (1.0 - self.MAX_ROUGHNESS) + n - 2)

    def build_mips(self, cutoff: float = 0.99) -> None:
        self.specular = [self.base]
        while self.specular[-1].shape[1] > self.LIGHT_MIN_RES:
            self.specular += [cubemap_mip(self.specular[-1])]
        self.diffuse = diffuse_cubemap(self.specular[-1])
        for idx in range(len(self.specular) - 1):
            roughness = (idx / (len(self.specular) - 2)) * (self.MAX_ROUGHNESS - self.MIN_ROUGHNESS) + self.MIN_ROUGHNESS
            self.specular[idx] = specular_cubemap(self.specular[idx], roughness, cutoff)
        self.specular[-1] = specular_cubemap(self.specular[-1], 1.0, cutoff)


def latlong_to_cubemap(latlong_map: torch.Tensor, res) -> torch.Tensor:
    """relight.py:92-112 / render.py:64-84: [H, W, C] lat-long environment map -> [6, res, res, C] cubemap (one launch;
    the reference loops over the faces with meshgrid / normalize / atan2 / acos / a texture lookup each)."""
    r = int(res[0]) if isinstance(res, (list, tuple)) else int(res)
    if isinstance(res, (list, tuple)) and int(res[1]) != r:
        raise RuntimeError("latlong_to_cubemap: square faces only")
    if latlong_map.dim() != 3 or not latlong_map.is_cuda:
        raise RuntimeError("latlong_to_cubemap: expected a CUDA tensor [H, W, C]")
    env = latlong_map.detach().float().contiguous()
    EH, EW, Cn = env.shape
    cube = torch.empty(6, r, r, Cn, dtype=torch.float32, device=env.device)
    with torch.cuda.device(env.device):
        check(_L.gigs_latlong_to_cubemap(EH, EW, Cn, env.data_ptr(), r, cube.data_ptr(), _stream()),
              "gigs_latlong_to_cubemap")
    return cube


def envmap_dirs(res=(512, 1024), device="cuda") -> torch.Tensor:
    """train.py:145-157 get_envmap_dirs: [H,W,3] directions of a lat-long grid (computed once, like train.py:209)."""
    gy, gx = torch.meshgrid(torch.linspace(0.0 + 1.0 / res[0], 1.0 - 1.0 / res[0], res[0], device=device),
                            torch.linspace(-1.0 + 1.0 / res[1], 1.0 - 1.0 / res[1], res[1], device=device),
                            indexing="ij")
    sintheta, costheta = torch.sin(gy * np.pi), torch.cos(gy * np.pi)
    sinphi, cosphi = torch.sin(gx * np.pi), torch.cos(gx * np.pi)
    return torch.stack((sintheta * sinphi, costheta, -sintheta * cosphi), dim=-1).contiguous()


_env_scratch: Dict = {}


def env_tv_fused(base: torch.Tensor, dirs: torch.Tensor, scale: float, grad_base: Optional[torch.Tensor] = None,
                 loss_out: Optional[torch.Tensor] = None, accumulate_loss: bool = False) -> Optional[torch.Tensor]:
    """gigs_env_tv: loss_out (+)= scale * env_tv(base), grad_base += scale * d env_tv / d base (either may be None)."""
    EH, EW = int(dirs.shape[0]), int(dirs.shape[1])
    key = (EH, EW, base.device.index or 0)
    sc = _env_scratch.get(key)
    if sc is None:
        need = C.c_uint64(0)
        check(_L.gigs_env_tv(int(base.shape[1]), None, None, EH, EW, 0.0, None, C.byref(need), None, None, 0, None),
              "gigs_env_tv")
        sc = _env_scratch[key] = torch.empty(need.value, dtype=torch.uint8, device=base.device)
    nb = C.c_uint64(sc.numel())
    with torch.cuda.device(base.device):
        check(_L.gigs_env_tv(int(base.shape[1]), base.data_ptr(), dirs.data_ptr(), EH, EW, float(scale), sc.data_ptr(),
                             C.byref(nb), None if grad_base is None else grad_base.data_ptr(),
                             None if loss_out is None else loss_out.data_ptr(), int(accumulate_loss), _stream()),
              "gigs_env_tv")
    return loss_out


class _EnvTV(torch.autograd.Function):
    @staticmethod
    def forward(ctx, base, dirs):
        b = _cube(base, "base")
        loss = torch.zeros(1, dtype=torch.float32, device=b.device)
        env_tv_fused(b, dirs, 1.0, None, loss)
        ctx.save_for_backward(b, dirs)
        return loss[0]

    @staticmethod
    def backward(ctx, g):
        b, dirs = ctx.saved_tensors
        gb = torch.zeros_like(b)
        env_tv_fused(b, dirs, 1.0, gb, None)
        return gb * g, None


def env_tv_loss(base: torch.Tensor, dirs: torch.Tensor) -> torch.Tensor:
    """train.py:406-420: tv_h1 + tv_w1 of the lat-long unwrapping of the base cubemap (differentiable in `base`)."""
    return _EnvTV.apply(base, dirs.float().contiguous())


class PrefilteredLight:
    """build_mips (pbr/light.py:154-170) + its backward through gigs_light_build / gigs_light_backward.

    `base` is the trainable [6,R,R,3] texture (R a power of two: 16, or >= 64 — two levels divide by zero in the
    reference's roughness schedule too). `specular` (fine to coarse) and `diffuse` are views into the workspace that
    build() refreshes in place; they look like leaves whose .grad the shading backward accumulates into (the contiguous
    gradient span of the workspace). backward() turns those texture gradients into d loss / d base (+= into
    base.grad, or into `grad_out`) and clears them. Duck-types gigs.shade.Light for the frame / operator paths."""
    LIGHT_MIN_RES = 16
    MIN_ROUGHNESS = 0.08
    MAX_ROUGHNESS = 0.5

    # "auto": store the operators only when they fit this share of the device's FREE memory and this absolute budget
    # (1.4 GB at base_res 256, 11.3 GB at 512: ~8x per doubling — more texels, each with a cone of more texels)
    STORED_FREE_FRACTION = 0.25
    STORED_MAX_BYTES = 8 << 30

    def __init__(self, base: torch.Tensor, cutoff: float = 0.99, stored_operators="auto"):
        """stored_operators: True evaluates the filter weights once into HBM (layout.weights_bytes, 1.4 GB at base_res
        256) and streams them every step; False recomputes them on the fly (no extra memory, ~4x slower per build:
        right for a relight sweep that builds once). "auto" (default) stores them when layout.weights_bytes is within
        STORED_FREE_FRACTION of the free device memory and STORED_MAX_BYTES, and says which it chose in
        `self.stored_operators` / `self.stored_operator_bytes`."""
        if base.dim() != 4 or base.shape[0] != 6 or base.shape[1] != base.shape[2] or base.shape[3] != 3:
            raise RuntimeError(f"Bad shape for base: {tuple(base.shape)} (expected [6,res,res,3])")
        if not base.is_cuda or base.dtype != torch.float32 or not base.is_contiguous():
            raise RuntimeError("base must be a contiguous float32 CUDA tensor")
        self.base = base
        self.device = base.device
        self.layout = GigsLightLayout()
        n, r = 1, base.shape[1]
        while r > self.LIGHT_MIN_RES:
            r //= 2
            n += 1
        if n == 2:
            raise ZeroDivisionError("float division by zero")   # pbr/light.py:166 with two levels
        check(_L.gigs_light_layout(int(base.shape[1]), self.LIGHT_MIN_RES, C.byref(self.layout)), "gigs_light_layout")
        lay = self.layout
        for i in range(lay.n_levels):
            # the reference evaluates the schedule in python doubles (light.py:165-170) before __ndfBounds
            rough = 1.0 if i == lay.n_levels - 1 else \
                (i / (lay.n_levels - 2)) * (self.MAX_ROUGHNESS - self.MIN_ROUGHNESS) + self.MIN_ROUGHNESS
            lay.cutoff[i] = ndf_cutoff(rough, cutoff)
        self.ws = torch.empty(lay.total_bytes, dtype=torch.uint8, device=self.device)
        with torch.cuda.device(self.device):
            check(_L.gigs_light_prepare(C.byref(lay), self.ws.data_ptr(), _stream()), "gigs_light_prepare")
            self.weights = None
            if stored_operators == "auto":
                free, _total = torch.cuda.mem_get_info(self.device)
                stored_operators = (int(lay.weights_bytes) <= self.STORED_MAX_BYTES
                                    and int(lay.weights_bytes) <= self.STORED_FREE_FRACTION * free)
            self.stored_operators = bool(stored_operators)
            self.stored_operator_bytes = int(lay.weights_bytes) if stored_operators else 0
            if stored_operators:
                self.weights = torch.empty(lay.weights_bytes, dtype=torch.uint8, device=self.device)
                check(_L.gigs_light_weights(C.byref(lay), self.ws.data_ptr(), self.weights.data_ptr(), _stream()),
                      "gigs_light_weights")

        def view(off, res, ch=3):
            return self.ws[off:off + 4 * 6 * res * res * ch].view(torch.float32).view(6, res, res, ch) if ch > 1 else \
                self.ws[off:off + 4 * 6 * res * res].view(torch.float32).view(6, res, res)

        self.specular: List[torch.Tensor] = []
        for i in range(lay.n_levels):
            t = view(lay.spec[i], lay.res[i]).requires_grad_(True)
            t.grad = view(lay.g_spec[i], lay.res[i])
            self.specular.append(t)
        rd = lay.res[lay.n_levels - 1]
        self.diffuse = view(lay.diffuse, rd).requires_grad_(True)
        self.diffuse.grad = view(lay.g_diffuse, rd)
        self.wsum = [view(lay.wsum[i], lay.res[i], 1) for i in range(lay.n_levels)]
        self.chain = [self.ws[lay.chain[i]:lay.chain[i] + 16 * 6 * lay.res[i] ** 2].view(torch.float32)
                      .view(6, lay.res[i], lay.res[i], 4) for i in range(lay.n_levels)]
        self.texture_grads = self.ws[lay.grad_begin:lay.grad_begin + lay.grad_bytes].view(torch.float32)

    def get_mip(self, roughness: torch.Tensor) -> torch.Tensor:
        return CubemapLight.get_mip(self, roughness)

    def build(self) -> "PrefilteredLight":
        with torch.cuda.device(self.device):
            check(_L.gigs_light_build(C.byref(self.layout), self.base.data_ptr(), self.ws.data_ptr(),
                                      None if self.weights is None else self.weights.data_ptr(), _stream()),
                  "gigs_light_build")
        return self

    build_mips = build

    def backward(self, grad_out: Optional[torch.Tensor] = None, accumulate: bool = True, clear: bool = True):
        """d loss / d base from the texture gradients accumulated since the last clear."""
        if grad_out is None:
            if self.base.grad is None:
                self.base.grad = torch.zeros_like(self.base)
            grad_out = self.base.grad
        if grad_out.dtype != torch.float32 or not grad_out.is_contiguous() or grad_out.numel() != self.base.numel():
            raise RuntimeError("PrefilteredLight.backward: grad_out must be contiguous float32 of base's size")
        with torch.cuda.device(self.device):
            check(_L.gigs_light_backward(C.byref(self.layout), self.ws.data_ptr(),
                                         None if self.weights is None else self.weights.data_ptr(), grad_out.data_ptr(),
                                         int(accumulate), int(clear), _stream()), "gigs_light_backward")
        return grad_out
