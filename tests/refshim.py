"""TEST INFRASTRUCTURE: ctypes access to the *unmodified reference CUDA kernels* compiled into
oracle/_ref/libgigs_ref.so by oracle/Makefile (C-ABI shim oracle/ref_shim.cu). Only tests/, bench.py's
reference arms and __graft_entry__.smoke() may import this; the product never does.
Needs a GPU to run anything; loading works on CPU.
"""
import ctypes as C
import os

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libgigs_ref.so")

_lib = None


def available() -> bool:
    return os.path.exists(REF_SO)


def load():
    global _lib
    if _lib is None:
        _lib = C.CDLL(REF_SO)
        _lib.ref_ctx_create.restype = C.c_void_p
        _lib.ref_last_error.restype = C.c_char_p
        _lib.ref_forward.restype = C.c_int
        _lib.ref_backward.restype = C.c_int
        _lib.ref_lite_forward.restype = C.c_int
    return _lib


def _p(t):
    if t is None or t.numel() == 0:
        return C.c_void_p(0)
    return C.c_void_p(t.data_ptr())


class RefRasterizer:
    """One reference context (its three cudaMalloc'd workspaces persist between forward and backward)."""

    def __init__(self):
        self.lib = load()
        self.ctx = C.c_void_p(self.lib.ref_ctx_create())
        self.P = self.R = 0

    def close(self):
        if self.ctx:
            self.lib.ref_ctx_destroy(self.ctx)
            self.ctx = None

    def lite_forward(self, g, cam, bg, sh_degree=3, scale_modifier=1.0, argmax_depth=False, prefiltered=False,
                     colors_precomp=None, cov3D_precomp=None):
        """The reference's radiance-only rasterizer (Rasterizer::lite_forward): colour, opacity, depth, radii."""
        dev = g["means3D"].device
        P = g["means3D"].shape[0]
        H, W = cam.image_height, cam.image_width
        f = dict(dtype=torch.float32, device=dev)
        out = dict(color=torch.zeros(3, H, W, **f), opacity=torch.zeros(1, H, W, **f), depth=torch.zeros(1, H, W, **f),
                   radii=torch.zeros(P, dtype=torch.int32, device=dev))
        shs = None if colors_precomp is not None else g["shs"]
        M = shs.shape[1] if shs is not None else 0
        scales = None if cov3D_precomp is not None else g["scales"]
        rots = None if cov3D_precomp is not None else g["rotations"]
        keep = [t.contiguous() if t is not None else None for t in
                (bg, g["means3D"], shs, colors_precomp, g["opacity"], scales, rots, cov3D_precomp,
                 cam.world_view_transform, cam.full_proj_transform, cam.camera_center)]
        (bg_, m3, sh_, cp_, op_, sc_, rt_, cv_, vm_, pm_, cc_) = keep
        torch.cuda.synchronize()
        R = self.lib.ref_lite_forward(self.ctx, C.c_int(P), C.c_int(sh_degree), C.c_int(M), _p(bg_), C.c_int(W),
                                      C.c_int(H), _p(m3), _p(sh_), _p(cp_), _p(op_), _p(sc_), C.c_float(scale_modifier),
                                      _p(rt_), _p(cv_), _p(vm_), _p(pm_), _p(cc_), C.c_float(cam.tanfovx),
                                      C.c_float(cam.tanfovy), C.c_int(int(prefiltered)), C.c_int(int(argmax_depth)),
                                      _p(out["color"]), _p(out["opacity"]), _p(out["depth"]), _p(out["radii"]))
        if R < 0:
            raise RuntimeError("reference lite_forward failed: " + self.lib.ref_last_error().decode())
        torch.cuda.synchronize()
        out["num_rendered"] = R
        return out

    def forward(self, g, cam, bg, sh_degree=3, scale_modifier=1.0, inference=False, argmax_depth=False,
                colors_precomp=None, cov3D_precomp=None, debug=False, prefiltered=False):
        dev = g["means3D"].device
        P = g["means3D"].shape[0]
        H, W = cam.image_height, cam.image_width
        f = dict(dtype=torch.float32, device=dev)
        out = dict(color=torch.zeros(3, H, W, **f), opacity=torch.zeros(1, H, W, **f), depth=torch.zeros(1, H, W, **f),
                   normal=torch.zeros(3, H, W, **f), normal_view=torch.zeros(3, H, W, **f),
                   pos=torch.zeros(3, H, W, **f), albedo=torch.zeros(3, H, W, **f),
                   roughness=torch.zeros(1, H, W, **f), metallic=torch.zeros(1, H, W, **f),
                   radii=torch.zeros(P, dtype=torch.int32, device=dev))
        shs = None if colors_precomp is not None else g["shs"]
        M = shs.shape[1] if shs is not None else 0
        scales = None if cov3D_precomp is not None else g["scales"]
        rots = None if cov3D_precomp is not None else g["rotations"]
        self._keep = [t.contiguous() if t is not None else None for t in
                      (bg, g["means3D"], shs, colors_precomp, g["opacity"], g["normal"], g["albedo"], g["roughness"],
                       g["metallic"], scales, rots, cov3D_precomp, cam.world_view_transform, cam.full_proj_transform,
                       cam.camera_center)]
        (bg_, m3, sh_, cp_, op_, nr_, al_, ro_, me_, sc_, rt_, cv_, vm_, pm_, cc_) = self._keep
        torch.cuda.synchronize()
        R = self.lib.ref_forward(self.ctx, C.c_int(P), C.c_int(sh_degree), C.c_int(M), _p(bg_), C.c_int(W), C.c_int(H),
                                 _p(m3), _p(sh_), _p(cp_), _p(op_), _p(nr_), _p(al_), _p(ro_), _p(me_), _p(sc_),
                                 C.c_float(scale_modifier), _p(rt_), _p(cv_), _p(vm_), _p(pm_), _p(cc_),
                                 C.c_float(cam.tanfovx), C.c_float(cam.tanfovy), C.c_int(int(prefiltered)),
                                 C.c_int(int(argmax_depth)), C.c_int(int(inference)), _p(out["color"]), _p(out["opacity"]), _p(out["depth"]),
                                 _p(out["normal"]), _p(out["normal_view"]), _p(out["pos"]), _p(out["albedo"]),
                                 _p(out["roughness"]), _p(out["metallic"]), _p(out["radii"]), C.c_int(int(debug)))
        if R < 0:
            raise RuntimeError("reference forward failed: " + self.lib.ref_last_error().decode())
        torch.cuda.synchronize()
        self.P, self.R, self.W, self.H, self.M, self.D = P, R, W, H, M, sh_degree
        self.scale_modifier = scale_modifier
        out["num_rendered"] = R
        return out

    def state(self):
        """Decode the reference's geom / binning / img blobs of the last forward into torch tensors (copies)."""
        ptrs = (C.c_void_p * 16)()
        self.lib.ref_state_ptrs(self.ctx, ptrs)
        P, R, N = self.P, self.R, self.W * self.H
        T = ((self.W + 15) // 16) * ((self.H + 15) // 16)

        def grab(i, nbytes, dtype, shape):
            buf = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
            if nbytes:
                self.lib.ref_memcpy_d2d(C.c_void_p(buf.data_ptr()), C.c_void_p(ptrs[i]), C.c_size_t(nbytes))
            return buf.view(dtype).reshape(shape)

        s = dict(
            depths=grab(0, 4 * P, torch.float32, (P,)), pos_view=grab(1, 12 * P, torch.float32, (P, 3)),
            clamped=grab(2, 3 * P, torch.uint8, (P, 3)), means2D=grab(3, 8 * P, torch.float32, (P, 2)),
            cov3D=grab(4, 24 * P, torch.float32, (P, 6)), conic_opacity=grab(5, 16 * P, torch.float32, (P, 4)),
            rgb=grab(6, 12 * P, torch.float32, (P, 3)), tiles_touched=grab(7, 4 * P, torch.int32, (P,)),
            point_offsets=grab(8, 4 * P, torch.int32, (P,)),
            keys_unsorted=grab(9, 8 * R, torch.int64, (R,)), keys_sorted=grab(10, 8 * R, torch.int64, (R,)),
            vals_unsorted=grab(11, 4 * R, torch.int32, (R,)), point_list=grab(12, 4 * R, torch.int32, (R,)),
            final_T=grab(13, 4 * N, torch.float32, (N,)), n_contrib=grab(14, 4 * N, torch.int32, (N,)),
            ranges=grab(15, 8 * T, torch.int32, (T, 2)),
        )
        torch.cuda.synchronize()
        return s

    def backward(self, g, cam, bg, radii, grads, colors_precomp=None, cov3D_precomp=None):
        """grads: dict with depth,color,opacity,normal,albedo,roughness,metallic maps (all required, like the
        reference's materialised zeros). Returns the 12 gradient tensors + dL_dconic."""
        dev = g["means3D"].device
        P, M = self.P, self.M
        f = dict(dtype=torch.float32, device=dev)
        o = dict(means2D=torch.zeros(P, 3, **f), conic=torch.zeros(P, 2, 2, **f), depth=torch.zeros(P, 1, **f),
                 opacity=torch.zeros(P, 1, **f), normal=torch.zeros(P, 3, **f), albedo=torch.zeros(P, 3, **f),
                 roughness=torch.zeros(P, 1, **f), metallic=torch.zeros(P, 1, **f), colors=torch.zeros(P, 3, **f),
                 means3D=torch.zeros(P, 3, **f), cov3D=torch.zeros(P, 6, **f), sh=torch.zeros(P, M, 3, **f),
                 scales=torch.zeros(P, 3, **f), rotations=torch.zeros(P, 4, **f))
        shs = None if colors_precomp is not None else g["shs"]
        scales = None if cov3D_precomp is not None else g["scales"]
        rots = None if cov3D_precomp is not None else g["rotations"]
        keep = [t.contiguous() if t is not None else None for t in
                (bg, g["means3D"], shs, colors_precomp, g["normal"], g["albedo"], g["roughness"], g["metallic"],
                 scales, rots, cov3D_precomp, cam.world_view_transform, cam.full_proj_transform, cam.camera_center,
                 radii, grads["depth"], grads["color"], grads["opacity"], grads["normal"], grads["albedo"],
                 grads["roughness"], grads["metallic"])]
        (bg_, m3, sh_, cp_, nr_, al_, ro_, me_, sc_, rt_, cv_, vm_, pm_, cc_, rad_, gd, gc, go, gn, ga, gr, gm) = keep
        torch.cuda.synchronize()
        rc = self.lib.ref_backward(self.ctx, C.c_int(P), C.c_int(self.D), C.c_int(M), C.c_int(self.R), _p(bg_),
                                   C.c_int(self.W), C.c_int(self.H), _p(m3), _p(sh_), _p(cp_), _p(nr_), _p(al_),
                                   _p(ro_), _p(me_), _p(sc_), _p(rt_), _p(cv_), _p(vm_), _p(pm_), _p(cc_), _p(rad_),
                                   C.c_float(self.scale_modifier), C.c_float(cam.tanfovx), C.c_float(cam.tanfovy),
                                   _p(gd), _p(gc), _p(go), _p(gn), _p(ga), _p(gr), _p(gm),
                                   _p(o["means2D"]), _p(o["conic"]), _p(o["depth"]), _p(o["opacity"]), _p(o["normal"]),
                                   _p(o["albedo"]), _p(o["roughness"]), _p(o["metallic"]), _p(o["colors"]),
                                   _p(o["means3D"]), _p(o["cov3D"]), _p(o["sh"]), _p(o["scales"]), _p(o["rotations"]),
                                   C.c_int(0))
        if rc != 0:
            raise RuntimeError("reference backward failed: " + self.lib.ref_last_error().decode())
        torch.cuda.synchronize()
        return o


def depth_to_normal(W, H, fx, fy, viewmatrix, depth):
    lib = load()
    n = torch.zeros(3, H, W, dtype=torch.float32, device=depth.device)
    p = torch.zeros(3, H, W, dtype=torch.float32, device=depth.device)
    d, v = depth.contiguous(), viewmatrix.contiguous()
    torch.cuda.synchronize()
    lib.ref_depth_to_normal(C.c_int(W), C.c_int(H), C.c_float(fx), C.c_float(fy), _p(v), _p(d), _p(n), _p(p))
    torch.cuda.synchronize()
    return n, p


def ssao(W, H, fx, fy, radius, bias, thick, delta, step, start, normal, pos):
    lib = load()
    occ = torch.ones(1, H, W, dtype=torch.float32, device=pos.device)
    n, p = normal.contiguous(), pos.contiguous()
    torch.cuda.synchronize()
    lib.ref_ssao(C.c_int(W), C.c_int(H), C.c_float(fx), C.c_float(fy), C.c_float(radius), C.c_float(bias),
                 C.c_float(thick), C.c_float(delta), C.c_int(step), C.c_int(start), _p(n), _p(p), _p(occ))
    torch.cuda.synchronize()
    return occ


def ssr(W, H, fx, fy, radius, bias, thick, delta, step, start, normal, pos, rgb, albedo, roughness, metallic, F0):
    lib = load()
    color = torch.zeros(3, H, W, dtype=torch.float32, device=pos.device)
    abd = torch.zeros(3, H, W, dtype=torch.float32, device=pos.device)
    ts = [t.contiguous() for t in (normal, pos, rgb, albedo, roughness, metallic, F0)]
    torch.cuda.synchronize()
    lib.ref_ssr(C.c_int(W), C.c_int(H), C.c_float(fx), C.c_float(fy), C.c_float(radius), C.c_float(bias),
                C.c_float(thick), C.c_float(delta), C.c_int(step), C.c_int(start), *[_p(t) for t in ts], _p(color),
                _p(abd))
    torch.cuda.synchronize()
    return color, abd


def mark_visible(means3D, viewmatrix, projmatrix):
    lib = load()
    P = means3D.shape[0]
    pres = torch.zeros(P, dtype=torch.bool, device=means3D.device)
    m, v, pm = means3D.contiguous(), viewmatrix.contiguous(), projmatrix.contiguous()
    torch.cuda.synchronize()
    lib.ref_mark_visible(C.c_int(P), _p(m), _p(v), _p(pm), _p(pres))
    torch.cuda.synchronize()
    return pres


def knn(points):
    lib = load()
    P = points.shape[0]
    out = torch.zeros(P, dtype=torch.float32, device=points.device)
    pts = points.contiguous()
    torch.cuda.synchronize()
    lib.ref_knn(C.c_int(P), _p(pts), _p(out))
    torch.cuda.synchronize()
    return out


# ---- the reference's cubemap prefilter kernels (pbr/renderutils/c_src/cubemap.cu via oracle/ref_cubemap_shim.cu) ----
REF_CUBEMAP_SO = os.path.join(ROOT, "oracle", "_ref", "libgigs_ref_cubemap.so")
_cm = None


def cubemap_available() -> bool:
    return os.path.exists(REF_CUBEMAP_SO)


class RefCubemap:
    """The reference's renderutils plugin entry points (torch_bindings.cpp:740-889), same names and tensor shapes:
    cubemaps [6,N,N,3], bounds float [6,N,N,24], specular_cubemap_fwd output [6,N,N,4] = (sum col*w, wsum)."""

    def __init__(self):
        global _cm
        if _cm is None:
            _cm = C.CDLL(REF_CUBEMAP_SO)
        self.lib = _cm

    @staticmethod
    def _chk(rc, what):
        if rc != 0:
            raise RuntimeError(f"reference cubemap kernel {what} failed: cudaError {rc}")

    def diffuse_cubemap_fwd(self, cubemap):
        x = cubemap.float().contiguous()
        out = torch.empty_like(x)
        torch.cuda.synchronize()
        self._chk(self.lib.ref_diffuse_cubemap_fwd(C.c_int(x.shape[1]), _p(x), _p(out)), "diffuse fwd")
        return out

    def diffuse_cubemap_bwd(self, cubemap, grad):
        x, g = cubemap.float().contiguous(), grad.float().contiguous()
        out = torch.empty_like(x)
        torch.cuda.synchronize()
        self._chk(self.lib.ref_diffuse_cubemap_bwd(C.c_int(x.shape[1]), _p(x), _p(g), _p(out)), "diffuse bwd")
        return out

    def specular_bounds(self, res, costheta_cutoff, device="cuda"):
        out = torch.empty(6, res, res, 24, dtype=torch.float32, device=device)
        torch.cuda.synchronize()
        self._chk(self.lib.ref_specular_bounds(C.c_int(res), C.c_float(costheta_cutoff), _p(out)), "bounds")
        return out

    def specular_cubemap_fwd(self, cubemap, bounds, roughness, costheta_cutoff):
        x = cubemap.float().contiguous()
        out = torch.empty(6, x.shape[1], x.shape[1], 4, dtype=torch.float32, device=x.device)
        torch.cuda.synchronize()
        self._chk(self.lib.ref_specular_cubemap_fwd(C.c_int(x.shape[1]), _p(x), _p(bounds), C.c_float(roughness),
                                                    C.c_float(costheta_cutoff), _p(out)), "specular fwd")
        return out

    def specular_cubemap_bwd(self, cubemap, bounds, grad4, roughness, costheta_cutoff):
        x, g = cubemap.float().contiguous(), grad4.float().contiguous()
        out = torch.empty_like(x)
        torch.cuda.synchronize()
        self._chk(self.lib.ref_specular_cubemap_bwd(C.c_int(x.shape[1]), _p(x), _p(bounds), _p(g), C.c_float(roughness),
                                                    C.c_float(costheta_cutoff), _p(out)), "specular bwd")
        return out

    def specular_cubemap(self, cubemap, roughness, cutoff, costheta_cutoff):
        """renderutils.specular_cubemap (ops.py:446-456) on the reference kernels -> (out, wsum, bounds)."""
        b = self.specular_bounds(cubemap.shape[1], costheta_cutoff, cubemap.device)
        o = self.specular_cubemap_fwd(cubemap, b, roughness, costheta_cutoff)
        return o[..., 0:3] / o[..., 3:], o[..., 3], b

    def specular_cubemap_grad(self, cubemap, bounds, wsum, dout, roughness, costheta_cutoff):
        """autograd of out[...,0:3] / out[...,3:] feeding specular_cubemap_bwd (which reads only channels 0..2)."""
        g4 = torch.zeros(*dout.shape[:3], 4, dtype=torch.float32, device=dout.device)
        g4[..., 0:3] = dout / wsum[..., None]
        return self.specular_cubemap_bwd(cubemap, bounds, g4, roughness, costheta_cutoff)
