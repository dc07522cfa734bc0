"""Seeded synthetic scenes and cameras of the BASELINE shapes.

Everything is generated with a CPU torch.Generator so the same seed gives the same scene on any host.
Camera matrices follow the reference's conventions exactly
(/root/reference/utils/graphics_utils.py:43-82 getWorld2View2 / getProjectionMatrix,
 /root/reference/scene/cameras.py:69-86: the matrices handed to the rasterizer are the TRANSPOSED
 world-view and full-projection matrices, i.e. column-major for the CUDA side).
"""
import math
from dataclasses import dataclass
from typing import Dict

import torch


@dataclass
class Camera:
    image_width: int
    image_height: int
    FoVx: float
    FoVy: float
    world_view_transform: torch.Tensor  # [4,4] transposed W2C
    full_proj_transform: torch.Tensor   # [4,4] transposed
    camera_center: torch.Tensor         # [3]
    znear: float = 0.01
    zfar: float = 100.0

    @property
    def tanfovx(self):
        return math.tan(self.FoVx * 0.5)

    @property
    def tanfovy(self):
        return math.tan(self.FoVy * 0.5)

    def to(self, device):
        return Camera(self.image_width, self.image_height, self.FoVx, self.FoVy,
                      self.world_view_transform.to(device), self.full_proj_transform.to(device),
                      self.camera_center.to(device), self.znear, self.zfar)


def _projection_matrix(znear, zfar, fovX, fovY):
    # utils/graphics_utils.py:62-82
    tanHalfFovY = math.tan(fovY / 2)
    tanHalfFovX = math.tan(fovX / 2)
    top = tanHalfFovY * znear
    bottom = -top
    right = tanHalfFovX * znear
    left = -right
    P = torch.zeros(4, 4)
    z_sign = 1.0
    P[0, 0] = 2.0 * znear / (right - left)
    P[1, 1] = 2.0 * znear / (top - bottom)
    P[0, 2] = (right + left) / (right - left)
    P[1, 2] = (top + bottom) / (top - bottom)
    P[3, 2] = z_sign
    P[2, 2] = z_sign * zfar / (zfar - znear)
    P[2, 3] = -(zfar * znear) / (zfar - znear)
    return P


def look_at_camera(eye, target, width, height, fovx=None, fx=None, znear=0.01, zfar=100.0) -> Camera:
    """Camera at `eye` looking at `target` (y-down, z-forward camera frame like COLMAP/3DGS)."""
    eye = torch.as_tensor(eye, dtype=torch.float64)
    target = torch.as_tensor(target, dtype=torch.float64)
    fwd = target - eye
    fwd = fwd / fwd.norm()
    up_world = torch.tensor([0.0, 0.0, 1.0], dtype=torch.float64)
    if abs(float(fwd @ up_world)) > 0.999:
        up_world = torch.tensor([0.0, 1.0, 0.0], dtype=torch.float64)
    right = torch.linalg.cross(fwd, up_world)
    right = right / right.norm()
    down = torch.linalg.cross(fwd, right)
    # world-to-camera rotation rows = camera axes (x right, y down, z forward)
    Rw2c = torch.stack([right, down, fwd], dim=0)
    t = -Rw2c @ eye
    W2C = torch.eye(4, dtype=torch.float64)
    W2C[:3, :3] = Rw2c
    W2C[:3, 3] = t
    if fovx is None:
        fovx = 2.0 * math.atan(width / (2.0 * fx))
    focal = width / (2.0 * math.tan(fovx / 2.0))
    fovy = 2.0 * math.atan(height / (2.0 * focal))
    world_view = W2C.float().transpose(0, 1).contiguous()
    proj = _projection_matrix(znear, zfar, fovx, fovy).transpose(0, 1)
    full = (world_view.unsqueeze(0).bmm(proj.unsqueeze(0))).squeeze(0).contiguous()
    center = world_view.inverse()[3, :3].contiguous()
    return Camera(width, height, fovx, fovy, world_view, full, center, znear, zfar)


def orbit_camera(k: int, K: int, width=800, height=800, radius=4.031, elevation_deg=30.0,
                 camera_angle_x=0.6911112, fx=None) -> Camera:
    az = 2.0 * math.pi * k / max(K, 1)
    el = math.radians(elevation_deg)
    eye = [radius * math.cos(el) * math.cos(az), radius * math.cos(el) * math.sin(az), radius * math.sin(el)]
    return look_at_camera(eye, [0.0, 0.0, 0.0], width, height, fovx=None if fx else camera_angle_x, fx=fx)


def _sigmoid_inv(x):
    return math.log(x / (1 - x))


def make_scene(P: int, seed: int = 0, regime: str = "trained", sh_degree: int = 3, shape: str = "lego") -> Dict:
    """Raw (pre-activation) Gaussian parameters, the layout of scene/gaussian_model.py.

    regime 'trained': log-scale ~ N(ln 0.02, 0.5^2) with one axis x0.2, random unit quaternions,
    opacity logit ~ N(1.5, 2^2), f_dc ~ N(0,.5^2), f_rest ~ N(0,.1^2), normal raw ~ N(0,I),
    material logits ~ N(0,1).   regime 'init': create_from_pcd-like (identity quats, opacity 0.1,
    normal (0,0,1), material logits 1), scales from a constant nearest-neighbour estimate.
    shape 'lego': xyz ~ U[-1.3,1.3]^3 (scene/dataset_readers.py:304-310);
    shape 'bicycle': 60% ground disk r<4, |y|<0.5; 40% shell r in [10,50]; log-scale ∝ distance.
    """
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    M = (sh_degree + 1) ** 2
    if shape == "lego":
        xyz = (torch.rand(P, 3, generator=g) * 2.6 - 1.3)
        dist_scale = torch.ones(P)
    elif shape == "bicycle":
        n_ground = int(P * 0.6)
        r = 4.0 * torch.sqrt(torch.rand(n_ground, generator=g))
        th = 2 * math.pi * torch.rand(n_ground, generator=g)
        ground = torch.stack([r * torch.cos(th), r * torch.sin(th), torch.rand(n_ground, generator=g) - 0.5], dim=1)
        n_shell = P - n_ground
        rs = 10.0 + 40.0 * torch.rand(n_shell, generator=g)
        d = torch.randn(n_shell, 3, generator=g)
        d = d / d.norm(dim=1, keepdim=True)
        shell = d * rs[:, None]
        xyz = torch.cat([ground, shell], dim=0)
        dist_scale = torch.cat([torch.ones(n_ground), rs / 4.0])
    else:
        raise ValueError(shape)

    if regime == "trained":
        log_scale = math.log(0.02) + 0.5 * torch.randn(P, 3, generator=g)
        axis = torch.randint(0, 3, (P,), generator=g)
        log_scale[torch.arange(P), axis] += math.log(0.2)
        log_scale = log_scale + torch.log(dist_scale)[:, None]
        rot = torch.randn(P, 4, generator=g)
        opacity = 1.5 + 2.0 * torch.randn(P, 1, generator=g)
        f_dc = 0.5 * torch.randn(P, 1, 3, generator=g)
        f_rest = 0.1 * torch.randn(P, M - 1, 3, generator=g)
        normal = torch.randn(P, 3, generator=g)
        albedo = torch.randn(P, 3, generator=g)
        roughness = torch.randn(P, 1, generator=g)
        metallic = torch.randn(P, 1, generator=g)
    elif regime == "init":
        # mean NN spacing of a uniform cloud: (V/P)^(1/3); create_from_pcd uses sqrt(mean 3-NN dist^2)
        spacing = (17.576 / max(P, 1)) ** (1.0 / 3.0) * 0.55
        log_scale = torch.full((P, 3), math.log(spacing)) + torch.log(dist_scale)[:, None]
        rot = torch.zeros(P, 4)
        rot[:, 0] = 1
        opacity = torch.full((P, 1), _sigmoid_inv(0.1))
        f_dc = 0.5 * torch.randn(P, 1, 3, generator=g)
        f_rest = torch.zeros(P, M - 1, 3)
        normal = torch.zeros(P, 3)
        normal[:, 2] = 1
        albedo = torch.ones(P, 3)
        roughness = torch.ones(P, 1)
        metallic = torch.ones(P, 1)
    else:
        raise ValueError(regime)
    return dict(xyz=xyz, log_scale=log_scale, rot=rot, opacity=opacity, f_dc=f_dc, f_rest=f_rest, normal=normal,
                albedo=albedo, roughness=roughness, metallic=metallic, sh_degree=sh_degree)


def activate(raw: Dict, device="cpu") -> Dict:
    """The GaussianModel getters (scene/gaussian_model.py:178-266): the rasterizer's input contract."""
    F = torch.nn.functional
    out = dict(
        means3D=raw["xyz"].to(device),
        opacity=torch.sigmoid(raw["opacity"].to(device)),
        normal=F.normalize(raw["normal"].to(device), dim=-1),
        albedo=torch.sigmoid(raw["albedo"].to(device)),
        roughness=torch.sigmoid(raw["roughness"].to(device)),
        metallic=torch.sigmoid(raw["metallic"].to(device)),
        scales=torch.exp(raw["log_scale"].to(device)),
        rotations=F.normalize(raw["rot"].to(device), dim=-1),
        shs=torch.cat([raw["f_dc"], raw["f_rest"]], dim=1).to(device).contiguous(),
        sh_degree=raw["sh_degree"],
    )
    return out


def canonical_rays(cam: Camera, device="cpu") -> torch.Tensor:
    """scene/__init__.py:157-167: ((x+0.5-W/2)/fx, (y+0.5-H/2)/fy, 1), shape [H*W,3]."""
    W, H = cam.image_width, cam.image_height
    fx = W / (2.0 * cam.tanfovx)
    fy = H / (2.0 * cam.tanfovy)
    ys, xs = torch.meshgrid(torch.arange(H, dtype=torch.float32, device=device),
                            torch.arange(W, dtype=torch.float32, device=device), indexing="ij")
    d = torch.stack([(xs + 0.5 - W / 2.0) / fx, (ys + 0.5 - H / 2.0) / fy, torch.ones_like(xs)], dim=-1)
    return d.reshape(-1, 3)


def make_light(seed: int = 0, base_res: int = 256, min_res: int = 16, device="cpu") -> Dict:
    """A CubemapLight-shaped set of textures (pbr/light.py:84-170): base U(0.25,0.75) [6,R,R,3]; the specular
    chain is the 2x2 average-pool mip chain 256->16 (the GGX prefilter of build_mips is the next §8f-1 row and is
    NOT applied here); diffuse = the 16x16 level."""
    g = torch.Generator(device="cpu")
    g.manual_seed(seed + 1000)
    base = torch.rand(6, base_res, base_res, 3, generator=g) * 0.5 + 0.25
    spec = [base]
    while spec[-1].shape[1] > min_res:
        y = spec[-1].permute(0, 3, 1, 2)
        y = torch.nn.functional.avg_pool2d(y, (2, 2))
        spec.append(y.permute(0, 2, 3, 1).contiguous())
    diffuse = spec[-1].clone()
    return dict(specular=[s.to(device).contiguous() for s in spec], diffuse=diffuse.to(device).contiguous())


_lut_cache = {}


def make_brdf_lut(res: int = 256, samples: int = 512) -> torch.Tensor:
    """A split-sum environment-BRDF LUT [1,res,res,2] (scale, bias of F0) by GGX importance sampling with a
    Hammersley sequence (Karis 2013, height-correlated Smith visibility). x = NoV, y = roughness, texel centres at (i+0.5)/res. Synthetic stand-in
    for the reference's data file pbr/brdf_256_256.bin, which is not copied into this repo."""
    key = (res, samples)
    if key in _lut_cache:
        return _lut_cache[key]
    nov = ((torch.arange(res, dtype=torch.float64) + 0.5) / res)[None, :, None]      # [1,res,1]
    rough = ((torch.arange(res, dtype=torch.float64) + 0.5) / res)[:, None, None]    # [res,1,1]
    i = torch.arange(samples, dtype=torch.int64)
    bits = i.clone()
    bits = ((bits << 16) | (bits >> 16)) & 0xFFFFFFFF
    bits = (((bits & 0x55555555) << 1) | ((bits & 0xAAAAAAAA) >> 1)) & 0xFFFFFFFF
    bits = (((bits & 0x33333333) << 2) | ((bits & 0xCCCCCCCC) >> 2)) & 0xFFFFFFFF
    bits = (((bits & 0x0F0F0F0F) << 4) | ((bits & 0xF0F0F0F0) >> 4)) & 0xFFFFFFFF
    bits = (((bits & 0x00FF00FF) << 8) | ((bits & 0xFF00FF00) >> 8)) & 0xFFFFFFFF
    xi1 = ((i.double() + 0.5) / samples)[None, None, :]
    xi2 = (bits.double() * 2.3283064365386963e-10)[None, None, :]
    a = rough * rough
    phi = 2.0 * math.pi * xi1
    cos_t = torch.sqrt((1.0 - xi2) / (1.0 + (a * a - 1.0) * xi2))
    sin_t = torch.sqrt(torch.clamp(1.0 - cos_t * cos_t, min=0.0))
    hx, hz = sin_t * torch.cos(phi), cos_t
    vx, vz = torch.sqrt(1.0 - nov * nov), nov
    vdh = vx * hx + vz * hz
    lz = 2.0 * vdh * hz - vz
    nol, noh, vdh_c = lz.clamp(min=0.0), hz.clamp(min=0.0), vdh.clamp(min=0.0)
    # height-correlated Smith-GGX visibility (the variant that reproduces the reference's data file to 1e-3)
    lam_v = nol * torch.sqrt(nov * nov * (1.0 - a * a) + a * a)
    lam_l = nov * torch.sqrt(nol * nol * (1.0 - a * a) + a * a)
    g = (0.5 / (lam_v + lam_l + 1e-12)) * 4.0 * nol * nov
    g_vis = g * vdh_c / (noh * nov + 1e-12)
    fc = (1.0 - vdh_c) ** 5
    ok = (nol > 0).double()
    A = ((1.0 - fc) * g_vis * ok).mean(-1)
    B = (fc * g_vis * ok).mean(-1)
    lut = torch.stack([A, B], -1).float()[None].contiguous()
    _lut_cache[key] = lut
    return lut
