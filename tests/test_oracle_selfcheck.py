"""CPU: internal consistency of the oracle and of the host-side logic (no reference data needed)."""
import math

import numpy as np
import pytest
import torch

import gigs_oracle as O
from gigs import scene, shade


def test_hemisphere_loop_counts_match_the_float_accumulated_reference_loops():
    # 32 phi x 16 theta = 512 directions, normaliser sum(cos*sin) ~= 162.4508
    phis, thetas = O.hemisphere_dirs(0.0625)
    assert (len(phis), len(thetas)) == (32, 16)
    assert abs(thetas[-1] - 1.4726217) < 1e-6
    s = 0.0
    for _ in phis:
        for t in thetas:
            s += math.cos(t) * math.sin(t)
    assert abs(s - 162.4508) < 1e-2


def test_higher_msb_key_bits():
    assert O.higher_msb(2500) == 12      # 800x800 -> 44-bit keys
    assert O.higher_msb(4056) == 12      # 1237x822
    assert O.higher_msb(32400) == 15     # 4K -> 47 bits
    assert O.higher_msb(1) == 1


def test_blend_backward_agrees_with_autograd_of_a_differentiable_blend():
    """Transcribed backward vs torch.autograd through a straightforward differentiable re-expression of the
    front-to-back blend (<=64 Gaussians, 32x32 px), for colour, opacity-map and material channels."""
    torch.manual_seed(0)
    raw = scene.make_scene(48, seed=2, regime="trained")
    g = scene.activate(raw)
    g["means3D"] = g["means3D"] * 0.25
    cam = scene.orbit_camera(0, 8, 32, 32)
    bg = torch.tensor([0.3, 0.2, 0.1])
    f = O.rasterize_forward(g, cam, bg)
    pre, binn = f["pre"], f["binning"]
    W = H = 32
    gen = torch.Generator().manual_seed(5)
    grads = {k: torch.randn(c, H, W, generator=gen) for k, c in
             (("color", 3), ("opacity", 1), ("albedo", 3), ("roughness", 1), ("metallic", 1), ("normal", 3),
              ("depth", 1))}
    grads["depth"].zero_()   # the reference's depth backward is knowingly inconsistent with its forward (A.4)
    grads["normal"][:, 0, :] = 0; grads["normal"][:, -1, :] = 0; grads["normal"][:, :, 0] = 0; grads["normal"][:, :, -1] = 0
    acc = O.blend_backward(pre, binn, g, cam, bg, f, grads)

    # differentiable re-expression w.r.t. the 2D quantities
    xy = pre["means2D"].clone().requires_grad_(True)
    conic = pre["conic"].clone().requires_grad_(True)
    op = g["opacity"].reshape(-1).clone().requires_grad_(True)
    rgb = pre["rgb"].clone().requires_grad_(True)
    alb = g["albedo"].clone().requires_grad_(True)
    ys, xs = torch.meshgrid(torch.arange(H, dtype=torch.float32), torch.arange(W, dtype=torch.float32), indexing="ij")
    total = torch.zeros(())
    ranges, plist = binn["ranges"], binn["point_list"]
    ncon = f["n_contrib"].reshape(H, W)
    for ty in range(2):
        for tx in range(2):
            lo, hi = ranges[ty * 2 + tx]
            px, py = xs[ty * 16:(ty + 1) * 16, tx * 16:(tx + 1) * 16], ys[ty * 16:(ty + 1) * 16, tx * 16:(tx + 1) * 16]
            T = torch.ones_like(px)
            C = torch.zeros(3, 16, 16); A = torch.zeros(3, 16, 16); Oo = torch.zeros(16, 16)
            nc = ncon[ty * 16:(ty + 1) * 16, tx * 16:(tx + 1) * 16]
            for j in range(int(hi - lo)):
                gid = int(plist[lo + j])
                dx, dy = xy[gid, 0] - px, xy[gid, 1] - py
                power = -0.5 * (conic[gid, 0] * dx * dx + conic[gid, 2] * dy * dy) - conic[gid, 1] * dx * dy
                a_raw = op[gid] * torch.exp(power)
                # the reference back-propagates THROUGH min(0.99, .) (backward.cu:545,609,627): straight-through clamp
                alpha = a_raw + (torch.clamp(a_raw, max=0.99) - a_raw).detach()
                use = (power <= 0) & (alpha >= 1.0 / 255.0) & (j < nc)
                w = torch.where(use, alpha * T, torch.zeros(()))
                C = C + rgb[gid][:, None, None] * w
                # material / normal / depth channels do NOT feed dL/dalpha in the reference (backward.cu:580-590)
                A = A + alb[gid][:, None, None] * w.detach()
                Oo = Oo + w
                T = torch.where(use, T * (1 - alpha), T)
            C = C + T * bg[:, None, None]
            sl = (slice(None), slice(ty * 16, (ty + 1) * 16), slice(tx * 16, (tx + 1) * 16))
            total = total + (C * grads["color"][sl]).sum() + (A * grads["albedo"][sl]).sum() \
                + (Oo * grads["opacity"][0][sl[1:]]).sum()
    total.backward()
    def rel(a, b):
        return float((a - b).norm() / (b.norm() + 1e-20))
    assert rel(acc["colors"], rgb.grad) < 1e-4
    assert rel(acc["albedo"], alb.grad) < 1e-4
    assert rel(acc["opacity"], op.grad) < 1e-3
    # the reference stores the off-diagonal gradient per matrix entry (it appears twice in the symmetric 2x2):
    # half of d/d(conic.y) of the scalar parametrisation (backward.cu:622-624, consumed at :212-214)
    cg = acc["conic"][:, [0, 1, 3]] * torch.tensor([1.0, 2.0, 1.0])
    assert rel(cg, conic.grad) < 1e-3
    # dL/dmean2D is scaled by 0.5*W / 0.5*H (ndc units) in the reference
    assert rel(acc["mean2D"][:, 0] / (0.5 * W), xy.grad[:, 0]) < 1e-3
    assert rel(acc["mean2D"][:, 1] / (0.5 * H), xy.grad[:, 1]) < 1e-3


def test_per_gaussian_backward_agrees_with_autograd_of_the_forward_transcription():
    raw = scene.make_scene(64, seed=4, regime="trained")
    g = scene.activate(raw)
    cam = scene.orbit_camera(1, 8, 64, 64)
    pre = O.preprocess(g, cam)
    P = 64
    gen = torch.Generator().manual_seed(9)
    acc = dict(conic=torch.randn(P, 4, generator=gen), mean2D=torch.randn(P, 3, generator=gen),
               colors=torch.randn(P, 3, generator=gen), depth=torch.zeros(P), opacity=torch.zeros(P),
               normal=torch.zeros(P, 3), albedo=torch.zeros(P, 3), roughness=torch.zeros(P), metallic=torch.zeros(P))
    out = O.gaussian_backward(pre, g, cam, acc)
    means = g["means3D"].clone().requires_grad_(True)
    scales = g["scales"].clone().requires_grad_(True)
    rots = g["rotations"].clone().requires_grad_(True)
    shs = g["shs"].clone().requires_grad_(True)
    g2 = dict(g, means3D=means, scales=scales, rotations=rots, shs=shs)
    W = H = 64
    V, PM = cam.world_view_transform, cam.full_proj_transform
    fx, fy = W / (2 * cam.tanfovx), H / (2 * cam.tanfovy)
    cov3D = O.compute_cov3d(scales, 1.0, rots)
    cov = O.compute_cov2d(means, fx, fy, cam.tanfovx, cam.tanfovy, cov3D, V)
    det = cov[:, 0] * cov[:, 2] - cov[:, 1] ** 2
    conic = torch.stack([cov[:, 2] / det, -cov[:, 1] / det, cov[:, 0] / det], 1)
    ph = O.transform_point_4x4(means, PM)
    pw = 1.0 / (ph[:, 3] + 1e-7)
    ndc = ph[:, :2] * pw[:, None]          # dL/dmean2D is w.r.t. ndc-scaled pixel coords (see blend test)
    rgb, _ = O.compute_color_from_sh(3, means, cam.camera_center, shs)
    vis = (pre["radii"] > 0).float()
    sym = torch.tensor([1.0, 2.0, 1.0])  # off-diagonal gradient is stored per matrix entry (appears twice)
    L = ((conic * acc["conic"][:, [0, 1, 3]] * sym).sum(1) * vis).sum() + ((ndc * acc["mean2D"][:, :2]).sum(1) * vis).sum() \
        + ((rgb * acc["colors"] * (~pre["clamped"]).float()).sum(1) * vis).sum()
    L.backward()
    def rel(a, b):
        return float((a - b).norm() / (b.norm() + 1e-20))
    assert rel(out["means3D"], means.grad) < 2e-3
    assert rel(out["scales"], scales.grad) < 2e-3
    assert rel(out["rotations"], rots.grad) < 2e-3
    assert rel(out["sh"], shs.grad) < 1e-4


def test_median3x3_semantics():
    x = torch.arange(25, dtype=torch.float32).reshape(1, 5, 5)
    m = O.median3x3(x)
    assert m[0, 2, 2] == 12           # interior: plain median
    assert m[0, 0, 0] == 0            # corner: 5 zero-pad values dominate -> lower median is 0
    x[0, 2, 2] = float("nan")
    m = O.median3x3(x)
    assert torch.isnan(m[0, 1:4, 1:4]).all() and not torch.isnan(m[0, 0, 0])   # NaN spreads to the 3x3 neighbourhood


def test_bilateral_preserves_constants_and_edges():
    x = torch.full((3, 8, 8), 0.37)
    assert torch.allclose(O.bilateral3x3(x), x, atol=1e-6)
    e = torch.zeros(3, 8, 8)
    e[:, :, 4:] = 10.0                 # colour distance 30 across the edge -> weight exp(-450) = 0
    assert torch.allclose(O.bilateral3x3(e), e, atol=1e-6)


def test_cube_sampling_is_seamless_and_exact_on_constants():
    torch.manual_seed(0)
    tex = torch.rand(6, 8, 8, 3)
    const = torch.full((6, 8, 8, 3), 0.5)
    d = torch.nn.functional.normalize(torch.randn(4000, 3), dim=-1)
    assert torch.allclose(O.tex_cube(const, d), torch.full((4000, 3), 0.5), atol=1e-6)
    # continuity across every edge: directions epsilon apart on both sides of a face boundary agree
    t = torch.linspace(-0.95, 0.95, 41)
    for axis_a, axis_b in ((0, 1), (0, 2), (1, 2)):
        for sa in (-1.0, 1.0):
            for sb in (-1.0, 1.0):
                third = 3 - axis_a - axis_b
                base = torch.zeros(41, 3)
                base[:, third] = t
                d1, d2 = base.clone(), base.clone()
                d1[:, axis_a] = sa; d1[:, axis_b] = sb * (1 - 1e-4)
                d2[:, axis_a] = sa * (1 - 1e-4); d2[:, axis_b] = sb
                assert (O.tex_cube(tex, d1) - O.tex_cube(tex, d2)).abs().max() < 2e-3
    # texel centres reproduce texel values
    w = 8
    for f in range(6):
        iu, iv = 3, 5
        s, tt = (2 * iu + 1 - w) / w, (2 * iv + 1 - w) / w
        dirs = {0: (1, -tt, -s), 1: (-1, -tt, s), 2: (s, 1, tt), 3: (s, -1, -tt), 4: (s, -tt, 1), 5: (-s, -tt, -1)}[f]
        v = O.tex_cube(tex, torch.tensor([dirs], dtype=torch.float32))
        assert torch.allclose(v[0], tex[f, iv, iu], atol=1e-5)


def test_lut_lookup_clamps_to_edge_texels():
    lut = torch.rand(16, 16, 2)
    uv = torch.tensor([[-1.0, -1.0], [2.0, 2.0], [0.5 / 16, 0.5 / 16]])
    v = O.tex_2d_clamp(lut, uv)
    assert torch.allclose(v[0], lut[0, 0]) and torch.allclose(v[1], lut[15, 15]) and torch.allclose(v[2], lut[0, 0])


def test_generated_brdf_lut_matches_the_reference_data_file_texels():
    # known texels of /root/reference/pbr/brdf_256_256.bin; our LUT is generated, not copied
    lut = shade.make_brdf_lut()[0]
    for (y, x), (a, b) in {(0, 0): (0.00973, 0.99025), (128, 128): (0.83426, 0.02193), (255, 255): (0.30928, 3.5e-05)}.items():
        assert abs(float(lut[y, x, 0]) - a) < 5e-3 and abs(float(lut[y, x, 1]) - b) < 5e-3


def test_dist2_brute_force_known_answer():
    pts = torch.tensor([[0.0, 0, 0], [1, 0, 0], [0, 2, 0], [0, 0, 3], [5, 5, 5]])
    d = O.dist2(pts)
    assert abs(float(d[0]) - (1 + 4 + 9) / 3) < 1e-6
    dup = torch.zeros(4, 3)
    assert float(O.dist2(dup).max()) == 0.0       # duplicates give 0 (other points, not self)


def test_camera_convention_matches_reference_helpers():
    cam = scene.orbit_camera(0, 8, 800, 800)
    fx = 800 / (2 * cam.tanfovx)
    assert abs(fx - 1111.1) < 0.2                  # NeRF-synthetic intrinsics (camera_angle_x = 0.6911112)
    # the origin projects to the image centre, 4.031 in front of the camera
    p = O.transform_point_4x3(torch.zeros(1, 3), cam.world_view_transform)
    assert abs(float(p[0, 2]) - 4.031) < 1e-4 and abs(float(p[0, 0])) < 1e-5 and abs(float(p[0, 1])) < 1e-5
    ph = O.transform_point_4x4(torch.zeros(1, 3), cam.full_proj_transform)
    assert abs(float(ph[0, 3]) - 4.031) < 1e-4
    assert torch.allclose(cam.camera_center.norm(), torch.tensor(4.031), atol=1e-4)


# ---- cubemap prefilter restatement (oracle section "Cubemap prefilter") ---------------------------------------------
def test_cubemap_texel_geometry():
    d = O.cm_cube_to_dir(16)
    assert torch.allclose(d.norm(dim=-1), torch.ones(6, 16, 16), atol=1e-6)
    # face s points along its major axis (c_src/cubemap.cu:36-45)
    for s, (ax, sg) in enumerate([(0, 1), (0, -1), (1, 1), (1, -1), (2, 1), (2, -1)]):
        assert (d[s, ..., ax] * sg > 0.5).all()
    # pixel_area is d(atan x) * d(atan y), the reference's (separable) approximation of the texel's solid angle: over
    # a face it sums to ~(pi/2)^2, not to 4*pi/6 (and |x - H| shifts the negative half by one texel)
    assert abs(O.cm_pixel_area(256).sum().item() - (math.pi / 2) ** 2) < 0.02


def test_cubemap_filters_preserve_constants_and_are_adjoint():
    c = torch.full((6, 16, 16, 3), 0.7)
    o, w = O.specular_cubemap(c, 1.0)
    assert (o - 0.7).abs().max().item() < 3e-6 and (w > 0).all()
    x = torch.randn(6, 16, 16, 3, generator=torch.Generator().manual_seed(1))
    g = torch.randn(6, 16, 16, 3, generator=torch.Generator().manual_seed(2))
    for fwd, bwd in ((lambda t: O.specular_cubemap(t, 0.5)[0], lambda t: O.specular_cubemap_backward(t, 0.5)),
                     (O.diffuse_cubemap, O.diffuse_cubemap_backward)):
        lhs, rhs = (fwd(x) * g).sum().item(), (bwd(g) * x).sum().item()
        assert abs(lhs - rhs) < 1e-3 * max(1.0, abs(lhs))
    # the cone of the bounds contains the texel itself and, at roughness 1, (almost) the whole hemisphere
    b = O.specular_bounds(16, O.ndf_cutoff(1.0))
    assert (b[0, 8, 8, 0] == torch.tensor([0, 15, 0, 15])).all()        # own face: everything
    assert (b[0, 8, 8, 1, 0] > b[0, 8, 8, 1, 1])                        # opposite face: empty box


def test_cubemap_mip_chain_and_its_reference_backward():
    x = torch.rand(6, 64, 64, 3, generator=torch.Generator().manual_seed(3))
    y = O.cubemap_mip(x)
    yt = torch.nn.functional.avg_pool2d(x.permute(0, 3, 1, 2), (2, 2)).permute(0, 2, 3, 1)
    assert torch.allclose(y, yt, atol=1e-7)
    # backward = bilinear lookup of 0.25 * dout: a constant upstream gradient c gives 0.25 c everywhere
    g = O.cubemap_mip_backward(torch.full((6, 32, 32, 3), 2.0))
    assert (g - 0.5).abs().max().item() < 1e-6 and g.shape == (6, 64, 64, 3)
    with pytest.raises(ZeroDivisionError):
        O.build_mips(torch.rand(6, 32, 32, 3))      # two levels: the reference's schedule divides by zero too
