"""GPU (B200): the exported variants of the rasterizer against the REFERENCE's own kernels (oracle/_ref), through
the C-ABI: argmax_depth, scale_modifier != 1, prefiltered, derive_normal=False, the radiance-only rasterizer
(lite_rasterize_gaussians: rasterizer_impl.cu:338-482, forward.cu:279-418) and BASELINE configs[1] at full size
(300k Gaussians, 800x800) forward + backward.

Gates (north_star): integers (radii, n_contrib, sorted order, ranges) bit-exact; maps <= 1e-4 max-abs; gradients
<= 1e-3 relative."""
import pytest
import torch

pytestmark = pytest.mark.gpu

import gpu_util as U
import refshim
from gigs import scene

DEV = "cuda:0"
MAPS = ("color", "opacity", "depth", "normal", "normal_view", "pos", "albedo", "roughness", "metallic")
GRADS = ("means2D", "colors", "opacity", "normal", "albedo", "roughness", "metallic", "means3D", "cov3D", "sh",
         "scales", "rotations")
needs_ref = pytest.mark.skipif(not refshim.available(), reason="oracle/_ref not built (needs /root/reference here)")


def make(P, W, H, seed=0, regime="trained", k=1):
    raw = scene.make_scene(P, seed=seed, regime=regime)
    g = scene.activate(raw, DEV)
    cam = scene.orbit_camera(k, 8, W, H).to(DEV)
    bg = torch.tensor([0.1, 0.2, 0.3], device=DEV)
    return g, cam, bg


def dense_grads(W, H, seed=1):
    gen = torch.Generator().manual_seed(seed)
    N = W * H
    return {k: (torch.randn(c, H, W, generator=gen) / N).to(DEV) for k, c in
            (("depth", 1), ("color", 3), ("opacity", 1), ("normal", 3), ("albedo", 3), ("roughness", 1),
             ("metallic", 1))}


def same_binning(fo, ro, rs, P, W, H):
    st = U.decode_state(fo, P, W, H)
    assert fo["num_rendered"] == ro["num_rendered"]
    assert torch.equal(fo["radii"], ro["radii"])
    for k in ("tiles_touched", "keys_sorted", "point_list", "ranges", "n_contrib"):
        assert torch.equal(st[k], rs[k]), f"{k} not bit-exact"


@needs_ref
@pytest.mark.parametrize("inference", [False, True])
def test_argmax_depth_vs_reference(inference):
    """argmax_depth=True: the depth map is the depth of the largest blend weight (forward.cu:575-592)."""
    P, W, H = 20000, 400, 300
    g, cam, bg = make(P, W, H, seed=3)
    ref = refshim.RefRasterizer()
    ro = ref.forward(g, cam, bg, argmax_depth=True, inference=inference)
    rs = ref.state()
    fo = U.ours_forward(g, cam, bg, argmax=True, inference=inference)
    same_binning(fo, ro, rs, P, W, H)
    for k in MAPS:
        U.assert_close_map(fo[k], ro[k], 1e-4, k)
    assert torch.equal(fo["depth"], ro["depth"]), "arg-max depth is a selection: it must be bit-exact"
    # and it differs from the expected-depth map (the flag did something)
    fe = U.ours_forward(g, cam, bg, argmax=False, inference=inference)
    assert float((fe["depth"] - fo["depth"]).abs().max()) > 1e-3
    ref.close()


@needs_ref
@pytest.mark.parametrize("modifier", [0.7, 1.6])
def test_scale_modifier_forward_and_backward_vs_reference(modifier):
    P, W, H = 12000, 320, 240
    g, cam, bg = make(P, W, H, seed=5)
    ref = refshim.RefRasterizer()
    ro = ref.forward(g, cam, bg, scale_modifier=modifier)
    rs = ref.state()
    fo = U.ours_forward(g, cam, bg, scale_modifier=modifier)
    same_binning(fo, ro, rs, P, W, H)
    for k in MAPS:
        U.assert_close_map(fo[k], ro[k], 1e-4, k)
    grads = dense_grads(W, H)
    rb = ref.backward(g, cam, bg, ro["radii"], grads)
    ob = U.ours_backward(g, cam, bg, fo, grads, scale_modifier=modifier)
    for k in GRADS:
        U.assert_grad_close(ob[k], rb[k].reshape(ob[k].shape), k)
    ref.close()


@needs_ref
def test_prefiltered_vs_reference():
    """prefiltered=True promises that no Gaussian is behind the near plane (auxiliary.h:166-172 traps otherwise):
    a scene entirely in front of the camera renders exactly as without the flag."""
    P, W, H = 8000, 256, 192
    g, cam, bg = make(P, W, H, seed=9)
    V = cam.world_view_transform
    z = (g["means3D"] @ V[:3, 2]) + V[3, 2]
    assert float(z.min()) > 0.25, "scene must lie in front of the near plane for this test"
    ref = refshim.RefRasterizer()
    ro = ref.forward(g, cam, bg, prefiltered=True)
    rs = ref.state()
    fo = U.ours_forward(g, cam, bg, prefiltered=True)
    same_binning(fo, ro, rs, P, W, H)
    for k in MAPS:
        U.assert_close_map(fo[k], ro[k], 1e-4, k)
    f0 = U.ours_forward(g, cam, bg, prefiltered=False)
    for k in MAPS:
        assert torch.equal(torch.nan_to_num(f0[k]), torch.nan_to_num(fo[k])), k
    ref.close()


@needs_ref
@pytest.mark.parametrize("argmax,modifier,precomp", [(False, 1.0, False), (True, 1.0, False), (False, 0.8, True)])
def test_lite_rasterizer_vs_reference(argmax, modifier, precomp):
    """lite_rasterize_gaussians: radiance, opacity, depth and radii only."""
    import diff_gaussian_rasterization as dgr
    P, W, H = 15000, 333, 257
    g, cam, bg = make(P, W, H, seed=8)
    E = torch.Tensor([])
    colors = None
    if precomp:
        colors = torch.rand(P, 3, generator=torch.Generator().manual_seed(4)).to(DEV)
    ref = refshim.RefRasterizer()
    ro = ref.lite_forward(g, cam, bg, scale_modifier=modifier, argmax_depth=argmax, colors_precomp=colors)
    R, color, opacity, radii, depth = dgr._C.lite_rasterize_gaussians(
        bg, g["means3D"], colors if precomp else E, g["opacity"], g["scales"], g["rotations"], E,
        E if precomp else g["shs"], cam.camera_center, cam.world_view_transform, cam.full_proj_transform, modifier,
        cam.tanfovx, cam.tanfovy, H, W, 3, False, argmax)
    assert R == ro["num_rendered"] > 0
    assert torch.equal(radii, ro["radii"])
    U.assert_close_map(color, ro["color"], 1e-4, "lite color")
    U.assert_close_map(opacity, ro["opacity"], 1e-4, "lite opacity")
    U.assert_close_map(depth, ro["depth"], 1e-4, "lite depth")
    if argmax:
        assert torch.equal(depth, ro["depth"])
    # the full rasterizer's radiance / opacity / depth are the lite rasterizer's (same blend, fewer channels)
    fo = U.ours_forward(g, cam, bg, argmax=argmax, scale_modifier=modifier, colors_precomp=colors)
    assert torch.equal(fo["color"], color) and torch.equal(fo["opacity"], opacity) and torch.equal(fo["depth"], depth)
    ref.close()


@needs_ref
def test_derive_normal_false_module_outputs_vs_reference():
    """GaussianRasterizer(..., derive_normal=False): pseudo normals and positions are zero maps (DGR/__init__.py:
    486-488, both filters map zeros to zeros) and SSAO marches from position 0 with the blended view normal."""
    import diff_gaussian_rasterization as dgr
    P, W, H = 6000, 192, 160
    g, cam, bg = make(P, W, H, seed=12)
    settings = dgr.GaussianRasterizationSettings(
        image_height=H, image_width=W, tanfovx=cam.tanfovx, tanfovy=cam.tanfovy, bg=bg, scale_modifier=1.0,
        viewmatrix=cam.world_view_transform, projmatrix=cam.full_proj_transform, sh_degree=3,
        campos=cam.camera_center, prefiltered=False, debug=False, argmax_depth=False, inference=False,
        radius=0.8, bias=0.01, thick=0.05, delta=0.0625, step=16, start=8)
    rast = dgr.GaussianRasterizer(settings)
    m2d = torch.zeros_like(g["means3D"])
    out = rast(g["means3D"], m2d, g["opacity"], g["normal"], g["albedo"], g["roughness"], g["metallic"], shs=g["shs"],
               scales=g["scales"], rotations=g["rotations"], derive_normal=False)
    (color, radii, opacity_map, depth, normal_from_depth, out_normal, occlusion, albedo_map, roughness_map,
     metallic_map, out_normal_view, depth_pos_filter) = out
    assert float(normal_from_depth.abs().max()) == 0.0 and float(depth_pos_filter.abs().max()) == 0.0
    ref = refshim.RefRasterizer()
    ro = ref.forward(g, cam, bg)
    U.assert_close_map(color, ro["color"], 1e-4, "color")
    U.assert_close_map(out_normal_view, ro["normal_view"], 1e-4, "normal_view")
    fx, fy = W / (2 * cam.tanfovx), H / (2 * cam.tanfovy)
    occ_r = refshim.ssao(W, H, fx, fy, 0.8, 0.01, 0.05, 0.0625, 16, 8, ro["normal_view"], torch.zeros_like(ro["pos"]))
    assert torch.equal(torch.nan_to_num(occlusion), torch.nan_to_num(occ_r)), "SSAO from the zero position map"
    # with derive_normal=True the same call produces non-trivial pseudo normals
    out2 = rast(g["means3D"], m2d, g["opacity"], g["normal"], g["albedo"], g["roughness"], g["metallic"], shs=g["shs"],
                scales=g["scales"], rotations=g["rotations"], derive_normal=True)
    assert float(out2[4].abs().max()) > 0.1
    ref.close()


@needs_ref
def test_c2_full_size_forward_backward_and_gi_vs_reference():
    """BASELINE configs[1] at its full size: 300k Gaussians, 800x800, degree 3; binning bit-exact, maps <= 1e-4,
    all twelve gradients <= 1e-3 relative, SSAO / SSR (march running: start 8) bit-identical."""
    import diff_gaussian_rasterization as dgr
    P, W, H = 300000, 800, 800
    g, cam, bg = make(P, W, H, seed=0, k=0)
    ref = refshim.RefRasterizer()
    ro = ref.forward(g, cam, bg)
    rs = ref.state()
    fo = U.ours_forward(g, cam, bg)
    same_binning(fo, ro, rs, P, W, H)
    for k in MAPS:
        U.assert_close_map(fo[k], ro[k], 1e-4, k)
    grads = dense_grads(W, H)
    rb = ref.backward(g, cam, bg, ro["radii"], grads)
    ob = U.ours_backward(g, cam, bg, fo, grads)
    for k in GRADS:
        U.assert_grad_close(ob[k], rb[k].reshape(ob[k].shape), k)
    # screen-space GI on this G-buffer, march running
    fx, fy = W / (2 * cam.tanfovx), H / (2 * cam.tanfovy)
    V = cam.world_view_transform
    n_r, p_r = refshim.depth_to_normal(W, H, fx, fy, V, ro["depth"])
    n_o, p_o = dgr._C.depth_to_normal(W, H, fx, fy, V, fo["depth"])
    assert torch.equal(n_o, n_r) and torch.equal(p_o, p_r)
    gi = (0.8, 0.01, 0.05, 0.0625, 16, 8)
    occ_r = refshim.ssao(W, H, fx, fy, *gi, ro["normal_view"], p_r)
    occ_o = dgr._C.SSAO(W, H, fx, fy, *gi, fo["normal_view"], p_o)
    assert torch.equal(occ_o, occ_r), "SSAO not bit-identical at 800x800"
    rgb = torch.rand(3, H, W, device=DEV, generator=torch.Generator(device=DEV).manual_seed(2))
    F0 = (1.0 - ro["metallic"]) * 0.04 + ro["albedo"] * ro["metallic"]
    c_r, a_r = refshim.ssr(W, H, fx, fy, *gi, ro["normal_view"], p_r, rgb, ro["albedo"], ro["roughness"],
                           ro["metallic"], F0)
    c_o, a_o = dgr._C.SSR(W, H, fx, fy, *gi, fo["normal_view"], p_o, rgb, fo["albedo"], fo["roughness"],
                          fo["metallic"], F0)
    assert torch.equal(c_o, c_r) and torch.equal(a_o, a_r), "SSR not bit-identical at 800x800"
    ref.close()


@needs_ref
@pytest.mark.parametrize("pairs,block_test", [(0, 0), (1, 0), (2, 1)])
def test_gi_march_tuning_variants_are_bit_identical(pairs, block_test):
    """Every setting of gigs_gi_tune gives the reference's bits (default (1, 1) is covered by the other tests)."""
    import diff_gaussian_rasterization as dgr
    from gigs import _lib
    L = _lib.load()
    P, W, H = 30000, 400, 304
    g, cam, bg = make(P, W, H, seed=6)
    fo = U.ours_forward(g, cam, bg)
    fx, fy = W / (2 * cam.tanfovx), H / (2 * cam.tanfovy)
    n_o, p_o = dgr._C.depth_to_normal(W, H, fx, fy, cam.world_view_transform, fo["depth"])
    gi = (0.8, 0.01, 0.05, 0.0625, 16, 8)
    occ_r = refshim.ssao(W, H, fx, fy, *gi, fo["normal_view"], p_o)
    rgb = torch.rand(3, H, W, device=DEV, generator=torch.Generator(device=DEV).manual_seed(2))
    F0 = (1.0 - fo["metallic"]) * 0.04 + fo["albedo"] * fo["metallic"]
    c_r, a_r = refshim.ssr(W, H, fx, fy, *gi, fo["normal_view"], p_o, rgb, fo["albedo"], fo["roughness"],
                           fo["metallic"], F0)
    try:
        _lib.check(L.gigs_gi_tune(pairs, block_test), "gigs_gi_tune")
        occ_o = dgr._C.SSAO(W, H, fx, fy, *gi, fo["normal_view"], p_o)
        c_o, a_o = dgr._C.SSR(W, H, fx, fy, *gi, fo["normal_view"], p_o, rgb, fo["albedo"], fo["roughness"],
                              fo["metallic"], F0)
    finally:
        _lib.check(L.gigs_gi_tune(1, 1), "gigs_gi_tune")
    assert torch.equal(occ_o, occ_r) and torch.equal(c_o, c_r) and torch.equal(a_o, a_r)
    # odd trip counts (step - start not a multiple of the probes per inner step) and a step that is not a power of two
    for step, start in ((16, 9), (16, 13), (12, 5)):
        gi2 = (0.8, 0.01, 0.05, 0.0625, step, start)
        assert torch.equal(dgr._C.SSAO(W, H, fx, fy, *gi2, fo["normal_view"], p_o),
                           refshim.ssao(W, H, fx, fy, *gi2, fo["normal_view"], p_o)), (step, start)


@needs_ref
@pytest.mark.parametrize("W,H", [(333, 257), (1920, 1080), (64, 48)])
def test_gi_march_special_values_and_sizes_vs_reference(W, H):
    """The march on synthetic G-buffers that exercise every exit of the fast path: NaN / zero / infinite normals,
    normals parallel to the up vector (NaN tangent frame), positions that are NaN, infinite, huge (beyond the fast
    path's range), exactly zero, behind the camera, and depths for which z + 1e-7 is exactly 0 or denormal (the shared
    reciprocal's range test) — at sizes that are not multiples of the tile, and at one large enough for 32-pixel
    blocks in the block test. Everything must equal the reference's kernels bit for bit."""
    import diff_gaussian_rasterization as dgr
    g = torch.Generator().manual_seed(W * 7 + H)
    fx, fy = 0.9 * W, 0.9 * W
    ys, xs = torch.meshgrid(torch.arange(H, dtype=torch.float32), torch.arange(W, dtype=torch.float32), indexing="ij")
    z = 3.0 + 0.8 * torch.sin(xs / 37.0) * torch.cos(ys / 23.0) + 0.05 * torch.rand(H, W, generator=g)
    pos = torch.stack([(xs - W / 2) / fx * z, (ys - H / 2) / fy * z, z])
    nrm = torch.nn.functional.normalize(torch.randn(3, H, W, generator=g) * 0.3 + torch.tensor([0.0, 0.0, -1.0])[:, None, None], dim=0)
    n = H * W
    idx = torch.randperm(n, generator=g)
    flat_n, flat_p = nrm.reshape(3, n).clone(), pos.reshape(3, n).clone()
    k = max(4, n // 400)

    def take(i):
        return idx[i * k:(i + 1) * k]
    flat_n[:, take(0)] = float("nan")
    flat_n[:, take(1)] = 0.0
    flat_n[0, take(2)] = float("inf")
    flat_n[:, take(3)] = torch.tensor([0.0, 1.0, 0.0])[:, None]          # parallel to `up`: NaN tangent
    flat_n[:, take(4)] = torch.tensor([0.0, -2.5, 0.0])[:, None]
    flat_p[2, take(5)] = float("nan")
    flat_p[0, take(6)] = float("nan")
    flat_p[1, take(7)] = float("inf")
    flat_p[:, take(8)] = 0.0
    flat_p[:, take(9)] = torch.tensor([1.0e8, -3.0e7, 5.0e8])[:, None]  # beyond the fast path's bounds
    flat_p[2, take(10)] = -2.0                                           # behind the camera
    flat_p[2, take(11)] = -0.0000001                                     # z + 1e-7 == 0 at the pixel itself
    flat_p[2, take(12)] = 1.0e-30
    flat_p[:, take(13)] = torch.tensor([3.0e5, 2.0e5, 9.0e5])[:, None]  # large but inside the fast path's bounds
    nrm, pos = flat_n.reshape(3, H, W).to(DEV).contiguous(), flat_p.reshape(3, H, W).to(DEV).contiguous()
    rgb = torch.rand(3, H, W, generator=g).to(DEV)
    rgb[:, 5, 7] = float("nan")                                           # a radiance texel that poisons what hits it
    alb = torch.rand(3, H, W, generator=g).to(DEV)
    rough = torch.rand(1, H, W, generator=g).to(DEV)
    met = torch.rand(1, H, W, generator=g).to(DEV)
    F0 = (1.0 - met) * 0.04 + alb * met
    for gi in ((0.8, 0.01, 0.05, 0.0625, 16, 8), (0.3, 0.02, 0.1, 0.125, 8, 3), (0.8, 0.01, 0.05, 0.0625, 16, 64)):
        occ_r = refshim.ssao(W, H, fx, fy, *gi, nrm, pos)
        occ_o = dgr._C.SSAO(W, H, fx, fy, *gi, nrm, pos)
        assert torch.equal(torch.nan_to_num(occ_o, nan=-7.0), torch.nan_to_num(occ_r, nan=-7.0)), ("ssao", gi)
        c_r, a_r = refshim.ssr(W, H, fx, fy, *gi, nrm, pos, rgb, alb, rough, met, F0)
        c_o, a_o = dgr._C.SSR(W, H, fx, fy, *gi, nrm, pos, rgb, alb, rough, met, F0)
        for x, y, nm in ((c_o, c_r, "ssr color"), (a_o, a_r, "ssr abd")):
            assert torch.equal(torch.isnan(x), torch.isnan(y)), (nm, gi)
            assert torch.equal(torch.nan_to_num(x, nan=-7.0), torch.nan_to_num(y, nan=-7.0)), (nm, gi)
