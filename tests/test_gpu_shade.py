"""GPU: filters, fused geometry chain, fused shading (forward + backward) and the drop-in autograd path against
the CPU oracle. The third-party semantics restated here (kornia filters, nvdiffrast textures) are "parity
unpinned" (not installable on this box); the oracle is the executable statement of the documented semantics."""
import pytest
import torch

pytestmark = pytest.mark.gpu

import diff_gaussian_rasterization as dgr
import gigs_oracle as O
import gpu_util as U
from gigs import renderer, scene, shade, step as gstep

DEV = "cuda:0"


def test_median_and_bilateral_vs_oracle():
    gen = torch.Generator().manual_seed(0)
    x = torch.rand(3, 37, 53, generator=gen)
    x[1, 10, 10] = float("nan")
    x[2, 0, 0] = float("inf")
    m_o = dgr.median_blur3x3(x.to(DEV)).cpu()
    m_r = O.median3x3(x)
    assert torch.equal(torch.isnan(m_o), torch.isnan(m_r))
    assert torch.equal(torch.nan_to_num(m_o), torch.nan_to_num(m_r))
    y = torch.rand(3, 37, 53, generator=gen)
    assert (dgr.bilateral_blur3x3(y.to(DEV), 1.0, 3.0).cpu() - O.bilateral3x3(y, 1.0, 3.0)).abs().max() < 1e-5


def test_median_backward_routes_gradient_to_the_selected_element():
    gen = torch.Generator().manual_seed(1)
    x = torch.rand(2, 20, 24, generator=gen)          # distinct values -> unique median element
    xg = x.to(DEV).requires_grad_(True)
    w = torch.rand(2, 20, 24, generator=gen)
    (dgr.median_blur3x3(xg) * w.to(DEV)).sum().backward()
    xc = x.clone().requires_grad_(True)
    (O.median3x3(xc) * w).sum().backward()
    assert (xg.grad.cpu() - xc.grad).abs().max() < 1e-6


def test_fused_geometry_chain_vs_oracle_and_vs_unfused():
    P, W, H = 8000, 211, 130
    raw = scene.make_scene(P, seed=3)
    g, cam = scene.activate(raw, DEV), scene.orbit_camera(1, 8, W, H).to(DEV)
    fo = U.ours_forward(g, cam, torch.zeros(3, device=DEV))
    fx, fy = W / (2 * cam.tanfovx), H / (2 * cam.tanfovy)
    V = cam.world_view_transform
    n_f, p_f = dgr.geometry_chain(W, H, fx, fy, V, fo["depth"], True)
    # unfused composition of our own kernels: identical bits
    n_m, p_m = dgr._C.depth_to_normal(W, H, fx, fy, V, dgr.median_blur3x3(fo["depth"]))
    assert torch.equal(n_f, dgr.bilateral_blur3x3(n_m, 1.0, 3.0))
    assert torch.equal(p_f, dgr.median_blur3x3(p_m))
    n_o, p_o = O.geometry_chain(W, H, fx, fy, V.cpu(), fo["depth"].cpu())
    assert (n_f.cpu() - n_o).abs().max() < 1e-4 and (p_f.cpu() - p_o).abs().max() < 1e-4
    n_z, p_z = dgr.geometry_chain(W, H, fx, fy, V, fo["depth"], False)
    assert float(n_z.abs().max()) == 0.0 and float(p_z.abs().max()) == 0.0


def _shade_inputs(H, W, seed, metallic=True):
    gen = torch.Generator().manual_seed(seed)
    F = torch.nn.functional
    n = F.normalize(torch.randn(H, W, 3, generator=gen), dim=-1)
    v = F.normalize(torch.randn(H, W, 3, generator=gen), dim=-1)
    n[0, :4] = torch.tensor([[1.0, 1.0, 1.0], [1.0, -1.0, 1.0], [-1.0, 1.0, -1.0], [1.0, 1.0, -1.0]]) / 3 ** 0.5  # corners
    n[1, :3] = torch.tensor([[1.0, 1.0, 0.0], [0.0, 1.0, 1.0], [1.0, 0.0, -1.0]]) / 2 ** 0.5                    # edges
    alb = torch.rand(H, W, 3, generator=gen)
    rough = torch.rand(H, W, 1, generator=gen) * 0.96 + 0.04
    met = torch.rand(H, W, 1, generator=gen) if metallic else None
    occ = torch.rand(H, W, 1, generator=gen)
    mask = torch.rand(H, W, 1, generator=gen) > 0.2
    return n, v, alb, rough, met, occ, mask


@pytest.mark.parametrize("tone,gamma,metallic", [(False, True, True), (True, False, True), (False, False, False)])
def test_fused_shading_forward_and_backward_vs_oracle(tone, gamma, metallic):
    H, W = 48, 64
    n, v, alb, rough, met, occ, mask = _shade_inputs(H, W, 0, metallic)
    light_h = scene.make_light(0, base_res=64)
    lut = shade.make_brdf_lut(64, 128)
    # oracle (autograd)
    lt = dict(diffuse=light_h["diffuse"].clone().requires_grad_(True),
              specular=[s.clone().requires_grad_(True) for s in light_h["specular"]])
    a_h, r_h = alb.clone().requires_grad_(True), rough.clone().requires_grad_(True)
    m_h = met.clone().requires_grad_(True) if metallic else None
    ro = O.pbr_shading(lt, n, v, a_h, r_h, mask, tone=tone, gamma=gamma, occlusion=occ, metallic=m_h, brdf_lut=lut)
    gen = torch.Generator().manual_seed(5)
    wts = torch.randn(H, W, 3, generator=gen)
    (ro["render_rgb"] * wts).sum().backward()
    # ours
    light = shade.Light([s.to(DEV).requires_grad_(True) for s in light_h["specular"]],
                        light_h["diffuse"].to(DEV).requires_grad_(True))
    a_d, r_d = alb.to(DEV).requires_grad_(True), rough.to(DEV).requires_grad_(True)
    m_d = met.to(DEV).requires_grad_(True) if metallic else None
    so = shade.pbr_shading(light, n.to(DEV), v.to(DEV), a_d, r_d, mask.to(DEV), tone=tone, gamma=gamma,
                           occlusion=occ.to(DEV), metallic=m_d, brdf_lut=lut.to(DEV))
    for k in ("render_rgb", "diffuse_rgb", "specular_rgb", "diffuse_light"):
        assert (so[k].cpu() - ro[k].detach()).abs().max() < 1e-4, k
    (so["render_rgb"] * wts.to(DEV)).sum().backward()
    U.assert_grad_close(a_d.grad.cpu(), a_h.grad, "albedo")
    U.assert_grad_close(r_d.grad.cpu(), r_h.grad, "roughness", rel_tol=5e-3)
    if metallic:
        U.assert_grad_close(m_d.grad.cpu(), m_h.grad, "metallic")
    U.assert_grad_close(light.diffuse.grad.cpu(), lt["diffuse"].grad, "diffuse texels")
    for i, s in enumerate(light.specular):
        U.assert_grad_close(s.grad.cpu(), lt["specular"][i].grad, f"specular level {i}")


def test_full_pbr_frame_runs_and_matches_composed_oracle_pieces():
    """The drop-in autograd path end to end (train.py:266-422 semantics): forward values of every stage equal the
    oracle's stage applied to OUR previous-stage output; gradients reach every parameter group."""
    P, W, H = 4000, 128, 96
    raw = scene.make_scene(P, seed=13)
    cam = scene.orbit_camera(3, 8, W, H).to(DEV)
    params = gstep.GaussianParams(raw, DEV, light=scene.make_light(0, base_res=64))
    lut = shade.make_brdf_lut(64, 64).to(DEV)
    rays = scene.canonical_rays(cam, DEV)
    gi = dict(radius=0.8, bias=0.01, thick=0.05, delta=0.0625, step=16, start=8)
    g = params.activated()
    res = renderer.pbr_forward(cam, g, params.light(), lut, rays, torch.zeros(3, device=DEV), gi=gi)
    fx, fy = W / (2 * cam.tanfovx), H / (2 * cam.tanfovy)
    raw_nv = U.ours_forward({k: (v.detach() if torch.is_tensor(v) else v) for k, v in g.items()}, cam,
                            torch.zeros(3, device=DEV))["normal_view"]          # SSAO consumes the RAW view normals
    occ_o = O.ssao(W, H, fx, fy, 0.8, 0.01, 0.05, 0.0625, 16, 8, raw_nv.cpu(), res["depth_pos"].cpu())
    # CPU oracle vs GPU: float32 without FMA contraction flips an isolated probe's hit test on a few pixels (one
    # probe = one direction weight <= 4e-3 of occlusion); the bit-exact SSAO gate is against the reference CUDA build
    d_occ = (res["occlusion_map"].cpu() - occ_o).abs()
    assert float((d_occ > 1e-4).float().mean()) < 2e-3 and float(d_occ.max()) < 1.3e-2
    assert torch.isfinite(res["render_rgb"]).all()
    gt = torch.rand(3, H, W, device=DEV)
    loss = renderer.pbr_loss(res, gt)
    loss.backward()
    for k in ("albedo", "roughness", "metallic"):
        assert float(params.leaves[k].grad.abs().sum()) > 0, k
    for t in params.light_leaves:
        assert float(t.grad.abs().sum()) > 0
    # the PBR-stage loss has no path to geometry (train.py:385-420): exactly zero, as in the reference
    for k in ("xyz", "log_scale", "rot", "opacity", "f_dc", "f_rest"):
        assert float(params.leaves[k].grad.abs().max()) == 0.0, k


def test_stage1_style_loss_reaches_geometry():
    """Colour + normal losses (train.py stage 1) exercise the full backward path through the autograd wrapper."""
    P, W, H = 3000, 96, 96
    raw = scene.make_scene(P, seed=21)
    cam = scene.orbit_camera(0, 8, W, H).to(DEV)
    params = gstep.GaussianParams(raw, DEV)
    rr = renderer.render(cam, params.activated(), torch.rand(3, device=DEV), derive_normal=True, start=64)
    loss = (rr["render"] - 0.5).abs().mean() + rr["normal_map"].abs().mean()
    loss.backward()
    for k in ("xyz", "log_scale", "rot", "opacity", "f_dc", "f_rest", "normal"):
        assert float(params.leaves[k].grad.abs().sum()) > 0, k
    assert rr["viewspace_points"].grad is not None and rr["viewspace_points"].grad.shape == (P, 3)
