"""GPU: the remaining BASELINE.json configs as parity / property cases (bench.py times configs[1] only).
C3 bicycle-shaped 6M Gaussians at 1237x822; C4 multi-view step (K views, gradient accumulation == what the
all-reduce sums); C5 camera-sharded inference sweep."""
import pytest
import torch

pytestmark = pytest.mark.gpu

import gpu_util as U
import refshim
from gigs import renderer, scene, shade, step as gstep

DEV = "cuda:0"
GI = dict(radius=0.8, bias=0.01, thick=0.05, delta=0.0625, step=16, start=64)


@pytest.mark.skipif(not refshim.available(), reason="oracle/_ref not built")
def test_c3_bicycle_shape_full_size_bit_exact_vs_reference():
    """~6M Gaussians, SH degree 3, 1237x822 (W,H not multiples of 16, 44-bit keys), fx = fy = 1040."""
    P, W, H = 6_000_000, 1237, 822
    raw = scene.make_scene(P, seed=0, regime="trained", shape="bicycle")
    g = scene.activate(raw, DEV)
    del raw
    cam = scene.look_at_camera([4.0, 0.0, 1.0], [0.0, 0.0, 0.0], W, H, fx=1040.0).to(DEV)
    bg = torch.zeros(3, device=DEV)
    ref = refshim.RefRasterizer()
    ro = ref.forward(g, cam, bg)
    rs = ref.state()
    fo = U.ours_forward(g, cam, bg)
    st = U.decode_state(fo, P, W, H)
    assert fo["num_rendered"] == ro["num_rendered"] > 0
    assert torch.equal(fo["radii"], ro["radii"])
    for k in ("keys_sorted", "point_list", "ranges", "n_contrib"):
        assert torch.equal(st[k], rs[k]), k
    for k in ("color", "depth", "albedo", "roughness", "metallic", "normal", "opacity", "pos"):
        U.assert_close_map(fo[k], ro[k], 1e-4, k)
    gen = torch.Generator().manual_seed(1)
    grads = {k: (torch.randn(c, H, W, generator=gen) / (W * H)).to(DEV) for k, c in
             (("depth", 1), ("color", 3), ("opacity", 1), ("normal", 3), ("albedo", 3), ("roughness", 1),
              ("metallic", 1))}
    rb = ref.backward(g, cam, bg, ro["radii"], grads)
    rb2 = ref.backward(g, cam, bg, ro["radii"], grads)      # the reference's own run-to-run atomic-order noise
    ob = U.ours_backward(g, cam, bg, fo, grads)
    for k in ("means2D", "colors", "opacity", "albedo", "means3D", "cov3D", "sh", "scales", "rotations"):
        a, b = ob[k].float().reshape(-1), rb[k].float().reshape(-1)
        rel = ((a - b).norm() / (b.norm() + 1e-9 * a.numel() ** 0.5)).item()
        assert rel <= 1e-3, (k, rel)                          # north_star gate: 1e-3 relative
        # per element: within 1e-3 of the tensor's scale, or within a small multiple of the reference's own
        # nondeterminism (dL/dcov3D amplifies atomic-order noise by 1/det^2 for distant, ill-conditioned splats)
        noise = (rb[k].float().reshape(-1) - rb2[k].float().reshape(-1)).abs().max().item()
        mx = (a - b).abs().max().item()
        assert mx <= max(1e-3 * b.abs().max().item() + 1e-8, 8.0 * noise), (k, mx, noise)
    ref.close()


def test_c4_multi_view_step_equals_mean_of_single_view_gradients():
    """K-view step on one rank == (1/K) * sum of K single-view steps; the sharded form only changes WHO adds."""
    P, W, H, K = 20000, 160, 128, 4
    raw = scene.make_scene(P, seed=3)
    light_h = scene.make_light(0, base_res=64)
    lut = shade.make_brdf_lut(64, 64).to(DEV)
    cams = [scene.orbit_camera(k, 8, W, H).to(DEV) for k in range(K)]
    gen = torch.Generator().manual_seed(0)
    gts = [torch.rand(3, H, W, generator=gen).to(DEV) for _ in range(K)]
    bg = torch.zeros(3, device=DEV)
    rays = scene.canonical_rays(cams[0], DEV)
    params = gstep.GaussianParams(raw, DEV, light=light_h)
    total = gstep.multi_view_step(params, cams, params.light(), lut, lambda c: rays, gts, bg, GI)
    multi = params.flat_grad.clone()
    acc = torch.zeros_like(multi)
    losses = []
    for k in range(K):
        params.zero_grad()
        losses.append(gstep.training_step(params, cams[k], params.light(), lut, rays, gts[k], bg, GI))
        acc += params.flat_grad
    U.assert_grad_close(multi, acc / K, "multi-view gradient")
    assert abs(float(total) - float(sum(losses)) / K) < 1e-5
    # two-"rank" emulation in one process: disjoint view shards add up to the same buffer
    shards = []
    for r in range(2):
        params.zero_grad()
        for k in gstep.shard_views(K, r, 2):
            gstep.training_step(params, cams[k], params.light(), lut, rays, gts[k], bg, GI, loss_scale=1.0 / K)
        shards.append(params.flat_grad.clone())
    U.assert_grad_close(shards[0] + shards[1], multi, "sharded sum")


def test_c5_camera_sharded_inference_sweep():
    """Eval / relight sweep: inference=True forward only, views round-robin over ranks, results independent of
    the sharding (no data-path collective)."""
    P, W, H, V = 15000, 200, 160, 6
    raw = scene.make_scene(P, seed=8)
    g = scene.activate(raw, DEV)
    light_h = scene.make_light(1, base_res=64)
    light = shade.Light([s.to(DEV) for s in light_h["specular"]], light_h["diffuse"].to(DEV))
    lut = shade.make_brdf_lut(64, 64).to(DEV)
    cams = [scene.orbit_camera(k, V, W, H).to(DEV) for k in range(V)]
    rays = scene.canonical_rays(cams[0], DEV)
    bg = torch.zeros(3, device=DEV)
    gi8 = dict(GI, start=8)

    def view(k):
        with torch.no_grad():
            r = renderer.pbr_forward(cams[k], g, light, lut, rays, bg, gi=gi8, inference=True)
        return r["render_rgb"]

    full = [view(k) for k in range(V)]
    for world in (2, 4):
        got = {}
        for r in range(world):
            for k in gstep.shard_views(V, r, world):
                got[k] = view(k)
        assert sorted(got) == list(range(V))
        for k in range(V):
            assert torch.equal(torch.nan_to_num(got[k]), torch.nan_to_num(full[k]))
    assert all(torch.isfinite(f).all() for f in full)
    # the fused forward-only frame renders the same views (eval sweeps do not need the operator path)
    from gigs import frame as gframe
    for k in (0, 3):
        ws = gframe.pbr_frame_eval(g, cams[k], light, lut, rays, bg, gi8, inference=True)
        d = (ws.map("render_rgb") - full[k]).abs()
        assert (torch.nan_to_num(d) > 1e-4).float().mean().item() < 2e-3, float(torch.nan_to_num(d).max())
        assert torch.equal(ws.map("roughness"), renderer.render(cams[k], g, bg, inference=True, derive_normal=True,
                                                                **gi8)["roughness_map"])
    # inference adds the residual transmittance to roughness (forward.cu:612-613)
    a = U.ours_forward(g, cams[0], bg, inference=True)
    b = U.ours_forward(g, cams[0], bg, inference=False)
    lay_T = U.decode_state(b, P, W, H)["final_T"].reshape(1, H, W)
    assert torch.allclose(a["roughness"], b["roughness"] + lay_T, atol=1e-6)
