"""GPU: the fused PBR-stage frame (gigs_frame_forward / gigs_frame_backward, csrc/deferred.cu) against the unfused
operator path (GaussianRasterizer -> render() post-processing -> pbr_shading -> Gaussian_SSR -> loss, autograd),
which the other -m gpu tests pin to the reference's own kernels and to the CPU oracle.

Two input modes:
  * activated parameters (raw_params = 0): the rasterizer sees bit-identical inputs in both paths, so the G-buffer,
    the geometry chain and the SSAO must be BIT-EXACT; the deferred kernels replace chains of float32 framework ops
    and agree to rounding (1e-5), gradients to 1e-3 relative (north_star tolerance);
  * raw leaves (raw_params = 1): the getters are evaluated inside preprocess with torch's own arithmetic (expf,
    IEEE division, the summation order of its norm reductions), so the G-buffer is bit-identical here too; gates:
    G-buffer torch.equal, shaded images <= 1e-4 max-abs, loss 1e-5 relative, gradients 1e-3 relative — at 20k/400x300
    and at BASELINE configs[1]'s full size (300k, 800x800).
"""
import ctypes as C

import pytest
import torch

pytestmark = pytest.mark.gpu

import gpu_util as U
from gigs import frame as gframe
from gigs import renderer, scene, shade, step as gstep

DEV = "cuda:0"
GI8 = dict(radius=0.8, bias=0.01, thick=0.05, delta=0.0625, step=16, start=8)
GI64 = dict(GI8, start=64)


def _setup(P, W, H, seed=3, base_res=64, lut_res=64):
    raw = scene.make_scene(P, seed=seed, regime="trained")
    cam = scene.orbit_camera(1, 8, W, H).to(DEV)
    lut = shade.make_brdf_lut(lut_res, 64).to(DEV)
    rays = scene.canonical_rays(cam, DEV)
    gt = torch.rand(3, H, W, generator=torch.Generator().manual_seed(5)).to(DEV)
    bg = torch.zeros(3, device=DEV)
    return raw, cam, lut, rays, gt, bg


def _unfused(raw, cam, lut, rays, gt, bg, gi, base_res, **kw):
    params = gstep.GaussianParams(raw, DEV, light=scene.make_light(0, base_res=base_res))
    params.zero_grad()
    loss = gstep.training_step(params, cam, params.light(), lut, rays, gt, bg, gi, fused=False, **kw)
    return params, float(loss)


def _named_grads(params):
    out, o = {}, 0
    names = list(params.leaves) + ["light_diffuse"] + [f"light_spec{i}" for i in range(len(params.light_leaves) - 1)]
    for nm, t in zip(names, list(params.leaves.values()) + params.light_leaves):
        out[nm] = params.flat_grad[o:o + t.numel()].clone()
        o += t.numel()
    return out


@pytest.mark.parametrize("gi,metallic", [(GI8, True), (GI64, True), (GI8, False)])
def test_frame_with_activated_inputs_matches_the_operator_path(gi, metallic):
    P, W, H, base = 20000, 400, 300, 64
    raw, cam, lut, rays, gt, bg = _setup(P, W, H)
    ref, loss_ref = _unfused(raw, cam, lut, rays, gt, bg, gi, base, metallic=metallic)
    g = {k: (v.detach() if torch.is_tensor(v) else v) for k, v in ref.activated().items()}
    with torch.no_grad():
        res = renderer.pbr_forward(cam, g, ref.light(), lut, rays, bg, gi=gi, metallic=metallic)

    # the frame, fed the SAME activated tensors; gradients w.r.t. the activated material tensors
    light = ref.light()
    ws = gframe.workspace(P, W, H, DEV)
    keep = []
    f = gframe._fill(ws, cam, bg, g, False, 3, light, lut, rays, gt, gi, True, metallic, False, True, 1.0, 0.001, keep)
    loss = float(gframe.frame_forward(ws, f))
    # rasterizer / geometry chain / SSAO: same kernels, same inputs -> same bits
    for nm, key in (("color", "render"), ("opacity", "opacity_map"), ("depth", "depth_map"), ("albedo", "albedo_map"),
                    ("roughness", "roughness_map"), ("metallic", "metallic_map"), ("depth_pos", "depth_pos"),
                    ("occlusion", "occlusion_map")):
        assert torch.equal(ws.map(nm), res[key]), nm
    assert torch.equal(ws.map("mask").bool(), res["normal_mask"][0])
    # deferred kernels vs the framework-op chains they replace
    U.assert_close_map(ws.map("shade_normal"), res["normal_map"], 2e-5, "shade_normal")
    U.assert_close_map(ws.map("ssr_normal"), res["out_normal_view"], 2e-5, "ssr_normal")
    U.assert_close_map(ws.map("rough_remap"), res["roughness_remap"], 1e-6, "rough_remap")
    d_direct = (ws.map("render_direct") - res["render_direct"]).abs()
    d_rgb = (ws.map("render_rgb") - res["render_rgb"]).abs()
    assert d_direct.max().item() <= 1e-4, d_direct.max().item()
    if gi["start"] >= gi["step"]:
        assert d_rgb.max().item() <= 1e-4, d_rgb.max().item()
    else:  # the march thresholds may flip a hit for a pixel whose inputs differ in the last bit
        assert (d_rgb > 1e-4).float().mean().item() < 1e-3
    assert abs(loss - loss_ref) <= 1e-5 * abs(loss_ref)

    ga, gr, gm = (torch.zeros_like(g["albedo"]), torch.zeros_like(g["roughness"]), torch.zeros_like(g["metallic"]))
    gd = torch.zeros_like(light.diffuse)
    gs = [torch.zeros_like(s) for s in light.specular]
    gframe.frame_backward(ws, f, ga, gr, gm if metallic else None, gd, gs)
    # reference gradients w.r.t. the activated tensors through autograd on the operator path
    leaves = {k: g[k].clone().requires_grad_(True) for k in ("albedo", "roughness", "metallic")}
    g2 = dict(g, **leaves)
    lt = shade.Light(specular=[s.detach().clone().requires_grad_(True) for s in light.specular],
                     diffuse=light.diffuse.detach().clone().requires_grad_(True))
    out = renderer.pbr_forward(cam, g2, lt, lut, rays, bg, gi=gi, metallic=metallic)
    renderer.pbr_loss(out, gt).backward()
    tol = 1e-3 if gi["start"] >= gi["step"] else 5e-3
    U.assert_grad_close(ga, leaves["albedo"].grad, "albedo", tol)
    U.assert_grad_close(gr, leaves["roughness"].grad, "roughness", tol)
    if metallic:
        U.assert_grad_close(gm, leaves["metallic"].grad, "metallic", tol)
    U.assert_grad_close(gd, lt.diffuse.grad, "light.diffuse", tol)
    for i, (a, b) in enumerate(zip(gs, lt.specular)):
        if b.grad is not None and float(b.grad.abs().max()) > 0:
            U.assert_grad_close(a, b.grad, f"light.specular[{i}]", tol)


@pytest.mark.parametrize("P,W,H,base", [(20000, 400, 300, 64), (300000, 800, 800, 256)])
def test_frame_with_raw_leaves_matches_the_operator_path(P, W, H, base):
    """The headline path (getters fused into preprocess_kernel<RAW>) at a small size and at BASELINE configs[1]'s full
    size. The fused getters return torch's bits (common.cuh: torch_norm_*, tools/raw_getter_check.py), so the whole
    G-buffer is BIT-IDENTICAL to the operator path's (which the other tests pin to the reference's kernels), and the
    shaded images meet north_star's 1e-4 max-abs without a statistical allowance."""
    raw, cam, lut, rays, gt, bg = _setup(P, W, H, lut_res=64 if base == 64 else 256)
    ref, loss_ref = _unfused(raw, cam, lut, rays, gt, bg, GI64, base)
    fus = gstep.GaussianParams(raw, DEV, light=scene.make_light(0, base_res=base))
    fus.zero_grad()
    loss = float(gstep.training_step(fus, cam, fus.light(), lut, rays, gt, bg, GI64, fused=True, radiance=True))
    assert abs(loss - loss_ref) <= 1e-5 * abs(loss_ref)
    ws = fus.last_workspace
    with torch.no_grad():
        g = ref.activated()
        res = renderer.pbr_forward(cam, g, ref.light(), lut, rays, bg, gi=GI64)
    for nm, key in (("color", "render"), ("opacity", "opacity_map"), ("depth", "depth_map"), ("albedo", "albedo_map"),
                    ("roughness", "roughness_map"), ("metallic", "metallic_map"), ("depth_pos", "depth_pos"),
                    ("occlusion", "occlusion_map")):
        assert torch.equal(ws.map(nm), res[key]), f"{nm}: G-buffer not bit-identical with fused getters"
    assert torch.equal(ws.map("mask").bool(), res["normal_mask"][0])
    for nm, key in (("render_direct", "render_direct"), ("render_rgb", "render_rgb")):
        d = (ws.map(nm) - res[key]).abs().max().item()
        assert d <= 1e-4, f"{nm}: max abs diff {d} > 1e-4 in linear RGB"
    ga, gb = _named_grads(ref), _named_grads(fus)
    for nm in ga:
        if float(ga[nm].abs().max()) == 0.0:
            assert float(gb[nm].abs().max()) == 0.0, nm   # geometry / SH / opacity get no gradient in the PBR stage
        else:
            U.assert_grad_close(gb[nm], ga[nm], nm, 1e-3)


def test_frame_accumulates_gradients_and_is_deterministic_in_the_forward():
    P, W, H, base = 6000, 160, 96, 32
    raw, cam, lut, rays, gt, bg = _setup(P, W, H)
    p = gstep.GaussianParams(raw, DEV, light=scene.make_light(0, base_res=base))
    p.zero_grad()
    l1 = float(gstep.training_step(p, cam, p.light(), lut, rays, gt, bg, GI8))
    rgb1 = p.last_workspace.map("render_rgb").clone()
    g1 = p.flat_grad.clone()
    l2 = float(gstep.training_step(p, cam, p.light(), lut, rays, gt, bg, GI8))   # no zero_grad: accumulates
    assert l1 == l2 and torch.equal(rgb1, p.last_workspace.map("render_rgb"))
    assert torch.allclose(p.flat_grad, 2 * g1, rtol=1e-4, atol=1e-9)
    # loss_scale scales loss and gradients
    p.zero_grad()
    l3 = float(gstep.training_step(p, cam, p.light(), lut, rays, gt, bg, GI8, loss_scale=0.25))
    assert abs(l3 - 0.25 * l1) <= 1e-6 * abs(l1)
    assert torch.allclose(p.flat_grad, 0.25 * g1, rtol=1e-4, atol=1e-9)


def test_frame_workspace_growth_and_error_paths():
    from gigs import _lib
    L = _lib.load()
    P, W, H = 3000, 96, 80
    raw, cam, lut, rays, gt, bg = _setup(P, W, H)
    p = gstep.GaussianParams(raw, DEV, light=scene.make_light(0, base_res=32))
    ws = gframe.FrameWorkspace(P, W, H, DEV)   # fresh: no binning / sort buffers yet
    keep = []
    f = gframe._fill(ws, cam, bg, p.leaves, True, 3, p.light(), lut, rays, gt, GI8, True, True, False, True, 1.0,
                     0.001, keep)
    gframe._set_sort(ws, f)
    st = L.gigs_frame_forward(C.byref(f))
    assert st == _lib.GIGS_E_GROW and f.need_binning_bytes >= 4 * f.num_rendered and f.need_sort_bytes > 0
    assert b"too small" in L.gigs_last_error()
    loss = float(gframe.frame_forward(ws, f))          # grows and resumes
    assert loss == loss and ws.num_rendered == f.num_rendered > 0
    # backward without gradient buffers is an argument error, not a crash
    f.g_albedo = None
    assert L.gigs_frame_backward(C.byref(f)) < 0
    # a maps blob that is too small is refused
    f.maps_bytes = 16
    assert L.gigs_frame_forward(C.byref(f)) == -2


def test_frame_forward_only_without_ground_truth():
    P, W, H = 4000, 128, 96
    raw, cam, lut, rays, gt, bg = _setup(P, W, H)
    p = gstep.GaussianParams(raw, DEV, light=scene.make_light(0, base_res=32))
    ws = gframe.workspace(P, W, H, DEV)
    keep = []
    f = gframe._fill(ws, cam, bg, p.leaves, True, 3, p.light(), lut, rays, None, GI8, True, True, False, True, 1.0,
                     0.001, keep)
    gframe.frame_forward(ws, f)
    rgb = ws.map("render_rgb").clone()
    with torch.no_grad():
        res = renderer.pbr_forward(cam, p.activated(), p.light(), lut, rays, bg, gi=GI8, inference=False)
    assert ((rgb - res["render_rgb"]).abs() > 1e-4).float().mean().item() < 2e-3


def test_frame_waits_for_the_ground_truth_event():
    """gt_ready: the ground-truth image is copied on a side stream and the frame waits for it only before the loss."""
    P, W, H = 4000, 128, 96
    raw, cam, lut, rays, gt, bg = _setup(P, W, H)
    p = gstep.GaussianParams(raw, DEV, light=scene.make_light(0, base_res=32))
    p.zero_grad()
    l_ref = float(gstep.training_step(p, cam, p.light(), lut, rays, gt, bg, GI8))
    g_ref = p.flat_grad.clone()
    host = gt.cpu().pin_memory()
    side = torch.cuda.Stream()
    p.zero_grad()
    with torch.cuda.stream(side):
        torch.cuda._sleep(20_000_000)                  # make the copy late on purpose
        gt2 = host.to(DEV, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(side)
    gt2.record_stream(torch.cuda.current_stream())
    l2 = float(gstep.training_step(p, cam, p.light(), lut, rays, gt2, bg, GI8, gt_ready=ev))
    assert l2 == l_ref
    assert torch.allclose(p.flat_grad, g_ref, rtol=1e-5, atol=1e-10)


def test_step_with_the_light_as_base_cubemap_matches_autograd_through_build_mips():
    """training_step with GaussianParams(light_base=...): build_mips before the frame, the mips' backward after (on a
    side stream, fused path) == autograd through the op-by-op CubemapLight.build_mips feeding the operator path."""
    from gigs import light as GL
    P, W, H = 20000, 400, 300
    raw, cam, lut, rays, gt, bg = _setup(P, W, H)
    base = (torch.rand(6, 64, 64, 3, generator=torch.Generator().manual_seed(21)) * 0.5 + 0.25).to(DEV)
    # autograd reference: base -> cubemap_mip / specular_cubemap / diffuse_cubemap -> shading -> loss
    ref = gstep.GaussianParams(raw, DEV)
    cl = GL.CubemapLight(64, device=DEV, base=base.clone())
    cl.build_mips()
    res = renderer.pbr_forward(cam, ref.activated(), shade.Light(specular=cl.specular, diffuse=cl.diffuse), lut, rays,
                               bg, gi=GI64)
    loss_ref = renderer.pbr_loss(res, gt)
    loss_ref.backward()
    g_ref = cl.base.grad
    assert float(g_ref.abs().max()) > 0
    for fused in (True, False):
        p = gstep.GaussianParams(raw, DEV, light_base=base.clone())
        p.zero_grad()
        loss = gstep.training_step(p, cam, p.light(), lut, rays, gt, bg, GI64, fused=fused)
        torch.cuda.synchronize()
        assert abs(float(loss) - float(loss_ref)) <= 1e-5 * abs(float(loss_ref)), fused
        U.assert_grad_close(p.light_base.grad, g_ref, f"light_base (fused={fused})", 1e-3)
        U.assert_grad_close(p.leaves["albedo"].grad, ref.leaves["albedo"].grad, f"albedo (fused={fused})", 1e-3)
        assert float(p.prefiltered.texture_grads.abs().max()) == 0.0    # consumed and cleared
        # a second step accumulates (zero_grad not called): twice the gradient
        gstep.training_step(p, cam, p.light(), lut, rays, gt, bg, GI64, fused=fused)
        torch.cuda.synchronize()
        U.assert_grad_close(p.light_base.grad, 2.0 * g_ref, f"light_base, two steps (fused={fused})", 1e-3)


@pytest.mark.parametrize("metallic", [True, False])
def test_brdf_tv_prior_in_the_frame_matches_the_operator_path(metallic):
    """train.py:388-402: the fused loss / backward kernels with brdf_tv_weight against the framework-op restatement
    (renderer.masked_tv_loss) through autograd."""
    P, W, H, base = 20000, 400, 300, 64
    raw, cam, lut, rays, gt, bg = _setup(P, W, H)
    # without the prior first (the frame workspace is shared per shape: read stats right after the step that wrote them)
    p0 = gstep.GaussianParams(raw, DEV, light=scene.make_light(0, base_res=base))
    l0 = float(gstep.training_step(p0, cam, p0.light(), lut, rays, gt, bg, GI64, fused=True, metallic=metallic))
    out = {}
    for fused in (False, True):
        p = gstep.GaussianParams(raw, DEV, light=scene.make_light(0, base_res=base))
        p.zero_grad()
        out[fused] = (p, float(gstep.training_step(p, cam, p.light(), lut, rays, gt, bg, GI64, fused=fused,
                                                   metallic=metallic, brdf_tv_weight=1.0)))
    (pu, lu), (pf, lf) = out[False], out[True]
    tv = float(pf.last_workspace.map("stats")[5])
    # the prior is a visible part of the loss on this scene (not a no-op in the comparison)
    assert lf - l0 > 1e-4 * l0
    assert abs((lf - l0) - tv) <= 1e-3 * tv
    assert abs(lf - lu) <= 1e-5 * abs(lu)
    for k in ("albedo", "roughness") + (("metallic",) if metallic else ()):
        U.assert_grad_close(pf.leaves[k].grad, pu.leaves[k].grad, f"{k} with the BRDF TV prior", 1e-3)
        assert not torch.allclose(pf.leaves[k].grad, p0.leaves[k].grad)


def test_env_tv_prior_matches_oracle_and_step_paths():
    """train.py:406-420 on the base cubemap: C-ABI gigs_env_tv vs the CPU oracle (autograd), and through the step."""
    import gigs_oracle as O
    from gigs import light as GL
    base = (torch.rand(6, 64, 64, 3, generator=torch.Generator().manual_seed(31)) * 0.5 + 0.25)
    dirs = GL.envmap_dirs(device=DEV)
    assert (dirs.cpu() - O.envmap_dirs()).abs().max().item() <= 1e-6
    b = base.clone().to(DEV).requires_grad_(True)
    loss = GL.env_tv_loss(b, dirs)
    loss.backward()
    bo = base.clone().requires_grad_(True)
    lo = O.env_tv_loss(bo, dirs.cpu())
    lo.backward()
    assert abs(float(loss) - float(lo)) <= 1e-5 * abs(float(lo))
    U.assert_grad_close(b.grad.cpu(), bo.grad, "env tv d/d base", 1e-4)
    # through the training step: fused (direct accumulation) == operator path (autograd)
    P, W, H = 20000, 400, 300
    raw, cam, lut, rays, gt, bg = _setup(P, W, H)
    res = {}
    for fused in (False, True):
        p = gstep.GaussianParams(raw, DEV, light_base=base.clone())
        p.zero_grad()
        l = gstep.training_step(p, cam, p.light(), lut, rays, gt, bg, GI64, fused=fused, brdf_tv_weight=1.0,
                                env_tv_weight=0.01)
        torch.cuda.synchronize()
        res[fused] = (float(l), p.light_base.grad.clone())
    assert abs(res[True][0] - res[False][0]) <= 1e-5 * abs(res[False][0])
    U.assert_grad_close(res[True][1], res[False][1], "light_base with both priors", 1e-3)


def test_material_only_frame_is_bit_identical_where_it_computes():
    """GigsRasterFwd.material_only (the default of the fused training frame: no SH radiance image, no blended position)
    changes nothing else: the loss and every other map are bit-for-bit those of the full frame; the gradients agree to
    the order of the backward's atomic reductions (two runs of the SAME frame differ at that level)."""
    P, W, H, base = 20000, 400, 300, 64
    raw, cam, lut, rays, gt, bg = _setup(P, W, H)
    res = {}
    for radiance in (True, False):
        p = gstep.GaussianParams(raw, DEV, light=scene.make_light(0, base_res=base))
        p.zero_grad()
        l = gstep.training_step(p, cam, p.light(), lut, rays, gt, bg, GI8, fused=True, brdf_tv_weight=1.0,
                                radiance=radiance)
        ws = p.last_workspace
        res[radiance] = (float(l), p.flat_grad.clone(),
                         {k: ws.map(k).clone() for k in ("albedo", "roughness", "metallic", "normal", "normal_view",
                                                         "depth", "opacity", "occlusion", "render_rgb")})
    assert res[True][0] == res[False][0]
    assert ((res[True][1] - res[False][1]).norm() / res[True][1].norm()).item() <= 1e-6
    for k, v in res[True][2].items():
        assert torch.equal(v, res[False][2][k]), k


@pytest.mark.parametrize("gi", [GI8, GI64])
def test_dependent_launch_changes_no_result(gi):
    """gigs_set_dependent_launch: the frame's kernels launched with programmatic stream serialization (each waits for
    the previous grid before its first global access) against ordinary stream-ordered launches: loss and every map are
    bit-for-bit the same, the gradients agree to the order of the backward's atomic reductions; the launch counter
    (gigs_launch_count) sees the same number of kernels either way."""
    from gigs import _lib
    L = _lib.load()
    P, W, H, base = 20000, 400, 300, 64
    raw, cam, lut, rays, gt, bg = _setup(P, W, H)
    res = {}
    prev = L.gigs_set_dependent_launch(-1)
    try:
        for on in (0, 1):
            assert L.gigs_set_dependent_launch(on) in (0, 1)
            assert L.gigs_set_dependent_launch(-1) == on
            p = gstep.GaussianParams(raw, DEV, light=scene.make_light(0, base_res=base))
            p.zero_grad()
            c0 = int(L.gigs_launch_count())
            for _ in range(3):      # back-to-back frames: the next frame's first kernels run behind this one's last
                p.zero_grad()
                l = gstep.training_step(p, cam, p.light(), lut, rays, gt, bg, gi, fused=True, brdf_tv_weight=1.0)
            torch.cuda.synchronize()
            ws = p.last_workspace
            res[on] = (float(l), p.flat_grad.clone(), int(L.gigs_launch_count()) - c0,
                       {k: ws.map(k).clone() for k in ("albedo", "roughness", "metallic", "normal", "normal_view", "depth",
                                                       "opacity", "occlusion", "render_rgb", "ssr_color", "ssr_abd")})
    finally:
        L.gigs_set_dependent_launch(prev)
    assert res[0][0] == res[1][0]
    assert res[0][2] == res[1][2] and res[0][2] > 0
    assert ((res[0][1] - res[1][1]).norm() / res[0][1].norm()).item() <= 1e-6
    for k, v in res[0][3].items():
        assert torch.equal(v, res[1][3][k]), k


def test_clear_spans_zeroes_exactly_the_spans():
    """gigs_clear_spans (zero_grad of the spans a step wrote, one launch): odd offsets and lengths, empty spans,
    a span shorter than one 16-byte unit, and nothing outside the spans touched."""
    from gigs import _lib
    L = _lib.load()
    n = 1 << 20
    buf = torch.arange(1, n + 1, dtype=torch.float32, device=DEV)
    ref = buf.clone()
    spans = [(0, 1), (3, 3), (5, 7), (13, 1000), (4099, 4099 + 65537), (n - 9, n), (200001, 200002), (300000, 300019)]
    lo = (C.c_uint64 * len(spans))(*[a for a, _ in spans])
    hi = (C.c_uint64 * len(spans))(*[b for _, b in spans])
    _lib.check(L.gigs_clear_spans(buf.data_ptr(), len(spans), lo, hi, torch.cuda.current_stream().cuda_stream), "clear")
    for a, b in spans:
        ref[a:b] = 0
    assert torch.equal(buf, ref)
    # the same through GaussianParams.zero_grad on an offset view of the buffer
    view = buf[1:]            # 4-byte aligned only
    view.fill_(1.0)
    lo1 = (C.c_uint64 * 1)(2); hi1 = (C.c_uint64 * 1)(n - 3)
    _lib.check(L.gigs_clear_spans(view.data_ptr(), 1, lo1, hi1, torch.cuda.current_stream().cuda_stream), "clear")
    assert view[:2].eq(1).all() and view[n - 3:].eq(1).all() and view[2:n - 3].eq(0).all()
    assert L.gigs_clear_spans(buf.data_ptr(), 9, lo, hi, None) != 0     # more than 8 spans: refused


def test_lean_frame_without_a_march_leaves_the_geometry_chain_out():
    """radiance=False and start >= step: normal_from_depth / depth_pos are not consumed by anything (train.py:290-381)
    and the depth -> normal / position chain is not run (GigsFrame.skip_geometry): the two maps stay untouched, the
    loss and every other map are bit-for-bit those of the full frame, the gradients agree to the order of the atomics."""
    P, W, H, base = 20000, 400, 300, 64
    raw, cam, lut, rays, gt, bg = _setup(P, W, H)
    res = {}
    for radiance in (True, False):
        p = gstep.GaussianParams(raw, DEV, light=scene.make_light(0, base_res=base))
        p.zero_grad()
        ws = gframe.workspace(P, W, H, DEV)
        ws.map("depth_pos").fill_(123.0)
        ws.map("normal_from_depth").fill_(-7.0)
        l = gstep.training_step(p, cam, p.light(), lut, rays, gt, bg, GI64, fused=True, brdf_tv_weight=1.0,
                                radiance=radiance)
        assert p.last_workspace is ws
        res[radiance] = (float(l), p.flat_grad.clone(),
                         {k: ws.map(k).clone() for k in ("albedo", "roughness", "metallic", "normal", "normal_view", "depth",
                                                         "opacity", "occlusion", "render_direct", "render_rgb",
                                                         "ssr_color", "ssr_abd", "g_rgb", "shade_normal", "rough_remap",
                                                         "metal_used")},
                         ws.map("depth_pos").clone(), ws.map("normal_from_depth").clone())
    assert float(res[False][3].min()) == 123.0 and float(res[False][3].max()) == 123.0      # not written
    assert float(res[False][4].min()) == -7.0 and float(res[False][4].max()) == -7.0
    assert float(res[True][3].max()) != 123.0                                               # the full frame ran it
    assert res[True][0] == res[False][0]
    assert ((res[True][1] - res[False][1]).norm() / res[True][1].norm()).item() <= 1e-6
    for k, v in res[True][2].items():
        assert torch.equal(v, res[False][2][k]) or bool(((v == res[False][2][k]) | (v.isnan() & res[False][2][k].isnan())).all()), k


def test_lean_frame_takes_a_pixel_position_from_the_depth_map_when_its_epilogue_needs_one():
    """SSR's epilogue is +0 without its position only for F0, metallic in [0, 1]; with albedo > 1 (activated inputs,
    metallic 1 -> F0 = albedo) the full expression runs and needs the pixel's position, which the lean frame then
    evaluates from the depth map with the chain's own expressions (filters.cuh: depth_pos_pixel): every map equals the
    full frame's bit for bit."""
    P, W, H, base = 20000, 400, 300, 64
    raw, cam, lut, rays, gt, bg = _setup(P, W, H)
    ref = gstep.GaussianParams(raw, DEV, light=scene.make_light(0, base_res=base))
    g = {k: (v.detach() if torch.is_tensor(v) else v) for k, v in ref.activated().items()}
    g["albedo"] = (g["albedo"] * 3.0).contiguous()           # blended albedo well above 1 on most pixels
    g["metallic"] = torch.ones_like(g["metallic"])
    light = ref.light()
    out = {}
    for skip in (False, True):
        ws = gframe.workspace(P, W, H, DEV)
        keep = []
        f = gframe._fill(ws, cam, bg, g, False, 3, light, lut, rays, gt, GI64, True, True, False, True, 1.0, 0.001, keep,
                         skip_geometry=skip)
        loss = float(gframe.frame_forward(ws, f))
        out[skip] = (loss, {k: ws.map(k).clone() for k in ("ssr_color", "ssr_abd", "render_rgb", "occlusion")},
                     ws.map("F0").clone())
    assert float(out[False][2].max()) > 1.0
    assert out[False][0] == out[True][0]
    for k, v in out[False][1].items():
        w = out[True][1][k]
        assert bool(((v == w) | (v.isnan() & w.isnan())).all()), k
    # and the epilogue really depended on the position there: its sign pattern is not all +0
    assert bool((torch.signbit(out[False][1]["ssr_abd"]) | out[False][1]["ssr_abd"].isnan()).any())
