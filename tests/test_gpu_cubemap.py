"""GPU parity of the cubemap prefilter (CubemapLight.build_mips): our kernels (C-ABI gigs_cubemap_* /
gigs_light_*) against
  (1) the reference's OWN kernels (pbr/renderutils/c_src/cubemap.cu, compiled unmodified into
      oracle/_ref/libgigs_ref_cubemap.so): bounds bit-exact, filters to float32 summation-order rounding;
  (2) the CPU oracle (oracle/gigs_oracle.py, dense / COO restatement);
and the fused PrefilteredLight against the op-by-op autograd CubemapLight.
Tolerances: textures 5e-6 relative to the texture's scale (measured <= 3e-6; only the order of the sums and the last
float division of ndfGGX differ from the reference), gradients 1e-5 relative norm-wise (the reference's own backward
is atomicAdd-ordered, i.e. nondeterministic at that level)."""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "gi-gs_b200"), os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

pytestmark = pytest.mark.gpu

TEX_TOL = 5e-6
DIFFUSE_TOL = 1e-5   # 1536..6144 same-sign terms per texel: the reference's own serial float sum is the coarser one
GRAD_TOL = 1e-5


def _dev():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    return torch.device("cuda:0")


def _ref():
    import refshim
    if not refshim.cubemap_available():
        pytest.skip("oracle/_ref/libgigs_ref_cubemap.so not built")
    return refshim.RefCubemap()


def _rel(a, b):
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def _cube(res, seed, dev, kind="uniform"):
    g = torch.Generator().manual_seed(seed)
    if kind == "uniform":
        x = torch.rand(6, res, res, 3, generator=g) * 0.5 + 0.25          # CubemapLight init (light.py:103-105)
    else:                                                                  # HDR-like: lognormal + a hot texel cluster
        x = torch.exp(torch.randn(6, res, res, 3, generator=g))
        x[2, res // 3:res // 3 + 2, res // 2:res // 2 + 2] = 1.0e4
    return x.to(dev)


@pytest.mark.parametrize("res,rough", [(16, 1.0), (16, 0.5), (32, 0.5), (32, 0.29), (64, 0.08), (64, 0.36), (128, 0.22)])
def test_bounds_bit_exact_vs_reference(res, rough):
    dev, ref = _dev(), _ref()
    from gigs import light as GL
    c, b = GL.specular_bounds(res, rough, 0.99, dev)
    rb = ref.specular_bounds(res, c, dev).view(6, res, res, 6, 4)
    assert torch.equal(b.float(), rb), f"bounds differ in {(b.float() != rb).sum().item()} entries"


@pytest.mark.parametrize("res", [16, 32])
def test_bounds_vs_oracle(res):
    dev = _dev()
    import gigs_oracle as O
    from gigs import light as GL
    c, b = GL.specular_bounds(res, 0.5, 0.99, dev)
    ob = O.specular_bounds(res, c)
    assert torch.equal(b.cpu().long(), ob)


@pytest.mark.parametrize("res,kind", [(16, "uniform"), (16, "hdr"), (32, "uniform")])
def test_diffuse_vs_reference_and_oracle(res, kind):
    dev, ref = _dev(), _ref()
    import gigs_oracle as O
    from gigs import light as GL
    x = _cube(res, 3, dev, kind).requires_grad_(True)
    out = GL.diffuse_cubemap(x)
    r = ref.diffuse_cubemap_fwd(x.detach())
    scale = r.abs().max().item()
    assert (out - r).abs().max().item() <= DIFFUSE_TOL * scale
    assert (out.cpu() - O.diffuse_cubemap(x.detach().cpu())).abs().max().item() <= DIFFUSE_TOL * scale
    g = torch.randn(6, res, res, 3, generator=torch.Generator().manual_seed(4)).to(dev)
    out.backward(g)
    rg = ref.diffuse_cubemap_bwd(x.detach(), g)
    assert _rel(x.grad, rg) <= GRAD_TOL
    assert _rel(x.grad.cpu(), O.diffuse_cubemap_backward(g.cpu())) <= GRAD_TOL


@pytest.mark.parametrize("res,rough,kind", [(16, 1.0, "uniform"), (32, 0.5, "uniform"), (64, 0.36, "hdr"),
                                            (64, 0.08, "uniform"), (128, 0.22, "uniform"), (256, 0.08, "uniform")])
def test_specular_vs_reference(res, rough, kind):
    dev, ref = _dev(), _ref()
    from gigs import light as GL
    x = _cube(res, 5, dev, kind).requires_grad_(True)
    out = GL.specular_cubemap(x, rough, 0.99)
    c, b = GL.specular_bounds(res, rough, 0.99, dev)
    r, rw, rb = ref.specular_cubemap(x.detach(), rough, 0.99, c)
    scale = r.abs().max().item()
    err = (out - r).abs().max().item()
    assert err <= TEX_TOL * scale, f"specular res {res} roughness {rough}: {err / scale:.3e} of scale"
    g = torch.randn(6, res, res, 3, generator=torch.Generator().manual_seed(6)).to(dev)
    out.backward(g)
    rg = ref.specular_cubemap_grad(x.detach(), rb, rw, g, rough, c)
    assert _rel(x.grad, rg) <= GRAD_TOL


@pytest.mark.parametrize("res,rough", [(16, 1.0), (32, 0.5), (64, 0.08)])
def test_specular_vs_oracle(res, rough):
    dev = _dev()
    import gigs_oracle as O
    from gigs import light as GL
    x = _cube(res, 7, dev).requires_grad_(True)
    out = GL.specular_cubemap(x, rough, 0.99)
    o, w = O.specular_cubemap(x.detach().cpu(), rough, 0.99)
    # at roughness 0.08 the GGX term amplifies one ulp of dot(V,H) to 0.3 % of a central weight; the oracle emulates
    # the CUDA arithmetic (FMA contraction order, correctly rounded sqrt) and agrees to rounding there too
    assert (out.cpu() - o).abs().max().item() <= 4 * TEX_TOL * o.abs().max().item()
    g = torch.randn(6, res, res, 3, generator=torch.Generator().manual_seed(8)).to(dev)
    out.backward(g)
    og = O.specular_cubemap_backward(g.cpu(), rough, 0.99)
    assert _rel(x.grad.cpu(), og) <= 3 * GRAD_TOL


def test_mip_forward_backward_vs_oracle():
    dev = _dev()
    import gigs_oracle as O
    from gigs import light as GL
    x = _cube(64, 9, dev).requires_grad_(True)
    y = GL.cubemap_mip(x)
    assert (y.cpu() - O.cubemap_mip(x.detach().cpu())).abs().max().item() <= 1e-7
    # the framework's own average pool, as the reference calls it (light.py:56-60)
    yt = torch.nn.functional.avg_pool2d(x.detach().permute(0, 3, 1, 2), (2, 2)).permute(0, 2, 3, 1).contiguous()
    assert torch.equal(y, yt)
    g = torch.randn(6, 32, 32, 3, generator=torch.Generator().manual_seed(10)).to(dev)
    y.backward(g)
    og = O.cubemap_mip_backward(g.cpu())
    # bilinear tap weights come from float texture-coordinate arithmetic (u * w - 0.5): rounding-level differences
    assert (x.grad.cpu() - og).abs().max().item() <= 1e-5 * og.abs().max().item()


@pytest.mark.parametrize("base_res,stored", [(64, True), (64, False), (256, True), (256, False), (16, True)])
def test_prefiltered_light_vs_operator_path(base_res, stored):
    """fused build (2 launches) / backward (6) == CubemapLight.build_mips op by op through autograd."""
    dev = _dev()
    from gigs import light as GL
    base = _cube(base_res, 11, dev)
    op = GL.CubemapLight(base_res, device=dev, base=base.clone())
    op.build_mips()
    fl = GL.PrefilteredLight(base.clone(), stored_operators=stored).build()
    assert len(fl.specular) == len(op.specular)
    for a, b in zip(fl.specular, op.specular):
        assert (a - b).abs().max().item() <= TEX_TOL * b.abs().max().item()
    assert (fl.diffuse - op.diffuse).abs().max().item() <= DIFFUSE_TOL * op.diffuse.abs().max().item()
    gen = torch.Generator().manual_seed(12)
    gs = [torch.randn(t.shape, generator=gen).to(dev) for t in op.specular]
    gd = torch.randn(op.diffuse.shape, generator=gen).to(dev)
    loss = sum((t * g).sum() for t, g in zip(op.specular, gs)) + (op.diffuse * gd).sum()
    loss.backward()
    for t, g in zip(fl.specular, gs):
        t.grad.copy_(g)
    fl.diffuse.grad.copy_(gd)
    gb = fl.backward(torch.empty_like(base), accumulate=False)
    assert _rel(gb, op.base.grad) <= GRAD_TOL
    assert fl.texture_grads.abs().max().item() == 0.0          # cleared for the next step
    # accumulate=True adds
    for t, g in zip(fl.specular, gs):
        t.grad.copy_(g)
    fl.diffuse.grad.copy_(gd)
    gb2 = fl.backward(gb.clone(), accumulate=True)
    assert _rel(gb2, 2 * op.base.grad) <= GRAD_TOL


@pytest.mark.parametrize("stored", [True, False])
def test_prefiltered_light_vs_oracle(stored):
    dev = _dev()
    import gigs_oracle as O
    from gigs import light as GL
    base = _cube(64, 13, dev)
    fl = GL.PrefilteredLight(base, stored_operators=stored).build()
    om = O.build_mips(base.cpu())
    for lvl, (a, b) in enumerate(zip(fl.specular, om["specular"])):
        assert (a.detach().cpu() - b).abs().max().item() <= 4 * TEX_TOL * b.abs().max().item(), f"level {lvl}"
    assert (fl.diffuse.detach().cpu() - om["diffuse"]).abs().max().item() <= DIFFUSE_TOL * om["diffuse"].abs().max().item()
    for a, b in zip(fl.chain, om["chain"]):
        assert (a[..., :3].cpu() - b).abs().max().item() <= 2e-7
    gen = torch.Generator().manual_seed(14)
    gs = [torch.randn(t.shape, generator=gen) for t in om["specular"]]
    gd = torch.randn(om["diffuse"].shape, generator=gen)
    for t, g in zip(fl.specular, gs):
        t.grad.copy_(g.to(dev))
    fl.diffuse.grad.copy_(gd.to(dev))
    gb = fl.backward(torch.empty_like(base), accumulate=False)
    og = O.build_mips_backward(64, gs, gd)
    assert _rel(gb.cpu(), og) <= 1e-4


def test_full_size_properties():
    """base_res 256 (the reference's training setting): constant light stays constant through every filter; the
    filters are linear; <g, F x> == <F^T g, x> for the GGX and cosine filters (their backward IS the adjoint)."""
    dev = _dev()
    from gigs import light as GL
    const = torch.full((6, 256, 256, 3), 0.7, device=dev)
    fl = GL.PrefilteredLight(const).build()
    for t in fl.specular:
        assert (t - 0.7).abs().max().item() <= 2e-6
    x1, x2 = _cube(256, 15, dev), _cube(256, 16, dev, "hdr")
    a = [t.detach().clone() for t in GL.PrefilteredLight(x1).build().specular]
    b = [t.detach().clone() for t in GL.PrefilteredLight(x2).build().specular]
    ab = GL.PrefilteredLight(x1 + 2.0 * x2).build().specular
    for u, v, w in zip(a, b, ab):
        assert (u + 2.0 * v - w).abs().max().item() <= 1e-5 * w.abs().max().item()
    for res, rough in ((256, 0.08), (128, 0.22), (64, 0.36), (32, 0.5), (16, 1.0)):
        x = _cube(res, 17, dev).requires_grad_(True)
        g = torch.randn(6, res, res, 3, generator=torch.Generator().manual_seed(18)).to(dev)
        y = GL.specular_cubemap(x, rough)
        y.backward(g)
        lhs, rhs = (y.detach().double() * g.double()).sum().item(), (x.grad.double() * x.detach().double()).sum().item()
        assert abs(lhs - rhs) <= 1e-5 * max(abs(lhs), abs(rhs), 1.0), (res, rough, lhs, rhs)


def test_stored_operator_matches_reference_wsum_bit_for_bit():
    """The stored forward operator is evaluated with the reference's arithmetic in the reference's order: its wsum
    equals the reference kernel's fourth output channel exactly (levels of base_res 64: roughness 0.08 / 0.5 / 1.0)."""
    dev, ref = _dev(), _ref()
    from gigs import light as GL
    base = _cube(64, 19, dev)
    fl = GL.PrefilteredLight(base).build()
    for lvl in range(fl.layout.n_levels):
        res, c = fl.layout.res[lvl], fl.layout.cutoff[lvl]
        rough = fl.layout.roughness[lvl]
        b = ref.specular_bounds(res, c, dev)
        o = ref.specular_cubemap_fwd(fl.chain[lvl][..., :3].contiguous(), b, rough, c)
        assert torch.equal(o[..., 3], fl.wsum[lvl]), f"level {lvl}: wsum differs in {(o[..., 3] != fl.wsum[lvl]).sum().item()} texels"
        r = o[..., 0:3] / o[..., 3:]
        assert (fl.specular[lvl] - r).abs().max().item() <= TEX_TOL * r.abs().max().item()


def test_prefiltered_light_storage_policy_and_base_res_512():
    """The stored filter operators grow ~8x per doubling of the base resolution (1.4 GB at 256, 11.3 GB at 512):
    "auto" stores them only inside its memory budget (8 GB, a quarter of the free memory) and falls back to the
    on-the-fly filter otherwise; both variants give the same textures at base_res 512."""
    from gigs import light as GL
    dev = "cuda:0"
    base = _cube(512, 21, dev)
    fly = GL.PrefilteredLight(base)                       # "auto": 11.3 GB is over the budget
    assert not fly.stored_operators and fly.stored_operator_bytes == 0 and fly.weights is None
    assert GL.PrefilteredLight(_cube(256, 3, dev)).stored_operators      # 1.4 GB: stored
    auto = GL.PrefilteredLight(base, stored_operators=True)
    assert auto.stored_operators and 8 << 30 < auto.stored_operator_bytes < 16 << 30
    auto.build(); fly.build()
    torch.cuda.synchronize()
    assert len(auto.specular) == 6 and auto.specular[0].shape[1] == 512
    for a, b in zip(auto.specular + [auto.diffuse], fly.specular + [fly.diffuse]):
        assert float((a.detach() - b.detach()).abs().max()) <= 5e-6 * float(b.detach().abs().max()) + 1e-7
    const = GL.PrefilteredLight(torch.full((6, 512, 512, 3), 0.7, device=dev), stored_operators=False)
    const.build()
    for t in const.specular:            # (the cosine-filtered diffuse level is not normalised in the reference either)
        assert float((t.detach() - 0.7).abs().max()) < 2e-5
    # a budget below the operators' size: automatic fall-back
    old = GL.PrefilteredLight.STORED_MAX_BYTES
    try:
        GL.PrefilteredLight.STORED_MAX_BYTES = 1 << 30
        small = GL.PrefilteredLight(_cube(256, 3, dev))
        assert not small.stored_operators
    finally:
        GL.PrefilteredLight.STORED_MAX_BYTES = old


@pytest.mark.parametrize("res", [2, 3, 16, 32, 100, 256, 512])
def test_cube_seam_table_matches_the_reference_fold(res):
    """The 24-entry edge table the shading kernels use for taps that step over a cube-face edge (shade_core.cuh:
    cube_texel) against the reference statement of the fold (cube_wrap_texel), on every tap position of a level."""
    from gigs import _lib
    L = _lib.load()
    bad = torch.full((1,), -1, dtype=torch.int32, device="cuda:0")
    _lib.check(L.gigs_cube_wrap_selfcheck(res, bad.data_ptr(), torch.cuda.current_stream().cuda_stream), "selfcheck")
    assert int(bad.item()) == 0
