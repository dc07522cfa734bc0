"""CPU: host-side mirror of the reference interface — names, argument validation, error behaviour."""
import pytest
import torch

import diff_gaussian_rasterization as dgr
from gigs import scene, step as gstep


def settings(**kw):
    base = dict(image_height=32, image_width=32, tanfovx=0.5, tanfovy=0.5, radius=0.8, bias=0.01, thick=0.05,
                delta=0.0625, step=16, start=8, bg=torch.zeros(3), scale_modifier=1.0, viewmatrix=torch.eye(4),
                projmatrix=torch.eye(4), sh_degree=3, campos=torch.zeros(3), prefiltered=False, debug=False,
                inference=False, argmax_depth=False)
    base.update(kw)
    return dgr.GaussianRasterizationSettings(**base)


def test_settings_fields_match_reference_order():
    # diff_gaussian_rasterization/__init__.py:31-51
    assert dgr.GaussianRasterizationSettings._fields == (
        "image_height", "image_width", "tanfovx", "tanfovy", "radius", "bias", "thick", "delta", "step", "start", "bg",
        "scale_modifier", "viewmatrix", "projmatrix", "sh_degree", "campos", "prefiltered", "debug", "inference",
        "argmax_depth")


def test_public_surface():
    for name in ("GaussianRasterizationSettings", "GaussianRasterizer", "Gaussian_SSR", "_C", "_RasterizeGaussians",
                 "_SSR"):
        assert hasattr(dgr, name)
    for fn in ("depth_to_normal", "SSAO", "SSR", "SSR_BACKWARD", "rasterize_gaussians", "lite_rasterize_gaussians",
               "rasterize_gaussians_backward", "mark_visible"):          # ext.cpp:16-24
        assert callable(getattr(dgr._C, fn))
    from simple_knn._C import distCUDA2
    assert callable(distCUDA2)


def test_sh_colour_and_cov_ambiguity_raise_like_the_reference():
    r = dgr.GaussianRasterizer(settings())
    P = 4
    a = dict(means3D=torch.zeros(P, 3), means2D=torch.zeros(P, 3), opacities=torch.zeros(P, 1),
             normal=torch.zeros(P, 3), albedo=torch.zeros(P, 3), roughness=torch.zeros(P, 1),
             metallic=torch.zeros(P, 1))
    with pytest.raises(Exception, match="Please provide excatly one of either SHs or precomputed colors!"):
        r(**a, shs=None, colors_precomp=None, scales=torch.ones(P, 3), rotations=torch.ones(P, 4))
    with pytest.raises(Exception, match="Please provide excatly one of either SHs or precomputed colors!"):
        r(**a, shs=torch.zeros(P, 16, 3), colors_precomp=torch.zeros(P, 3), scales=torch.ones(P, 3),
          rotations=torch.ones(P, 4))
    with pytest.raises(Exception, match="exactly one of either scale/rotation pair or precomputed 3D covariance"):
        r(**a, shs=torch.zeros(P, 16, 3), scales=None, rotations=None, cov3D_precomp=None)
    with pytest.raises(Exception, match="exactly one of either scale/rotation pair or precomputed 3D covariance"):
        r(**a, shs=torch.zeros(P, 16, 3), scales=torch.ones(P, 3), rotations=torch.ones(P, 4),
          cov3D_precomp=torch.zeros(P, 6))


def test_cpu_tensors_fail_loudly_not_silently():
    E = torch.Tensor([])
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        dgr._C.rasterize_gaussians(torch.zeros(3), torch.zeros(4, 3), E, torch.zeros(4, 1), torch.zeros(4, 3),
                                   torch.zeros(4, 3), torch.zeros(4, 1), torch.zeros(4, 1), torch.ones(4, 3),
                                   torch.ones(4, 4), E, torch.zeros(4, 16, 3), torch.zeros(3), torch.eye(4),
                                   torch.eye(4), 1.0, 0.5, 0.5, 32, 32, 3, False, False, False, False)
    with pytest.raises(RuntimeError, match="dimensions"):
        dgr._C.rasterize_gaussians(torch.zeros(3), torch.zeros(4, 2), E, E, E, E, E, E, E, E, E, E, torch.zeros(3),
                                   torch.eye(4), torch.eye(4), 1.0, 0.5, 0.5, 32, 32, 3, False, False, False, False)
    from simple_knn._C import distCUDA2
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        distCUDA2(torch.zeros(5, 3))


def test_view_sharding_is_a_partition():
    for world in (1, 2, 4, 8):
        seen = sorted(v for r in range(world) for v in gstep.shard_views(200, r, world))
        assert seen == list(range(200))
        sizes = [len(gstep.shard_views(200, r, world)) for r in range(world)]
        assert max(sizes) - min(sizes) <= 1


def test_flat_gradient_buffer_receives_autograd_in_place():
    raw = scene.make_scene(10, seed=0)
    p = gstep.GaussianParams(raw, "cpu", light=scene.make_light(0, base_res=32))
    assert p.flat_grad.numel() == 10 * 67 + sum(6 * r * r * 3 for r in (16, 32, 16))
    loss = sum((t * (i + 1)).sum() for i, t in enumerate(p.leaves.values()))
    loss.backward()
    o = 0
    for i, (k, t) in enumerate(p.leaves.items()):
        n = t.numel()
        assert torch.all(p.flat_grad[o:o + n] == i + 1), k      # .grad is a view into the flat buffer
        assert t.grad.data_ptr() == p.flat_grad[o:o + n].data_ptr()
        o += n
    assert gstep.PARAM_WIDTH and sum(gstep.PARAM_WIDTH.values()) == 67   # 268 B / Gaussian
    p.zero_grad()
    assert float(p.flat_grad.abs().sum()) == 0.0
