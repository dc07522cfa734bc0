"""latlong_to_cubemap (relight.py:92-112): the oracle restatement against an independent formulation of the same
lookup (torch grid_sample, valid away from the wrap seam) and against analytic properties on the CPU; the CUDA kernel
(gigs_latlong_to_cubemap) against the oracle on the GPU. nvdiffrast is absent: parity unpinned, see the oracle."""
import math

import pytest
import torch

import gigs_oracle as O


def test_oracle_constant_and_linear_maps():
    env = torch.full((64, 128, 3), 0.37)
    cube = O.latlong_to_cubemap(env, 16)
    assert cube.shape == (6, 16, 16, 3) and float((cube - 0.37).abs().max()) < 1e-6
    # a map that is linear in the row coordinate reproduces tv = acos(y) / pi (bilinear is exact for linear data)
    EH, EW = 256, 512
    rows = ((torch.arange(EH, dtype=torch.float32) + 0.5) / EH)[:, None, None].expand(EH, EW, 1).contiguous()
    cube = O.latlong_to_cubemap(rows, 32)
    lin = torch.linspace(-1.0 + 1.0 / 32, 1.0 - 1.0 / 32, 32)
    gy, gx = torch.meshgrid(lin, lin, indexing="ij")
    v = torch.nn.functional.normalize(torch.stack((gx, torch.ones_like(gx), gy), -1), dim=-1)   # face 2: +y
    tv = torch.acos(v[..., 1]) / math.pi
    ok = (tv * EH > 1.0) & (tv * EH < EH - 1.0)
    assert float((cube[2, ..., 0] - tv)[ok].abs().max()) < 1e-5


def test_oracle_matches_grid_sample_away_from_the_seam():
    g = torch.Generator().manual_seed(0)
    EH, EW, R = 128, 256, 24
    env = torch.rand(EH, EW, 3, generator=g)
    cube = O.latlong_to_cubemap(env, R)
    lin = torch.linspace(-1.0 + 1.0 / R, 1.0 - 1.0 / R, R)
    gy, gx = torch.meshgrid(lin, lin, indexing="ij")
    one = torch.ones_like(gx)
    for s, d in enumerate([(one, -gy, -gx), (-one, -gy, gx), (gx, one, gy), (gx, -one, -gy), (gx, -gy, one),
                           (-gx, -gy, -one)]):
        v = torch.nn.functional.normalize(torch.stack(d, -1), dim=-1)
        tu = torch.atan2(v[..., 0], -v[..., 2]) / (2 * math.pi) + 0.5
        tv = torch.acos(torch.clamp(v[..., 1], -1, 1)) / math.pi
        grid = torch.stack((tu * 2 - 1, tv * 2 - 1), -1)[None]
        ref = torch.nn.functional.grid_sample(env.permute(2, 0, 1)[None], grid, mode="bilinear", padding_mode="border",
                                              align_corners=False)[0].permute(1, 2, 0)
        inner = ((tu * EW > 1.0) & (tu * EW < EW - 1.0) & (tv * EH > 1.0) & (tv * EH < EH - 1.0))
        assert float((cube[s] - ref)[inner].abs().max()) < 1e-5, s


@pytest.mark.gpu
@pytest.mark.parametrize("EH,EW,R", [(1024, 2048, 256), (64, 128, 16), (100, 333, 512)])
def test_kernel_matches_oracle(EH, EW, R):
    from gigs import light
    # a smooth HDR-like map: the lookup's (u, v) come out of atan2 / acos, whose last bits differ between the CPU and
    # CUDA math libraries; on smooth data that moves a texel value by ~1e-6
    # periodic in u (the lookup wraps), texel centres at (i + 0.5) / size
    yy, xx = torch.meshgrid((torch.arange(EH) + 0.5) / EH, (torch.arange(EW) + 0.5) / EW, indexing="ij")
    env = torch.stack([2.0 + torch.sin(2 * math.pi * xx) * torch.cos(3.0 * yy), 1.0 + yy * torch.cos(2 * math.pi * xx),
                       3.0 * torch.exp(-4.0 * (yy - 0.3) ** 2)],
                      -1).contiguous()
    want = O.latlong_to_cubemap(env, R)
    got = light.latlong_to_cubemap(env.cuda(), [R, R]).cpu()
    assert float((got - want).abs().max()) < 5e-5
    # white noise up to 4.0 (neighbouring texels unrelated): an error of 3e-7 in u is 6e-4 of a texel at 2048 columns,
    # times the difference of two neighbours
    g = torch.Generator().manual_seed(EH)
    noise = torch.rand(EH, EW, 3, generator=g) * 4.0
    d = (light.latlong_to_cubemap(noise.cuda(), [R, R]).cpu() - O.latlong_to_cubemap(noise, R)).abs()
    assert float(d.max()) < 4.0 * EW * 2e-6 + 1e-5, float(d.max())
