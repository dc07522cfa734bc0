"""GPU: one view of the FIRST training stage (gigs.step.first_stage_step: render() + (1-lambda) L1 + lambda (1-SSIM) +
normal L1 + normal TV, backward through the general rasterizer backward) with the fused loss kernels against the same
step with the loss written in framework ops exactly as the reference writes it (train.py:318-328,
utils/loss_utils.py:54-100). The rasterizer underneath is pinned to the reference's kernels by test_gpu_parity.py."""
import pytest
import torch

pytestmark = pytest.mark.gpu

import gpu_util as U
from gigs import scene, step as gstep

DEV = "cuda:0"
GI = dict(radius=0.8, bias=0.01, thick=0.05, delta=0.0625, step=16, start=64)


@pytest.mark.parametrize("P,W,H", [(20000, 400, 300), (5000, 333, 257)])
def test_first_stage_step_fused_losses_match_framework_ops(P, W, H):
    raw = scene.make_scene(P, seed=2, regime="trained")
    cam = scene.orbit_camera(1, 8, W, H).to(DEV)
    gt = torch.rand(3, H, W, generator=torch.Generator().manual_seed(5)).to(DEV)
    bg = torch.ones(3, device=DEV)
    out = {}
    for fused in (True, False):
        p = gstep.GaussianParams(raw, DEV)
        p.zero_grad()
        loss, res = gstep.first_stage_step(p, cam, gt, bg, GI, fused_losses=fused)
        torch.cuda.synchronize()
        out[fused] = (float(loss), p.flat_grad.clone(), res["viewspace_points"].grad.clone(), p)
    la, lb = out[True][0], out[False][0]
    assert la == pytest.approx(lb, rel=1e-5)
    pa = out[True][3]
    for k in gstep.PARAM_KEYS:
        lo, hi = pa._span[k]
        if k not in ("albedo", "roughness", "metallic"):                 # the first-stage loss reads no material map
            assert float(out[False][1][lo:hi].abs().max()) > 0, k
        U.assert_grad_close(out[True][1][lo:hi], out[False][1][lo:hi], k, 1e-3)
    U.assert_grad_close(out[True][2], out[False][2], "viewspace_points", 1e-3)
