"""GPU: the training-loop glue (gigs.train.Trainer.iteration = train.py:246-523) across both stages on a small scene:
first stage with densification, pruning and an opacity reset, the switch to the PBR stage with the trainable base
cubemap, learning-rate schedule and the fused optimiser throughout. Checks the schedule's side effects and that the
loss of a fixed view goes down in each stage."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from gigs import scene, shade, train as gtrain
from gigs import step as gstep
from gigs.optim import OptimizationParams

DEV = "cuda:0"


def test_trainer_runs_both_stages_with_densification():
    P, W, H = 4000, 200, 160
    raw = scene.make_scene(P, seed=6, regime="trained")
    base = torch.rand(6, 64, 64, 3, generator=torch.Generator().manual_seed(1)) * 0.5 + 0.25
    params = gstep.GaussianParams(raw, DEV, light_base=base)
    cams = [scene.orbit_camera(k, 4, W, H).to(DEV) for k in range(4)]
    g = torch.Generator().manual_seed(2)
    gts = [torch.rand(3, 1, 1, generator=g).expand(3, H, W).contiguous().to(DEV) * 0.5 + 0.25 for _ in range(4)]
    opt = OptimizationParams(iterations=60, densify_from_iter=5, densification_interval=10, opacity_reset_interval=25,
                             densify_until_iter=28, densify_grad_threshold=0.00002)
    # (an opacity reset shortly BEFORE a pruning pass would prune everything: the reference leaves 100 iterations between)
    cfg = gtrain.TrainConfig(pbr_iteration=36, opt=opt)
    lut = shade.make_brdf_lut(64, 64).to(DEV)
    tr = gtrain.Trainer(params, lut, cameras_extent=4.0, cfg=cfg, rays_of=lambda c: scene.canonical_rays(c, DEV))
    base0 = params.light_base.detach().clone()
    alb0 = params.leaves["albedo"].detach().clone()
    for it in range(1, 56):
        k = it % 4
        loss = tr.iteration(it, cams[k], gts[k])
        assert torch.isfinite(loss), it
    log = tr.log
    dens = [e for e in log if e["event"] and "P_after" in e["event"]]
    assert [e["iteration"] for e in dens] == [10, 20]                          # > densify_from_iter, every 10, < 28
    assert any(e["event"]["n_clone"] + e["event"]["n_split"] > 0 for e in dens)
    assert [e["iteration"] for e in log if e["event"] and e["event"].get("reset_opacity")] == [25]
    assert log[35]["stage"] == 1 and log[36]["stage"] == 2
    assert params.P == log[-1]["P"] == params.leaves["xyz"].shape[0] == tr.stats.denom.shape[0]
    # the light only trains from pbr_iteration on (train.py:520); materials only in the PBR stage
    assert not torch.equal(params.light_base.detach(), base0) and float(params.light_base.min()) >= 0.0
    # light_optimizer.step() runs from it == pbr_iteration (train.py:520), but that iteration is still first-stage: the
    # light has no gradient (None) and torch's Adam skips it, so its first counted step is iteration 37
    assert tr.optimizer.adam.state["cubemap"]["step"] == 55 - 36
    assert tr.optimizer.adam.group("albedo")["lr"] == 0.0                       # the reference's schedule: 0 before 30k
    assert tr.optimizer.adam.group("xyz")["lr"] < 0.00016
    # losses go down within each stage (same four views cycled); first stage: before the first densification / opacity
    # reset disturbs the model
    s1 = [e["loss"] for e in log if e["stage"] == 1 and e["iteration"] < 10]
    s2 = [e["loss"] for e in log if e["stage"] == 2]
    assert sum(s1[-4:]) < sum(s1[:4]) and sum(s2[-4:]) < sum(s2[:4])
