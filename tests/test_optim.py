"""The optimiser step. CPU: the oracle's Adam restatement against torch.optim.Adam (the optimiser the
reference calls) and the committed golden trajectory; the lr schedule against values of the reference's own
get_expon_lr_func; the reference's update_learning_rate control flow. GPU: gigs_adam_step through the C-ABI against
torch.optim.Adam on the device, the oracle and the golden trajectory."""
import os

import numpy as np
import pytest
import torch

import gigs_oracle as O
from gigs import optim as gopt
from gigs import scene, step as gstep

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "optim_ref.npz")
# float32 Adam: our kernel and the oracle apply torch's operations in torch's order; what is left is whether a
# compiler contracted `a + b*c`. Tolerance = a few ulp of the parameter, plus a few ulp of the update itself (an Adam
# update is of the order of lr whatever the parameter's size): ATOL_PER_LR * lr.
RTOL, ATOL_PER_LR = 2e-6, 4e-6
ATOL = ATOL_PER_LR * 0.05


def test_expon_lr_matches_reference_values():
    z = np.load(GOLD)
    steps = z["lr_steps"].tolist()
    f = gopt.get_expon_lr_func(0.00016 * 5.2, 0.0000016 * 5.2, lr_delay_mult=0.01, max_steps=30000)
    f2 = gopt.get_expon_lr_func(0.05, 0.005, lr_delay_mult=0.01, max_steps=10000)
    f3 = gopt.get_expon_lr_func(0.01, 0.0001, lr_delay_steps=500, lr_delay_mult=0.01, max_steps=1000)
    for i, s in enumerate(steps):
        assert f(s) == pytest.approx(z["lr_xyz"][i], rel=1e-14, abs=0)
        assert f2(s) == pytest.approx(z["lr_brdf"][i], rel=1e-14, abs=0)
        assert f3(s) == pytest.approx(z["lr_delay"][i], rel=1e-14, abs=0)
        assert O.expon_lr(s, 0.01, 0.0001, 500, 0.01, 1000) == pytest.approx(z["lr_delay"][i], rel=1e-14, abs=0)


def test_oracle_adam_matches_golden_trajectory():
    z = np.load(GOLD)
    p, c = torch.from_numpy(z["p0"]), torch.from_numpy(z["c0"])
    m = v = mc = vc = None
    m, v, mc, vc = (torch.zeros_like(p) for _ in range(4))
    grads = torch.from_numpy(z["grads"])
    for t in range(grads.shape[0]):
        p, m, v = O.adam_step(p, grads[t], m, v, t + 1, 0.0025, eps=1e-15)
        c, mc, vc = O.adam_step(c, grads[t].flip(0), mc, vc, t + 1, 0.05, eps=1e-8, clamp_min0=True)
        assert torch.allclose(p, torch.from_numpy(z["traj_p"][t]), rtol=RTOL, atol=ATOL_PER_LR * 0.0025), t
        assert torch.allclose(c, torch.from_numpy(z["traj_c"][t]), rtol=RTOL, atol=ATOL), t
    gmax = grads.abs().amax(0)             # the moments are sums of terms of the gradients' size: compare at that scale
    assert float(((m - torch.from_numpy(z["m_p"])).abs() / gmax).max()) <= 1e-6
    assert float(((v - torch.from_numpy(z["v_p"])).abs() / gmax ** 2).max()) <= 1e-6
    assert float(c.min()) >= 0.0 and int((c == 0).sum()) > 0     # the clamp acted


def test_oracle_adam_matches_torch_adam_live():
    g = torch.Generator().manual_seed(3)
    p0 = torch.randn(1000, generator=g)
    p = p0.clone().requires_grad_(True)
    opt = torch.optim.Adam([p], lr=0.05, eps=1e-15)
    q, m, v = p0.clone(), torch.zeros(1000), torch.zeros(1000)
    for t in range(6):
        gr = torch.randn(1000, generator=g) * (0.0 if t in (2, 3) else 1e-3)
        p.grad = gr.clone()
        opt.step()
        q, m, v = O.adam_step(q, gr, m, v, t + 1, 0.05, eps=1e-15)
        assert torch.allclose(q, p.detach(), rtol=RTOL, atol=ATOL)


def test_update_learning_rate_follows_the_reference_control_flow():
    raw = scene.make_scene(8, seed=0)
    p = gstep.GaussianParams(raw, "cpu")
    o = gopt.GaussianOptimizer(p, spatial_lr_scale=5.2)
    names = [g["name"] for g in o.adam.param_groups]
    assert names == ["xyz", "f_dc", "f_rest", "opacity", "normal", "albedo", "roughness", "metallic", "scaling",
                     "rotation"]                                         # scene/gaussian_model.py:325-345
    assert o.adam.group("f_rest")["lr"] == 0.0025 / 20.0 and o.adam.group("xyz")["eps"] == 1e-15
    r = o.update_learning_rate(100)
    assert r == 0.0 and o.adam.group("albedo")["lr"] == 0.0              # BRDF_scheduler(iteration - 30000), step < 0
    assert o.adam.group("roughness")["lr"] == 0.05                       # never reached: the loop returns at albedo
    assert o.adam.group("xyz")["lr"] == pytest.approx(O.expon_lr(100, 0.00016 * 5.2, 0.0000016 * 5.2, 0, 0.01, 30000))
    r = o.update_learning_rate(35000)
    assert r == pytest.approx(O.expon_lr(5000, 0.05, 0.005, 0, 0.01, 10000)) and r == o.adam.group("albedo")["lr"]
    p2 = gstep.GaussianParams(raw, "cpu")
    p2.light_base = torch.rand(6, 16, 16, 3) * 0.5 + 0.25        # (the prefilter itself needs the device)
    o2 = gopt.GaussianOptimizer(p2)
    cg = o2.adam.group("cubemap")
    assert cg["eps"] == 1e-8 and cg["lr"] == 0.05 and cg["clamp_min0"]   # train.py:215-218, :523


def test_fused_adam_refuses_cpu_tensors():
    p = torch.zeros(8)
    fa = gopt.FusedAdam([dict(params=[p], name="x")])
    with pytest.raises(RuntimeError, match="CUDA"):
        fa.step(grads={"x": torch.zeros(8)})


def test_densify_stats_oracle_shapes():
    g = torch.Generator().manual_seed(0)
    P = 50
    radii = torch.randint(-1, 5, (P,), generator=g).int()
    grad = torch.randn(P, 3, generator=g)
    z = [torch.zeros(P, 1) for _ in range(4)] + [torch.zeros(P)]
    acc, acc_abs, acc_max, den, mr = O.densify_stats(radii, grad, *z)
    vis = radii > 0
    assert torch.equal(den[:, 0], vis.float()) and torch.all(acc[~vis] == 0)
    assert torch.allclose(acc_abs[vis, 0], grad[vis, 0].abs() + grad[vis, 1].abs())
    assert torch.equal(mr[vis], radii[vis].float())


# ------------------------------------------------------------------------------------------------ GPU
def _dev():
    return torch.device("cuda:0")


@pytest.mark.gpu
def test_fused_adam_matches_torch_adam_on_device():
    dev = _dev()
    g = torch.Generator().manual_seed(5)
    shapes = {"xyz": (4099, 3), "f_rest": (4099, 45), "opacity": (4099, 1), "rotation": (4099, 4), "cubemap": (6, 32, 32, 3)}
    lrs = {"xyz": 0.00016, "f_rest": 0.0025 / 20, "opacity": 0.05, "rotation": 0.001, "cubemap": 0.05}
    init = {k: (torch.randn(s, generator=g) if k != "cubemap" else torch.rand(s, generator=g) * 0.02) for k, s in shapes.items()}
    ours = {k: v.to(dev).clone() for k, v in init.items()}
    ref = {k: v.to(dev).clone().requires_grad_(True) for k, v in init.items()}
    fa = gopt.FusedAdam([dict(params=[ours[k]], lr=lrs[k], name=k, eps=(1e-8 if k == "cubemap" else 1e-15),
                              clamp_min0=(k == "cubemap")) for k in shapes])
    topt = torch.optim.Adam([dict(params=[ref[k]], lr=lrs[k], name=k) for k in shapes if k != "cubemap"], lr=0.0, eps=1e-15)
    lopt = torch.optim.Adam([dict(params=[ref["cubemap"]], lr=0.05, name="cubemap")], lr=0.05)
    worst = 0.0
    exact = True
    for t in range(10):
        grads = {k: (torch.randn(s, generator=g) * 10.0 ** float(torch.randint(-6, 1, (1,), generator=g))).to(dev)
                 for k, s in shapes.items()}
        zero = ["xyz", "rotation"] if t in (4, 5, 6) else []      # PBR-stage: known-zero gradients are not even read
        for k in zero:
            grads[k].zero_()
        mine = {k: v.clone() for k, v in grads.items()}
        for k in shapes:
            ref[k].grad = grads[k].clone()
        topt.step()
        lopt.step()
        with torch.no_grad():
            ref["cubemap"].clamp_(min=0.0)
        fa.step(zero_grads=zero, clear_grad=True, grads=mine)
        for k in shapes:
            if k not in zero:
                assert float(mine[k].abs().max()) == 0.0          # zero_grad fused into the pass
            a, b = ours[k], ref[k].detach()
            assert torch.allclose(a, b, rtol=RTOL, atol=ATOL_PER_LR * lrs[k]), (t, k, float((a - b).abs().max()))
            worst = max(worst, float(((a - b).abs() / (b.abs() + 1e-12)).max()))
            exact = exact and torch.equal(a, b)
            st = topt.state[ref[k]] if k != "cubemap" else lopt.state[ref[k]]
            gs = float(st["exp_avg"].abs().max()) + 1e-30
            assert float((fa.state[k]["exp_avg"] - st["exp_avg"]).abs().max()) <= 1e-6 * gs
            assert float((fa.state[k]["exp_avg_sq"] - st["exp_avg_sq"]).abs().max()) <= 1e-6 * float(st["exp_avg_sq"].max())
    assert float(ours["cubemap"].min()) >= 0.0 and int((ours["cubemap"] == 0).sum()) > 0
    print(f"fused Adam vs torch.optim.Adam (CUDA): worst relative difference {worst:.3g}, bit-identical: {exact}")


@pytest.mark.gpu
def test_fused_adam_matches_golden_trajectory_and_unaligned_tensors():
    dev = _dev()
    z = np.load(GOLD)
    # views at odd offsets: the scalar path
    buf = torch.zeros(4 * 1537 + 8, device=dev)
    p = buf[1:1538]
    p.copy_(torch.from_numpy(z["p0"]))
    c = buf[1539:3076]
    c.copy_(torch.from_numpy(z["c0"]))
    fa = gopt.FusedAdam([dict(params=[p], lr=0.0025, name="f_dc", eps=1e-15),
                         dict(params=[c], lr=0.05, name="cubemap", eps=1e-8, clamp_min0=True)])
    grads = torch.from_numpy(z["grads"]).to(dev)
    for t in range(grads.shape[0]):
        fa.step(grads={"f_dc": grads[t].clone(), "cubemap": grads[t].flip(0).contiguous()})
        assert torch.allclose(p.cpu(), torch.from_numpy(z["traj_p"][t]), rtol=RTOL, atol=ATOL_PER_LR * 0.0025), t
        assert torch.allclose(c.cpu(), torch.from_numpy(z["traj_c"][t]), rtol=RTOL, atol=ATOL), t
    assert float(buf[0]) == 0.0 and float(buf[1538]) == 0.0      # neighbours untouched
    sd = fa.state_dict()
    assert float(sd["state"][0]["step"]) == 12.0 and sd["param_groups"][1]["name"] == "cubemap"


@pytest.mark.gpu
def test_adam_argument_errors_and_empty_groups():
    from gigs import _lib
    L = _lib.load()
    assert L.gigs_adam_step(0, None, None) == 0
    assert L.gigs_adam_step(25, None, None) < 0
    arr = (_lib.GigsAdamGroup * 1)()
    arr[0].count = 10
    arr[0].step = 1
    assert L.gigs_adam_step(1, arr, None) < 0 and b"NULL" in L.gigs_last_error()
    e = torch.zeros(0, device=_dev())
    gopt.FusedAdam([dict(params=[e], name="empty")]).step(grads={"empty": e})     # P == 0: no launch, no error


@pytest.mark.gpu
def test_densify_stats_matches_oracle():
    from gigs import _lib
    dev = _dev()
    g = torch.Generator().manual_seed(9)
    P = 70001
    radii = torch.randint(-2, 40, (P,), generator=g).int()
    grad = torch.randn(P, 3, generator=g) * 1e-3
    state = [torch.rand(P, 1, generator=g) for _ in range(4)] + [torch.rand(P, generator=g) * 20]
    want = O.densify_stats(radii, grad, *state)
    d = [t.to(dev).clone() for t in state]
    L = _lib.load()
    _lib.check(L.gigs_densify_stats(P, radii.to(dev).data_ptr(), grad.to(dev).data_ptr(), 3, d[0].data_ptr(),
                                    d[1].data_ptr(), d[2].data_ptr(), d[3].data_ptr(), d[4].data_ptr(),
                                    torch.cuda.current_stream().cuda_stream), "gigs_densify_stats")
    for a, b in zip(d, want):
        assert torch.allclose(a.cpu(), b, rtol=1e-6, atol=1e-9)
    assert torch.equal(d[3].cpu(), want[3]) and torch.equal(d[4].cpu(), want[4])


@pytest.mark.gpu
def test_training_iterations_with_fused_optimizer_match_operator_path_with_torch_adam():
    """Three full PBR-stage iterations (frame + optimiser step) on the fused path with FusedAdam against the operator
    path stepped by torch.optim.Adam with the reference's groups: parameters agree after every iteration."""
    from gigs import shade
    dev = _dev()
    P, W, H = 3000, 160, 120
    raw = scene.make_scene(P, seed=4, regime="trained")
    cam = scene.orbit_camera(1, 8, W, H).to(dev)
    lut = shade.make_brdf_lut(64, 64).to(dev)
    rays = scene.canonical_rays(cam, dev)
    gt = torch.rand(3, H, W, generator=torch.Generator().manual_seed(5)).to(dev)
    bg = torch.zeros(3, device=dev)
    gi = dict(radius=0.8, bias=0.01, thick=0.05, delta=0.0625, step=16, start=64)
    base = torch.rand(6, 64, 64, 3, generator=torch.Generator().manual_seed(21)) * 0.5 + 0.25
    pa = gstep.GaussianParams(raw, dev, light_base=base.clone())
    pb = gstep.GaussianParams(raw, dev, light_base=base.clone())

    def run(p, fused):
        return gstep.training_step(p, cam, p.light(), lut, rays, gt, bg, gi, fused=fused, brdf_tv_weight=1.0,
                                   env_tv_weight=0.01)
    oa = gopt.GaussianOptimizer(pa)
    groups = [dict(params=[pb.leaves[k]], lr=oa.adam.group(gopt.REFERENCE_GROUP_NAME[k])["lr"],
                   name=gopt.REFERENCE_GROUP_NAME[k]) for k in gstep.PARAM_KEYS]
    tb = torch.optim.Adam(groups, lr=0.0, eps=1e-15)
    lb = torch.optim.Adam([dict(params=[pb.light_base], lr=0.05, name="cubemap")], lr=0.05)
    pa.zero_grad()
    pb.zero_grad()
    for it in range(3):
        la = run(pa, True)
        lb_ = run(pb, False)
        assert float(la) == pytest.approx(float(lb_), rel=2e-4)
        oa.step()
        tb.step()
        lb.step()
        with torch.no_grad():
            pb.light_base.clamp_(min=0.0)
        pb.zero_grad()
        assert float(pa.flat_grad.abs().max()) == 0.0
        for k in gstep.PARAM_KEYS:
            a, b = pa.leaves[k].detach(), pb.leaves[k].detach()
            # Adam normalises the step to ~lr whatever the gradient's size, so rounding-level gradient differences
            # between the two paths move a parameter by a fraction of lr at most
            lr = oa.adam.group(gopt.REFERENCE_GROUP_NAME[k])["lr"]
            assert float((a - b).abs().max()) <= 0.05 * lr * (it + 1) + 1e-7, (it, k, float((a - b).abs().max()))
        assert float((pa.light_base - pb.light_base).abs().max()) <= 0.05 * 0.05 * (it + 1)
