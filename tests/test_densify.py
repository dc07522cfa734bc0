"""Densification / pruning. The goldens (tests/golden/densify_ref.npz) are produced by the reference's own
GaussianModel methods compiled from its source (tests/make_golden_densify.py). CPU: the oracle's round-by-round
restatement against them. GPU: gigs.densify (one source map + one gigs_densify_gather launch) against the goldens:
statistics, row order, parameters, Adam moments, step counts, reset_opacity."""
import os

import numpy as np
import pytest
import torch

import gigs_oracle as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "densify_ref.npz")
KEYS = O.DENSIFY_KEYS


def _case(z, i):
    c = {k[len(f"c{i}_"):]: z[k] for k in z.files if k.startswith(f"c{i}_")}
    P, seed, extent, max_grad, mss, n_clone, n_child = c["meta"].tolist()
    return c, int(P), float(extent), float(max_grad), (int(mss) or None)


def _close(a, b, what):
    a, b = torch.as_tensor(a).float(), torch.as_tensor(b).float()
    assert a.shape == b.shape, (what, a.shape, b.shape)
    # copied values are exact; sampled positions / split scales go through exp, a 3x3 product and log
    assert torch.allclose(a, b, rtol=2e-6, atol=2e-6), (what, float((a - b).abs().max()))


@pytest.mark.parametrize("i", [0, 1])
def test_oracle_densify_matches_reference_methods(i):
    z = np.load(GOLD)
    c, P, extent, max_grad, mss = _case(z, i)
    p = {k: torch.from_numpy(c[f"p_{k}"]) for k in KEYS}
    m = {k: torch.from_numpy(c[f"m_{k}"]) for k in KEYS}
    v = {k: torch.from_numpy(c[f"v_{k}"]) for k in KEYS}
    # statistics first (add_densification_stats + max_radii2D)
    st = [torch.zeros(P, 1) for _ in range(4)] + [torch.zeros(P)]
    for r, g in zip(c["radii"], c["grad2D"]):
        st = list(O.densify_stats(torch.from_numpy(r), torch.from_numpy(g), *st))
    for got, name in zip(st, ("accum", "accum_abs", "accum_abs_max", "denom", "max_radii2D")):
        _close(got, c[name], name)
    q, qm, qv = O.densify_and_prune(p, m, v, st[0], st[1], st[3], max_grad, 0.05, extent, mss,
                                    torch.from_numpy(c["noise_clone"]), torch.from_numpy(c["noise_split"]))
    for k in KEYS:
        _close(q[k], c[f"q_{k}"], k)
        _close(qm[k], c[f"qm_{k}"], "exp_avg " + k)
        _close(qv[k], c[f"qv_{k}"], "exp_avg_sq " + k)
    assert q["xyz"].shape[0] != P                      # the case really changes the model


@pytest.mark.gpu
@pytest.mark.parametrize("i", [0, 1])
def test_fused_densify_matches_reference_methods(i):
    from gigs import densify, optim as gopt, step as gstep
    dev = torch.device("cuda:0")
    z = np.load(GOLD)
    c, P, extent, max_grad, mss = _case(z, i)
    raw = {k: torch.from_numpy(c[f"p_{k}"]) for k in KEYS}
    raw["sh_degree"] = 3
    params = gstep.GaussianParams(raw, dev)
    opt = gopt.GaussianOptimizer(params)
    for k in KEYS:
        opt.adam.state[gopt.REFERENCE_GROUP_NAME[k]] = dict(step=2, exp_avg=torch.from_numpy(c[f"m_{k}"]).to(dev),
                                                            exp_avg_sq=torch.from_numpy(c[f"v_{k}"]).to(dev))
    st = densify.DensifyState(P, dev)
    for r, g in zip(c["radii"], c["grad2D"]):
        st.add_view(torch.from_numpy(g).to(dev), torch.from_numpy(r).to(dev))
    for got, name in ((st.xyz_gradient_accum, "accum"), (st.xyz_gradient_accum_abs, "accum_abs"),
                      (st.xyz_gradient_accum_abs_max, "accum_abs_max"), (st.denom, "denom"), (st.max_radii2D, "max_radii2D")):
        _close(got.cpu(), c[name], name)
    pl = densify.plan(params, st, max_grad, 0.05, extent, mss)
    nc, ns2 = c["noise_clone"].shape[0], c["noise_split"].shape[0]
    assert pl["n_clone"] == nc and 2 * pl["n_split"] == ns2
    noise = torch.cat([torch.zeros(pl["n_candidates"] - nc - ns2, 3), torch.from_numpy(c["noise_clone"]),
                       torch.from_numpy(c["noise_split"])])
    info = densify.densify_and_prune(params, opt, st, max_grad, 0.05, extent, mss, noise=noise)
    assert info["P_after"] == c["q_xyz"].shape[0] and params.P == info["P_after"]
    for k in KEYS:
        _close(params.leaves[k].detach().cpu(), c[f"q_{k}"], k)
        s = opt.adam.state[gopt.REFERENCE_GROUP_NAME[k]]
        _close(s["exp_avg"].cpu(), c[f"qm_{k}"], "exp_avg " + k)
        _close(s["exp_avg_sq"].cpu(), c[f"qv_{k}"], "exp_avg_sq " + k)
        assert s["step"] == int(c[f"step_{k}"]) == 2
        assert opt.adam.group(gopt.REFERENCE_GROUP_NAME[k])["params"][0] is params.leaves[k]
        assert params.leaves[k].grad.data_ptr() == params.flat_grad[params._span[k][0]:].data_ptr()
    assert float(st.xyz_gradient_accum.abs().sum()) == 0.0 and st.max_radii2D.shape[0] == params.P
    # copied rows are bit-exact (only sampled positions / split scales are computed)
    kept = pl["kind"].cpu() == 0
    assert torch.equal(params.leaves["f_rest"].detach().cpu()[kept], torch.from_numpy(c["q_f_rest"])[kept])
    densify.reset_opacity(params, opt)
    _close(params.leaves["opacity"].detach().cpu(), c["reset_opacity"], "reset_opacity")
    assert float(opt.adam.state["opacity"]["exp_avg"].abs().sum()) == 0.0
    # the rebuilt model trains: one optimiser step on the new tensors
    for k in KEYS:
        params.leaves[k].grad.normal_()
    params.mark_dirty(None)
    opt.step()
    assert opt.adam.state["xyz"]["step"] == 3 and float(params.flat_grad.abs().max()) == 0.0


@pytest.mark.gpu
def test_densify_gather_argument_errors_and_default_noise():
    from gigs import _lib, densify, optim as gopt, scene, step as gstep
    dev = torch.device("cuda:0")
    L = _lib.load()
    assert L.gigs_densify_gather(0, None, None, None, None, None, 1.6, 0, None, None) == 0
    assert L.gigs_densify_gather(5, None, None, None, None, None, 1.6, 17, None, None) < 0
    arr = (_lib.GigsDensifyGroup * 1)()
    idx = torch.zeros(5, dtype=torch.int32, device=dev)
    kind = torch.zeros(5, dtype=torch.int8, device=dev)
    assert L.gigs_densify_gather(5, idx.data_ptr(), kind.data_ptr(), None, None, None, 1.6, 1, arr, None) < 0
    assert b"NULL tensor" in L.gigs_last_error()
    # device-drawn noise: new points lie within a few sigma of their sources; no optimiser attached
    raw = scene.make_scene(3000, seed=3, regime="trained")
    params = gstep.GaussianParams(raw, dev)
    st = densify.DensifyState(3000, dev)
    g = torch.Generator().manual_seed(0)
    st.add_view((torch.randn(3000, 3, generator=g) * 1e-3).to(dev), torch.ones(3000, dtype=torch.int32, device=dev))
    before = params.leaves["xyz"].detach().clone()
    info = densify.densify_and_prune(params, None, st, 0.0002, 0.005, 5.0, None)
    assert info["P_after"] > 3000 and info["n_clone"] + info["n_split"] > 0
    assert torch.isfinite(params.leaves["xyz"]).all() and float(params.leaves["xyz"].abs().max()) < float(before.abs().max()) + 2.0


@pytest.mark.gpu
def test_densify_with_nothing_selected_only_prunes():
    """Thresholds nobody reaches: no clones, no splits; the pass degenerates to the opacity prune, moments of the kept
    rows are carried over bit for bit."""
    from gigs import densify, optim as gopt, scene, step as gstep
    dev = torch.device("cuda:0")
    raw = scene.make_scene(2000, seed=5, regime="trained")
    params = gstep.GaussianParams(raw, dev)
    opt = gopt.GaussianOptimizer(params)
    for k in gstep.PARAM_KEYS:
        params.leaves[k].grad.normal_()
    params.mark_dirty(None)
    opt.step()
    m_before = opt.adam.state["f_rest"]["exp_avg"].clone()
    op = torch.sigmoid(params.leaves["opacity"].detach())[:, 0]
    st = densify.DensifyState(2000, dev)
    g = torch.rand(2000, 3, generator=torch.Generator().manual_seed(1)) * 1e-9      # distinct values: no ties at the maximum
    st.add_view(g.to(dev), torch.ones(2000, dtype=torch.int32, device=dev))
    info = densify.densify_and_prune(params, opt, st, 1e3, 0.05, 5.0, None)
    keep = op >= 0.05
    # grads_abs >= Q with Q = quantile(., 1 - 0) = max: the reference's rule still selects the maximum element(s)
    assert info["n_clone"] + info["n_split"] <= 2
    assert abs(info["P_after"] - int(keep.sum())) <= 4
    if info["n_clone"] + info["n_split"] == 0:
        assert torch.equal(opt.adam.state["f_rest"]["exp_avg"], m_before[keep])
