"""Interchange formats: the capture()/restore() tuple loads into torch.optim.Adam exactly as the
reference's GaussianModel.restore does; the PLY layout follows save_ply / load_ply. CPU only (host-side I/O)."""
import os

import numpy as np
import torch

from gigs import checkpoint as ck
from gigs import optim as gopt
from gigs import scene, step as gstep


class _State:          # DensifyState needs the device library only for add_view; the fields are plain tensors
    def __init__(self, P):
        self.max_radii2D = torch.arange(P, dtype=torch.float32)
        self.xyz_gradient_accum = torch.rand(P, 1)
        self.xyz_gradient_accum_abs = torch.rand(P, 1)
        self.xyz_gradient_accum_abs_max = torch.rand(P, 1)
        self.denom = torch.ones(P, 1)


def _model(P=37):
    raw = scene.make_scene(P, seed=4, regime="trained")
    p = gstep.GaussianParams(raw, "cpu")
    o = gopt.GaussianOptimizer(p, spatial_lr_scale=3.0)
    for i, g in enumerate(o.adam.param_groups):
        t = g["params"][0]
        o.adam.state[g["name"]] = dict(step=7, exp_avg=torch.full_like(t, 0.1 * (i + 1)), exp_avg_sq=torch.full_like(t, 0.01 * (i + 1)))
    return p, o, _State(P)


def test_capture_tuple_has_the_reference_layout_and_loads_into_torch_adam(tmp_path):
    p, o, st = _model()
    tup = ck.capture(p, o, st, spatial_lr_scale=3.0)
    assert len(tup) == 18 and tup[0] == 3 and tup[17] == 3.0
    assert tup[1] is p.leaves["xyz"] and tup[4] is p.leaves["log_scale"] and tup[5] is p.leaves["rot"]
    assert tup[2].shape == (37, 1, 3) and tup[3].shape == (37, 15, 3) and tup[11].shape == (37,)
    path = str(tmp_path / "chkpnt7.pth")
    torch.save({"gaussians": tup, "iteration": 7}, path)          # train.py:466-477
    back = torch.load(path, weights_only=False)["gaussians"]
    # the reference's restore(): training_setup builds torch.optim.Adam over the 10 groups, then load_state_dict
    order = ("xyz", "f_dc", "f_rest", "opacity", "normal", "albedo", "roughness", "metallic", "log_scale", "rot")
    tensors = dict(zip(("xyz", "f_dc", "f_rest", "log_scale", "rot", "opacity", "normal", "albedo", "roughness", "metallic"), back[1:11]))
    leaves = [torch.nn.Parameter(tensors[k].clone()) for k in order]
    ref = torch.optim.Adam([dict(params=[t], lr=0.0, name=gopt.REFERENCE_GROUP_NAME[k]) for k, t in zip(order, leaves)],
                           lr=0.0, eps=1e-15)
    ref.load_state_dict(back[16])
    assert [g["name"] for g in ref.param_groups] == [gopt.REFERENCE_GROUP_NAME[k] for k in order]
    assert ref.param_groups[0]["lr"] == 0.00016 * 3.0 and ref.param_groups[0]["eps"] == 1e-15
    assert float(ref.state[leaves[2]]["exp_avg"][0, 0, 0]) == np.float32(0.3) and float(ref.state[leaves[2]]["step"]) == 7.0


def test_restore_round_trip_and_reference_written_state(tmp_path):
    p, o, st = _model()
    tup = ck.capture(p, o, st, 3.0)
    p2, o2, st2, s = ck.restore(tup, "cpu", optimizer_factory=lambda pp, sl: gopt.GaussianOptimizer(pp, spatial_lr_scale=sl))
    assert s == 3.0 and p2.P == 37 and p2.sh_degree == 3
    for k in gstep.PARAM_KEYS:
        assert torch.equal(p2.leaves[k], p.leaves[k])
        a, b = o2.adam.state[gopt.REFERENCE_GROUP_NAME[k]], o.adam.state[gopt.REFERENCE_GROUP_NAME[k]]
        assert a["step"] == 7 and torch.equal(a["exp_avg"], b["exp_avg"]) and torch.equal(a["exp_avg_sq"], b["exp_avg_sq"])
    assert torch.equal(st2.max_radii2D, st.max_radii2D) and torch.equal(st2.denom, st.denom)
    # a state_dict produced by torch.optim.Adam itself (what a reference-written checkpoint holds)
    order = ("xyz", "f_dc", "f_rest", "opacity", "normal", "albedo", "roughness", "metallic", "log_scale", "rot")
    leaves = [torch.nn.Parameter(p.leaves[k].detach().clone()) for k in order]
    ref = torch.optim.Adam([dict(params=[t], lr=0.01 * (i + 1), name=gopt.REFERENCE_GROUP_NAME[k])
                            for i, (k, t) in enumerate(zip(order, leaves))], lr=0.0, eps=1e-15)
    for t in leaves:
        t.grad = torch.ones_like(t)
    ref.step()
    tup2 = tup[:16] + (ref.state_dict(), 3.0)
    _, o3, _, _ = ck.restore(tup2, "cpu", optimizer_factory=lambda pp, sl: gopt.GaussianOptimizer(pp, spatial_lr_scale=sl))
    assert o3.adam.group("rotation")["lr"] == 0.1 and o3.adam.state["f_rest"]["step"] == 1
    assert torch.allclose(o3.adam.state["scaling"]["exp_avg"], torch.full((37, 3), 0.1))
    # inference-only restore: no optimiser, statistics untouched
    p4, o4, st4, _ = ck.restore(tup, "cpu")
    assert o4 is None and float(st4.denom.sum()) == 0.0


def test_ply_layout_and_round_trip(tmp_path):
    p, _, _ = _model(P=23)
    names = ck.ply_attributes(p)
    assert names[:6] == ["x", "y", "z", "f_dc_0", "f_dc_1", "f_dc_2"] and names[6] == "f_rest_0" and names[50] == "f_rest_44"
    assert names[51:] == ["opacity", "normal_0", "normal_1", "normal_2", "albedo_0", "albedo_1", "albedo_2", "roughness",
                          "metallic", "scale_0", "scale_1", "scale_2", "rot_0", "rot_1", "rot_2", "rot_3"]
    path = str(tmp_path / "point_cloud.ply")
    ck.save_ply(p, path)
    raw = open(path, "rb").read()
    head = raw[:raw.index(b"end_header\n") + 11].decode()
    assert head.startswith("ply\nformat binary_little_endian 1.0\nelement vertex 23\nproperty float x\n")
    assert len(raw) == len(head) + 23 * 67 * 4                     # 67 floats per Gaussian
    rec = np.frombuffer(raw[len(head):], dtype="<f4").reshape(23, 67)
    # SH features are stored channel-major (save_ply transposes): f_rest_0..14 = channel 0 of the 15 coefficients
    assert np.array_equal(rec[:, 6:21], p.leaves["f_rest"].detach().numpy()[:, :, 0])
    q = ck.load_ply(path, "cpu")
    for k in gstep.PARAM_KEYS:
        assert torch.equal(q.leaves[k], p.leaves[k].detach()), k
    assert q.sh_degree == 3
    # an ASCII PLY with the attributes in another order loads too (attributes are looked up by name)
    apath = str(tmp_path / "ascii.ply")
    perm = list(reversed(names))
    with open(apath, "w") as fh:
        fh.write("ply\nformat ascii 1.0\ncomment test\nelement vertex 23\n" + "".join(f"property float {n}\n" for n in perm) + "end_header\n")
        for row in rec:
            fh.write(" ".join(repr(float(row[names.index(n)])) for n in perm) + "\n")
    q2 = ck.load_ply(apath, "cpu")
    assert torch.equal(q2.leaves["rot"], p.leaves["rot"].detach()) and torch.equal(q2.leaves["f_rest"], p.leaves["f_rest"].detach())
