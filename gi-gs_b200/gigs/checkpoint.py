"""Interchange of trained scenes with the reference (data formats either side of the path):
  * capture() / restore(): the 18-tuple of GaussianModel.capture / restore (/root/reference/scene/gaussian_model.py:82-175)
    that train.py:466-487 stores under "gaussians" in chkpnt<iter>.pth, optimiser state in torch.optim.Adam's
    state_dict layout — a checkpoint written here loads into the reference's GaussianModel.restore and vice versa;
  * save_ply() / load_ply(): the vertex attribute layout of GaussianModel.save_ply / load_ply (:397-579), written and
    parsed directly (binary little-endian PLY, the format plyfile writes by default; plyfile is not a dependency).
Host-side I/O only; nothing here is on the timed path.
"""
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch

from .optim import REFERENCE_GROUP_NAME
from .step import PARAM_KEYS, GaussianParams

# tuple positions 1..10 of capture() (:104-115)
_CAPTURE_ORDER = ("xyz", "f_dc", "f_rest", "log_scale", "rot", "opacity", "normal", "albedo", "roughness", "metallic")


def capture(params: GaussianParams, optimizer, state, spatial_lr_scale: float = 1.0) -> Tuple:
    """GaussianModel.capture(): (active_sh_degree, _xyz, _features_dc, _features_rest, _scaling, _rotation, _opacity,
    _normal, _albedo, _roughness, _metallic, max_radii2D, xyz_gradient_accum, xyz_gradient_accum_abs,
    xyz_gradient_accum_abs_max, denom, optimizer.state_dict(), spatial_lr_scale)."""
    L = params.leaves
    sd = _gaussian_state_dict(optimizer)
    return (params.sh_degree, *[L[k] for k in _CAPTURE_ORDER], state.max_radii2D, state.xyz_gradient_accum,
            state.xyz_gradient_accum_abs, state.xyz_gradient_accum_abs_max, state.denom, sd, spatial_lr_scale)


def _gaussian_state_dict(optimizer) -> Dict:
    """The Gaussians' optimiser only (the light's is saved separately as "light_optimizer", train.py:473), groups in
    training_setup's order so that parameter ids line up with the reference's."""
    sd = optimizer.adam.state_dict()
    keep = [i for i, g in enumerate(sd["param_groups"]) if g["name"] in REFERENCE_GROUP_NAME.values()]
    groups = [dict(sd["param_groups"][i], params=[j]) for j, i in enumerate(keep)]
    state = {j: sd["state"][i] for j, i in enumerate(keep) if i in sd["state"]}
    return dict(state=state, param_groups=groups)


def restore(model_args: Tuple, device, optimizer_factory=None, light_base: Optional[torch.Tensor] = None):
    """GaussianModel.restore(): returns (params, optimizer, densify_state, spatial_lr_scale). optimizer_factory(params,
    spatial_lr_scale) builds the GaussianOptimizer (= training_setup) whose state is then loaded; None = no training."""
    from .densify import DensifyState
    (sh_degree, *tensors) = model_args[:11]
    max_radii2D, accum, accum_abs, accum_abs_max, denom, opt_dict, spatial_lr_scale = model_args[11:18]
    raw = {k: t.detach() for k, t in zip(_CAPTURE_ORDER, tensors)}
    raw["sh_degree"] = int(sh_degree)
    params = GaussianParams(raw, device, light_base=light_base)
    st = DensifyState(params.P, device)
    st.max_radii2D = max_radii2D.to(device).float()
    optimizer = None
    if optimizer_factory is not None:
        optimizer = optimizer_factory(params, spatial_lr_scale)
        st.xyz_gradient_accum, st.xyz_gradient_accum_abs = accum.to(device).float(), accum_abs.to(device).float()
        st.xyz_gradient_accum_abs_max, st.denom = accum_abs_max.to(device).float(), denom.to(device).float()
        names = [g["name"] for g in opt_dict["param_groups"]]
        for i, name in enumerate(names):
            g = optimizer.adam.group(name)
            src = opt_dict["param_groups"][i]
            g["lr"], g["betas"], g["eps"] = float(src["lr"]), tuple(src["betas"]), float(src["eps"])
            s = opt_dict["state"].get(i)
            if s is not None:
                p = g["params"][0]
                optimizer.adam.state[name] = dict(step=int(float(s["step"])),
                                                  exp_avg=s["exp_avg"].to(p.device, torch.float32).contiguous().clone(),
                                                  exp_avg_sq=s["exp_avg_sq"].to(p.device, torch.float32).contiguous().clone())
    return params, optimizer, st, spatial_lr_scale


def ply_attributes(params: GaussianParams) -> List[str]:
    """construct_list_of_attributes (:397-415)."""
    L = params.leaves
    l = ["x", "y", "z"]
    l += [f"f_dc_{i}" for i in range(L["f_dc"].shape[1] * L["f_dc"].shape[2])]
    l += [f"f_rest_{i}" for i in range(L["f_rest"].shape[1] * L["f_rest"].shape[2])]
    l.append("opacity")
    l += [f"normal_{i}" for i in range(L["normal"].shape[1])]
    l += [f"albedo_{i}" for i in range(L["albedo"].shape[1])]
    l += ["roughness", "metallic"]
    l += [f"scale_{i}" for i in range(L["log_scale"].shape[1])]
    l += [f"rot_{i}" for i in range(L["rot"].shape[1])]
    return l


def save_ply(params: GaussianParams, path: str) -> None:
    """GaussianModel.save_ply (:417-465): one float32 property per attribute; SH features channel-major
    (transpose(1,2).flatten)."""
    L = {k: v.detach().cpu() for k, v in params.leaves.items()}
    cols = [L["xyz"], L["f_dc"].transpose(1, 2).flatten(start_dim=1), L["f_rest"].transpose(1, 2).flatten(start_dim=1),
            L["opacity"], L["normal"], L["albedo"], L["roughness"], L["metallic"], L["log_scale"], L["rot"]]
    data = torch.cat(cols, dim=1).contiguous().numpy().astype("<f4")
    names = ply_attributes(params)
    assert data.shape[1] == len(names)
    header = "ply\nformat binary_little_endian 1.0\nelement vertex %d\n" % data.shape[0]
    header += "".join(f"property float {n}\n" for n in names) + "end_header\n"
    with open(path, "wb") as fh:
        fh.write(header.encode("ascii"))
        fh.write(data.tobytes())


def _read_ply(path: str) -> Dict[str, np.ndarray]:
    types = {"float": "f4", "float32": "f4", "double": "f8", "float64": "f8", "uchar": "u1", "uint8": "u1",
             "char": "i1", "int8": "i1", "short": "i2", "int16": "i2", "ushort": "u2", "uint16": "u2", "int": "i4",
             "int32": "i4", "uint": "u4", "uint32": "u4"}
    with open(path, "rb") as fh:
        if fh.readline().strip() != b"ply":
            raise ValueError(f"{path}: not a PLY file")
        fmt, n, props, in_vertex = None, 0, [], False
        while True:
            line = fh.readline()
            if not line:
                raise ValueError(f"{path}: truncated PLY header")
            tok = line.decode("ascii").split()
            if not tok or tok[0] == "comment":
                continue
            if tok[0] == "format":
                fmt = tok[1]
            elif tok[0] == "element":
                in_vertex = tok[1] == "vertex" and not props
                if in_vertex:                   # further elements (faces) are ignored
                    n = int(tok[2])
            elif tok[0] == "property" and in_vertex:
                if tok[1] == "list":
                    raise ValueError(f"{path}: list property in the vertex element")
                props.append((tok[2], types[tok[1]]))
            elif tok[0] == "end_header":
                break
        if fmt == "ascii":
            arr = np.loadtxt(fh, max_rows=n, ndmin=2)
            return {name: arr[:, i] for i, (name, _) in enumerate(props)}
        order = "<" if fmt == "binary_little_endian" else ">"
        dt = np.dtype([(name, order + t) for name, t in props])
        rec = np.frombuffer(fh.read(n * dt.itemsize), dtype=dt, count=n)
        return {name: rec[name] for name, _ in props}


def load_ply(path: str, device, max_sh_degree: int = 3, light_base: Optional[torch.Tensor] = None) -> GaussianParams:
    """GaussianModel.load_ply (:474-578): attributes by name, f_rest_* / scale_* / rot_* sorted by their index."""
    v = _read_ply(path)
    P = v["x"].shape[0]

    def stack(names):
        return np.stack([np.asarray(v[n], dtype=np.float32) for n in names], axis=1)

    def numbered(prefix):
        return sorted([n for n in v if n.startswith(prefix)], key=lambda x: int(x.split("_")[-1]))
    extra = numbered("f_rest_")
    assert len(extra) == 3 * (max_sh_degree + 1) ** 2 - 3
    f_dc = stack(["f_dc_0", "f_dc_1", "f_dc_2"]).reshape(P, 3, 1)
    f_rest = stack(extra).reshape(P, 3, (max_sh_degree + 1) ** 2 - 1)
    raw = dict(xyz=stack(["x", "y", "z"]), opacity=stack(["opacity"]),
               normal=stack(["normal_0", "normal_1", "normal_2"]), albedo=stack(["albedo_0", "albedo_1", "albedo_2"]),
               roughness=stack(["roughness"]), metallic=stack(["metallic"]),
               f_dc=np.ascontiguousarray(f_dc.transpose(0, 2, 1)), f_rest=np.ascontiguousarray(f_rest.transpose(0, 2, 1)),
               log_scale=stack(numbered("scale_")), rot=stack([n for n in numbered("rot")]))
    raw = {k: torch.from_numpy(np.ascontiguousarray(a)) for k, a in raw.items()}
    raw["sh_degree"] = max_sh_degree          # active_sh_degree = max_sh_degree (:578)
    return GaussianParams(raw, device, light_base=light_base)
