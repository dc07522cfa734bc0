"""Are the getters fused into preprocess_kernel<RAW> (sigmoid / exp / normalize on load) bit-identical to torch's CUDA
kernels? Compares the packed blend records and cov3D written by the frame path fed raw leaves with those written by
the operator path fed torch's activations (scene.activate). One JSON line with differing-value counts per field."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gi-gs_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch

import diff_gaussian_rasterization as dgr
import gpu_util as U
from gigs import scene, shade, step as gstep


def main():
    P = int(sys.argv[1]) if len(sys.argv) > 1 else 300000
    W = H = int(sys.argv[2]) if len(sys.argv) > 2 else 800
    dev = torch.device("cuda:0")
    raw = scene.make_scene(P, seed=0, regime="trained")
    mode = os.environ.get("GETTER_MODE", "")
    if mode == "unitq":
        raw["rot"] = torch.zeros_like(raw["rot"]); raw["rot"][:, 0] = 1.0
    if mode == "unitscale":
        raw["log_scale"] = torch.full_like(raw["log_scale"], -4.0)
    cam = scene.orbit_camera(0, 8, W, H).to(dev)
    bg = torch.zeros(3, device=dev)
    g = scene.activate(raw, dev)
    fo = U.ours_forward(g, cam, bg)
    lay = dgr.raster_layout(P, W, H, fo["num_rendered"])
    rec_op = fo["geom"][lay.g_record:lay.g_record + P * 96].view(torch.float32).view(P, 24).clone()
    cov_op = fo["geom"][lay.g_cov3D:lay.g_cov3D + P * 24].view(torch.float32).view(P, 6).clone()
    vis = fo["radii"] > 0
    params = gstep.GaussianParams(raw, dev, light=scene.make_light(1, base_res=64))
    lut = shade.make_brdf_lut(64, 64).to(dev)
    rays = scene.canonical_rays(cam, dev)
    gt = torch.rand(3, H, W, device=dev)
    gi = dict(radius=0.8, bias=0.01, thick=0.05, delta=0.0625, step=16, start=64)
    gstep.training_step(params, cam, params.light(), lut, rays, gt, bg, gi, radiance=True)
    torch.cuda.synchronize()
    ws = params.last_workspace
    rec_fr = ws.geom[lay.g_record:lay.g_record + P * 96].view(torch.float32).view(P, 24)
    cov_fr = ws.geom[lay.g_cov3D:lay.g_cov3D + P * 24].view(torch.float32).view(P, 6)
    out = {"P": P, "visible": int(vis.sum()), "radii_equal": bool(torch.equal(ws.radii, fo["radii"]))}
    names = {"xy": (0, 2), "conic": (2, 5), "opacity": (5, 6), "depth": (6, 7), "tau": (7, 8), "rgb": (8, 11),
             "rough": (11, 12), "albedo": (12, 15), "metal": (15, 16), "normal": (16, 19), "pos": (19, 22)}
    for k, (a, b) in names.items():
        x, y = rec_fr[vis, a:b], rec_op[vis, a:b]
        out[k] = int((x.view(torch.int32) != y.view(torch.int32)).sum())
    out["cov3D"] = int((cov_fr[vis].view(torch.int32) != cov_op[vis].view(torch.int32)).sum())
    # which activation is responsible: torch's own kernels against the formulas used in the kernel, evaluated by torch
    x = raw["opacity"].to(dev)
    out["sigmoid_formula_vs_torch"] = int((torch.sigmoid(x) != 1.0 / (1.0 + torch.exp(-x))).sum())
    q = raw["rot"].to(dev)
    n1 = torch.nn.functional.normalize(q, dim=-1)
    nn = torch.sqrt(((q[:, 0] * q[:, 0] + q[:, 1] * q[:, 1]) + q[:, 2] * q[:, 2]) + q[:, 3] * q[:, 3]).clamp_min(1e-12)
    out["normalize4_sequential_sum_vs_torch"] = int((n1 != q / nn[:, None]).sum())
    nn2 = torch.sqrt((q[:, 0] * q[:, 0] + q[:, 2] * q[:, 2]) + (q[:, 1] * q[:, 1] + q[:, 3] * q[:, 3])).clamp_min(1e-12)
    out["normalize4_pairwise02_13_vs_torch"] = int((n1 != q / nn2[:, None]).sum())
    nn3 = torch.sqrt((q[:, 0] * q[:, 0] + q[:, 1] * q[:, 1]) + (q[:, 2] * q[:, 2] + q[:, 3] * q[:, 3])).clamp_min(1e-12)
    out["normalize4_pairwise01_23_vs_torch"] = int((n1 != q / nn3[:, None]).sum())
    v = raw["normal"].to(dev)
    m1 = torch.nn.functional.normalize(v, dim=-1)
    mm = torch.sqrt((v[:, 0] * v[:, 0] + v[:, 1] * v[:, 1]) + v[:, 2] * v[:, 2]).clamp_min(1e-12)
    out["normalize3_sequential_sum_vs_torch"] = int((m1 != v / mm[:, None]).sum())
    sq = v * v
    for nm, t in (("02_1", (sq[:, 0] + sq[:, 2]) + sq[:, 1]), ("0_12", sq[:, 0] + (sq[:, 1] + sq[:, 2])),
                  ("12_0", (sq[:, 1] + sq[:, 2]) + sq[:, 0])):
        out["normalize3_" + nm] = int((m1 != v / torch.sqrt(t).clamp_min(1e-12)[:, None]).sum())
    # map-level F.normalize(x, dim=0) and torch.norm(x, dim=0) on a [3, H, W] map (outer reduction)
    mp = torch.randn(3, 300, 400, device=dev)
    t1 = torch.nn.functional.normalize(mp, dim=0, p=2)
    tn = torch.norm(mp, dim=0, keepdim=True)
    sq = mp * mp
    for nm, t in (("seq", (sq[0] + sq[1]) + sq[2]), ("02_1", (sq[0] + sq[2]) + sq[1]), ("0_12", sq[0] + (sq[1] + sq[2]))):
        nrm = torch.sqrt(t)
        out["map_normalize_dim0_" + nm] = int((t1 != mp / nrm.clamp_min(1e-12)[None]).sum())
        out["map_norm_dim0_" + nm] = int((tn[0] != nrm).sum())
    print(json.dumps(out))


if __name__ == "__main__":
    main()
