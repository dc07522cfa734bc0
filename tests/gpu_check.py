"""Developer report (GPU): our kernels vs the reference CUDA build on the same inputs.
Prints mismatch statistics instead of asserting; the pytest -m gpu files hold the gates.
Usage: python tests/gpu_check.py [P] [W] [H] [start]
"""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gi-gs_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

import diff_gaussian_rasterization as dgr  # noqa: E402
import refshim  # noqa: E402
from gigs import scene  # noqa: E402


def cmp(name, a, b, report, exact=False):
    a = a.detach().float().cpu()
    b = b.detach().float().cpu()
    nan_a, nan_b = torch.isnan(a), torch.isnan(b)
    nan_mismatch = int((nan_a != nan_b).sum())
    ok = ~(nan_a | nan_b)
    diff = (a[ok] - b[ok]).abs()
    mx = float(diff.max()) if diff.numel() else 0.0
    nbad = int((diff > 0).sum())
    denom = float(b[ok].abs().max()) if diff.numel() else 0.0
    rel = float((a[ok] - b[ok]).norm() / (b[ok].norm() + 1e-30)) if diff.numel() else 0.0
    report[name] = dict(max_abs=mx, n_diff=nbad, n=int(a.numel()), nan_mismatch=nan_mismatch, ref_absmax=denom,
                        rel_l2=rel)
    flag = "OK " if (nbad == 0 and nan_mismatch == 0) else ("~  " if not exact else "BAD")
    print(f"{flag} {name:28s} max_abs={mx:.3e} rel_l2={rel:.3e} n_diff={nbad}/{a.numel()} nan_mismatch={nan_mismatch} "
          f"ref_absmax={denom:.3e}")


def main():
    P = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
    W = int(sys.argv[2]) if len(sys.argv) > 2 else 400
    H = int(sys.argv[3]) if len(sys.argv) > 3 else 300
    start = int(sys.argv[4]) if len(sys.argv) > 4 else 8
    dev = torch.device("cuda:0")
    report = {}
    raw = scene.make_scene(P, seed=0, regime="trained")
    g = scene.activate(raw, dev)
    cam = scene.orbit_camera(1, 8, W, H).to(dev)
    bg = torch.tensor([0.1, 0.2, 0.3], device=dev)

    # ---------------- reference forward
    ref = refshim.RefRasterizer()
    ro = ref.forward(g, cam, bg)
    rs = ref.state()
    R = ro["num_rendered"]
    print("reference num_rendered", R)

    # ---------------- ours forward (through the reference-shaped _C entry point)
    res = dgr._C.rasterize_gaussians(bg, g["means3D"], torch.Tensor([]), g["opacity"], g["normal"], g["albedo"],
                                     g["roughness"], g["metallic"], g["scales"], g["rotations"], torch.Tensor([]),
                                     g["shs"], cam.camera_center, cam.world_view_transform, cam.full_proj_transform,
                                     1.0, cam.tanfovx, cam.tanfovy, H, W, 3, False, False, False, False)
    (R2, color, radii, geom, binning, img, opacity, depth, normal, normal_view, pos, albedo, rough, metal) = res
    torch.cuda.synchronize()
    print("ours num_rendered", R2)
    report["num_rendered"] = dict(ref=R, ours=R2)
    lay = dgr.raster_layout(P, W, H, R2)
    sc = dgr.sort_scratch(dev)

    def view(buf, off, nbytes, dtype, shape):
        return buf[off:off + nbytes].view(dtype).reshape(shape)

    rec = view(geom, lay.g_record, P * 96, torch.float32, (P, 24))
    vis = rs["tiles_touched"] > 0
    print("visible", int(vis.sum()), "of", P)
    cmp("radii", radii, ro["radii"], report, True)
    cmp("tiles_touched", view(geom, lay.g_tiles_touched, 4 * P, torch.int32, (P,)), rs["tiles_touched"], report, True)
    cmp("means2D", rec[vis][:, 0:2], rs["means2D"][vis], report, True)
    cmp("conic", rec[vis][:, 2:5], rs["conic_opacity"][vis][:, 0:3], report, True)
    cmp("depths", rec[vis][:, 6], rs["depths"][vis], report, True)
    cmp("rgb", rec[vis][:, 8:11], rs["rgb"][vis], report, True)
    cmp("cov3D", view(geom, lay.g_cov3D, 24 * P, torch.float32, (P, 6))[vis], rs["cov3D"][vis], report, True)
    if R2 == R:
        import gpu_util as U
        st = U.decode_state(dict(num_rendered=R2, color=color, geom=geom, img=img, binning=binning), P, W, H)
        for nm in ("keys_unsorted", "vals_unsorted", "keys_sorted", "point_list"):
            a, b = st[nm], rs[nm]
            nd = int((a != b).sum())
            report[nm] = dict(n_diff=nd, n=int(a.numel()))
            print(("OK " if nd == 0 else "BAD"), nm, "n_diff", nd, "/", a.numel())
    T = ((W + 15) // 16) * ((H + 15) // 16)
    cmp("ranges", view(img, lay.i_ranges, 8 * T, torch.int32, (T, 2)), rs["ranges"], report, True)
    cmp("n_contrib", view(img, lay.i_n_contrib, 4 * W * H, torch.int32, (W * H,)), rs["n_contrib"], report, True)
    cmp("final_T", view(img, lay.i_final_T, 4 * W * H, torch.float32, (W * H,)), rs["final_T"], report)
    for nm, a, b in (("color", color, ro["color"]), ("opacity", opacity, ro["opacity"]), ("depth", depth, ro["depth"]),
                     ("normal", normal, ro["normal"]), ("normal_view", normal_view, ro["normal_view"]),
                     ("pos", pos, ro["pos"]), ("albedo", albedo, ro["albedo"]), ("roughness", rough, ro["roughness"]),
                     ("metallic", metal, ro["metallic"])):
        cmp("map_" + nm, a, b, report)

    # ---------------- backward
    gen = torch.Generator(device="cpu").manual_seed(1)
    N = W * H
    grads = {k: (torch.randn(c, H, W, generator=gen) / N).to(dev) for k, c in
             (("depth", 1), ("color", 3), ("opacity", 1), ("normal", 3), ("albedo", 3), ("roughness", 1),
              ("metallic", 1))}
    rb = ref.backward(g, cam, bg, ro["radii"], grads)
    ob = dgr._C.rasterize_gaussians_backward(bg, g["means3D"], radii, torch.Tensor([]), g["normal"], g["albedo"],
                                             g["roughness"], g["metallic"], g["scales"], g["rotations"],
                                             torch.Tensor([]), g["shs"], cam.camera_center, cam.world_view_transform,
                                             cam.full_proj_transform, 1.0, cam.tanfovx, cam.tanfovy, 3, grads["depth"],
                                             grads["color"], grads["opacity"], grads["normal"], grads["albedo"],
                                             grads["roughness"], grads["metallic"], geom, binning, img, R2, False)
    torch.cuda.synchronize()
    names = ("means2D", "colors", "opacity", "normal", "albedo", "roughness", "metallic", "means3D", "cov3D", "sh",
             "scales", "rotations")
    for nm, t in zip(names, ob):
        cmp("grad_" + nm, t, rb[nm].reshape(t.shape), report)
    # material-only fast path vs reference fed zeros
    zg = {k: torch.zeros_like(v) for k, v in grads.items()}
    for k in ("albedo", "roughness", "metallic"):
        zg[k] = grads[k]
    rb2 = ref.backward(g, cam, bg, ro["radii"], zg)
    ob2 = dgr._C.rasterize_gaussians_backward(bg, g["means3D"], radii, torch.Tensor([]), g["normal"], g["albedo"],
                                              g["roughness"], g["metallic"], g["scales"], g["rotations"],
                                              torch.Tensor([]), g["shs"], cam.camera_center, cam.world_view_transform,
                                              cam.full_proj_transform, 1.0, cam.tanfovx, cam.tanfovy, 3, None, None,
                                              None, None, grads["albedo"], grads["roughness"], grads["metallic"], geom,
                                              binning, img, R2, False, image_height=H, image_width=W)
    torch.cuda.synchronize()
    for nm, t in zip(names, ob2):
        cmp("matgrad_" + nm, t, rb2[nm].reshape(t.shape), report)

    # ---------------- screen-space
    fx = W / (2.0 * cam.tanfovx)
    fy = H / (2.0 * cam.tanfovy)
    V = cam.world_view_transform
    n_ref, p_ref = refshim.depth_to_normal(W, H, fx, fy, V, ro["depth"])
    n_our, p_our = dgr._C.depth_to_normal(W, H, fx, fy, V, ro["depth"])
    cmp("d2n_normal", n_our, n_ref, report)
    cmp("d2n_pos", p_our, p_ref, report)
    # unfused chain vs fused chain (ours vs ours + reference d2n in the middle)
    dmed = dgr.median_blur3x3(ro["depth"])
    n_mid, p_mid = refshim.depth_to_normal(W, H, fx, fy, V, dmed)
    n_chain_ref = dgr.bilateral_blur3x3(n_mid, 1.0, 3.0)
    p_chain_ref = dgr.median_blur3x3(p_mid)
    n_chain, p_chain = dgr.geometry_chain(W, H, fx, fy, V, ro["depth"], True)
    cmp("chain_normal", n_chain, n_chain_ref, report)
    cmp("chain_pos", p_chain, p_chain_ref, report)

    gi = dict(radius=0.8, bias=0.01, thick=0.05, delta=0.0625, step=16, start=start)
    t0 = time.time()
    occ_ref = refshim.ssao(W, H, fx, fy, gi["radius"], gi["bias"], gi["thick"], gi["delta"], gi["step"], gi["start"],
                           ro["normal_view"], p_chain_ref)
    t1 = time.time()
    occ = dgr._C.SSAO(W, H, fx, fy, gi["radius"], gi["bias"], gi["thick"], gi["delta"], gi["step"], gi["start"],
                      ro["normal_view"], p_chain_ref)
    torch.cuda.synchronize()
    t2 = time.time()
    print(f"ssao ref {t1 - t0:.4f}s ours {t2 - t1:.4f}s")
    cmp("ssao", occ, occ_ref, report)
    rgb = torch.rand(3, H, W, device=dev)
    F0 = (1.0 - ro["metallic"]) * 0.04 + ro["albedo"] * ro["metallic"]
    nv = torch.nn.functional.normalize(torch.nan_to_num(ro["normal_view"]), dim=0)
    t0 = time.time()
    c_ref, a_ref = refshim.ssr(W, H, fx, fy, gi["radius"], gi["bias"], gi["thick"], gi["delta"], gi["step"],
                               gi["start"], nv, p_chain_ref, rgb, ro["albedo"], ro["roughness"], ro["metallic"], F0)
    t1 = time.time()
    c_our, a_our = dgr._C.SSR(W, H, fx, fy, gi["radius"], gi["bias"], gi["thick"], gi["delta"], gi["step"],
                              gi["start"], nv, p_chain_ref, rgb, ro["albedo"], ro["roughness"], ro["metallic"], F0)
    torch.cuda.synchronize()
    t2 = time.time()
    print(f"ssr ref {t1 - t0:.4f}s ours {t2 - t1:.4f}s")
    cmp("ssr_color", c_our, c_ref, report)
    cmp("ssr_abd", a_our, a_ref, report)

    # ---------------- knn
    from simple_knn._C import distCUDA2
    pts = g["means3D"]
    d_ref = refshim.knn(pts)
    d_our = distCUDA2(pts)
    torch.cuda.synchronize()
    cmp("dist2", d_our, d_ref, report, True)
    cmp("mark_visible", dgr._C.mark_visible(pts, V, cam.full_proj_transform).float(),
        refshim.mark_visible(pts, V, cam.full_proj_transform).float(), report, True)

    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", f"gpu_check_{P}_{W}x{H}_s{start}.json"), "w") as f:
        json.dump(report, f, indent=1)


if __name__ == "__main__":
    main()
