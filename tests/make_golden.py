"""Generate the golden fixtures under tests/golden/ by running the REFERENCE's own CUDA kernels
(oracle/_ref/libgigs_ref.so, built from /root/reference by oracle/Makefile) on a B200.

    gpurun -- 'python tests/make_golden.py'      # writes gpurun_out/golden/*.npz, copy them to tests/golden/

The inputs are regenerated from (P, seed, regime, W, H, camera k/K) by gigs.scene, so only the outputs and the
upstream gradients (seeded) are stored. These fixtures pin the CPU oracle (tests/test_oracle_golden.py) and
are re-checked against our kernels on the GPU (tests/test_gpu_parity.py::test_against_committed_golden_vectors).
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gi-gs_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

import refshim  # noqa: E402
from gigs import scene  # noqa: E402

CASES = {
    # name: P, seed, regime, W, H, cam k, K, sh_degree, inference, start
    "tiny_trained": dict(P=600, seed=3, regime="trained", W=64, H=48, k=1, K=8, deg=3, inference=False, start=8),
    "odd_inference": dict(P=400, seed=5, regime="trained", W=53, H=37, k=3, K=8, deg=1, inference=True, start=8),
    "init_like": dict(P=1500, seed=7, regime="init", W=80, H=64, k=5, K=8, deg=0, inference=False, start=64),
}
GI = dict(radius=0.8, bias=0.01, thick=0.05, delta=0.0625, step=16)


def upstream_grads(W, H, seed):
    gen = torch.Generator(device="cpu").manual_seed(1000 + seed)
    N = W * H
    return {k: (torch.randn(c, H, W, generator=gen) / N) for k, c in
            (("depth", 1), ("color", 3), ("opacity", 1), ("normal", 3), ("albedo", 3), ("roughness", 1),
             ("metallic", 1))}


def case_inputs(c, device="cpu"):
    raw = scene.make_scene(c["P"], seed=c["seed"], regime=c["regime"], sh_degree=3)
    g = scene.activate(raw, device)
    g["sh_degree"] = c["deg"]
    cam = scene.orbit_camera(c["k"], c["K"], c["W"], c["H"]).to(device)
    bg = torch.tensor([0.1, 0.2, 0.3], device=device)
    return g, cam, bg


def main():
    dev = torch.device("cuda:0")
    outdir = os.path.join(ROOT, "gpurun_out", "golden")
    os.makedirs(outdir, exist_ok=True)
    for name, c in CASES.items():
        g, cam, bg = case_inputs(c, dev)
        W, H = c["W"], c["H"]
        ref = refshim.RefRasterizer()
        ro = ref.forward(g, cam, bg, sh_degree=c["deg"], inference=c["inference"])
        st = ref.state()
        grads = {k: v.to(dev) for k, v in upstream_grads(W, H, c["seed"]).items()}
        rb = ref.backward(g, cam, bg, ro["radii"], grads)
        fx, fy = W / (2.0 * cam.tanfovx), H / (2.0 * cam.tanfovy)
        V = cam.world_view_transform
        d2n_n, d2n_p = refshim.depth_to_normal(W, H, fx, fy, V, ro["depth"])
        # SSAO / SSR on reference-produced inputs (pos = plain depth_to_normal position; no third-party filters)
        occ = refshim.ssao(W, H, fx, fy, GI["radius"], GI["bias"], GI["thick"], GI["delta"], GI["step"], c["start"],
                           ro["normal_view"], d2n_p)
        gen = torch.Generator(device="cpu").manual_seed(2000 + c["seed"])
        rgb = torch.rand(3, H, W, generator=gen).to(dev)
        F0 = (1.0 - ro["metallic"]) * 0.04 + ro["albedo"] * ro["metallic"]
        ssr_c, ssr_a = refshim.ssr(W, H, fx, fy, GI["radius"], GI["bias"], GI["thick"], GI["delta"], GI["step"],
                                   c["start"], ro["normal_view"], d2n_p, rgb, ro["albedo"], ro["roughness"],
                                   ro["metallic"], F0)
        d2 = refshim.knn(g["means3D"])
        vis = refshim.mark_visible(g["means3D"], V, cam.full_proj_transform)
        out = {}
        for k in ("color", "opacity", "depth", "normal", "normal_view", "pos", "albedo", "roughness", "metallic"):
            out["map_" + k] = ro[k].cpu().numpy()
        out["radii"] = ro["radii"].cpu().numpy()
        out["num_rendered"] = np.int64(ro["num_rendered"])
        for k in ("depths", "means2D", "conic_opacity", "rgb", "cov3D", "tiles_touched", "point_offsets",
                  "keys_unsorted", "keys_sorted", "vals_unsorted", "point_list", "final_T", "n_contrib", "ranges",
                  "clamped"):
            out["st_" + k] = st[k].cpu().numpy()
        for k, v in rb.items():
            out["grad_" + k] = v.cpu().numpy()
        out["d2n_normal"] = d2n_n.cpu().numpy(); out["d2n_pos"] = d2n_p.cpu().numpy()
        out["ssao"] = occ.cpu().numpy(); out["ssr_color"] = ssr_c.cpu().numpy(); out["ssr_abd"] = ssr_a.cpu().numpy()
        out["ssr_rgb_in"] = rgb.cpu().numpy()
        out["dist2"] = d2.cpu().numpy(); out["mark_visible"] = vis.cpu().numpy()
        path = os.path.join(outdir, name + ".npz")
        np.savez_compressed(path, **out)
        print(name, "R =", ro["num_rendered"], "->", path, os.path.getsize(path) // 1024, "KiB")
        ref.close()


if __name__ == "__main__":
    main()
