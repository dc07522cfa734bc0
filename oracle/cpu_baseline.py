"""TEST / BENCH INFRASTRUCTURE — times the CPU oracle (oracle/gigs_oracle.py) on a BOUNDED SAMPLE of the
PBR-stage training step and extrapolates to the full frame. Used only by bench.py (`cpu_baseline` and
`--impl reference`). The reference has no CPU path; this transcription is the reported CPU baseline
(BASELINE.md §2), not an optimisation target.

Sample: per-Gaussian stages (preprocess, binning incl. the stable sort, per-Gaussian backward) and the
full-frame filter chain run on the WHOLE workload; the per-tile blend forward/backward run on `n_tiles`
tiles out of T and the per-pixel GI march / shading on `n_pix` pixels out of N, each scaled by T/n_tiles
or N/n_pix.
"""
import os
import sys
import time
from typing import Dict

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
if _HERE not in sys.path:
    sys.path.insert(0, _HERE)
import gigs_oracle as O  # noqa: E402


def cpu_step_sample(g: Dict, cam, bg, light: Dict, lut, gi: Dict, n_tiles: int = 12, n_pix: int = 4096,
                    seed: int = 0) -> Dict:
    W, H = cam.image_width, cam.image_height
    N = W * H
    gen = torch.Generator().manual_seed(seed)
    t = {}
    t0 = time.perf_counter()
    pre = O.preprocess(g, cam)
    binn = O.binning(pre)
    t["per_gaussian_fwd"] = time.perf_counter() - t0
    T = pre["grid"][0] * pre["grid"][1]
    n_tiles = min(n_tiles, T)
    tiles = torch.randperm(T, generator=gen)[:n_tiles].sort().values
    t0 = time.perf_counter()
    fwd = O.blend_forward(pre, binn, g, cam, bg, tiles=tiles)
    t["blend_fwd_sample"] = time.perf_counter() - t0
    grads = {"albedo": torch.randn(3, H, W, generator=gen) / N, "roughness": torch.randn(1, H, W, generator=gen) / N,
             "metallic": torch.randn(1, H, W, generator=gen) / N}
    t0 = time.perf_counter()
    acc = O.blend_backward(pre, binn, g, cam, bg, fwd, grads, tiles=tiles)
    t["blend_bwd_sample"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    O.gaussian_backward(pre, g, cam, acc)
    t["per_gaussian_bwd"] = time.perf_counter() - t0
    fx, fy = W / (2.0 * cam.tanfovx), H / (2.0 * cam.tanfovy)
    t0 = time.perf_counter()
    nfd, pos = O.geometry_chain(W, H, fx, fy, cam.world_view_transform, fwd["depth"])
    O.median3x3(fwd["normal"]); O.median3x3(fwd["normal_view"]); O.median3x3(fwd["albedo"])  # render() / IRR medians
    t["filters_full"] = time.perf_counter() - t0
    n_pix = min(n_pix, N)
    pix = torch.randperm(N, generator=gen)[:n_pix].sort().values
    rgb = torch.rand(3, H, W, generator=gen)
    F0 = (1.0 - fwd["metallic"]) * 0.04 + fwd["albedo"] * fwd["metallic"]
    t0 = time.perf_counter()
    O.ssao(W, H, fx, fy, gi["radius"], gi["bias"], gi["thick"], gi["delta"], gi["step"], gi["start"],
           fwd["normal_view"], pos, pix_sel=pix)
    O.ssr(W, H, fx, fy, gi["radius"], gi["bias"], gi["thick"], gi["delta"], gi["step"], gi["start"],
          torch.nan_to_num(fwd["normal_view"]), pos, rgb, fwd["albedo"], fwd["roughness"], fwd["metallic"], F0,
          pix_sel=pix)
    t["gi_sample"] = time.perf_counter() - t0
    # shading forward + backward (autograd of the oracle) on an n_pix-pixel strip
    hs = max(1, n_pix // W)
    sl = slice(0, hs)
    nrm = torch.nn.functional.normalize(torch.randn(hs, W, 3, generator=gen), dim=-1)
    vd = torch.nn.functional.normalize(torch.randn(hs, W, 3, generator=gen), dim=-1)
    alb = fwd["albedo"].permute(1, 2, 0)[sl].clone().requires_grad_(True)
    rgh = (fwd["roughness"].permute(1, 2, 0)[sl] * 0.96 + 0.04).clone().requires_grad_(True)
    met = fwd["metallic"].permute(1, 2, 0)[sl].clone().requires_grad_(True)
    lt = dict(diffuse=light["diffuse"].clone().requires_grad_(True),
              specular=[s.clone().requires_grad_(True) for s in light["specular"]])
    t0 = time.perf_counter()
    res = O.pbr_shading(lt, nrm, vd, alb, rgh, torch.ones(hs, W, 1, dtype=torch.bool), gamma=True,
                        occlusion=torch.ones(hs, W, 1), metallic=met, brdf_lut=lut)
    res["render_rgb"].abs().mean().backward()
    t["shade_sample"] = time.perf_counter() - t0
    # BRDF smoothness prior of the loss (train.py:388-402), forward + backward on the full frame
    pred = torch.cat([fwd["albedo"], fwd["roughness"] * 0.96 + 0.04, fwd["metallic"]], 0).clone().requires_grad_(True)
    gt_img = torch.rand(3, H, W, generator=gen)
    t0 = time.perf_counter()
    O.masked_tv_loss(torch.ones(1, H, W, dtype=torch.bool), gt_img, pred).backward()
    t["brdf_tv_full"] = time.perf_counter() - t0
    est = (t["per_gaussian_fwd"] + t["per_gaussian_bwd"] + t["filters_full"] + t["brdf_tv_full"]
           + (t["blend_fwd_sample"] + t["blend_bwd_sample"]) * (T / n_tiles)
           + t["gi_sample"] * (N / n_pix) + t["shade_sample"] * (N / (hs * W)))
    return dict(times=t, est_frame_s=est, frames_per_s=1.0 / est, wall_s=sum(t.values()),
                sample=f"per-Gaussian stages + 3x3 filter chain + BRDF TV prior on the full {g['means3D'].shape[0]}-Gaussian {W}x{H} "
                       f"frame; blend fwd+bwd on {n_tiles}/{T} tiles; SSAO+SSR on {n_pix}/{N} pixels; shade fwd+bwd on "
                       f"{hs * W}/{N} pixels; each scaled to the full frame")
