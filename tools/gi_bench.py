"""SSAO / SSR march on the BASELINE configs[1] G-buffer (300k Gaussians, 800x800, start 8): every tuning variant of
gi_march_kernel timed with CUDA events, compared bit for bit with the reference-order loop (variant 0) and, when
oracle/_ref is built, with the reference's own SSAOCUDA / SSRCUDA; executed-probe counts for the roofline.
python tools/gi_bench.py [P] [W] [H]  ->  one JSON line"""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gi-gs_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch

from gigs import _lib, scene, shade, step as gstep

L = _lib.load()


def same(a, b):
    return bool(((a == b) | (torch.isnan(a) & torch.isnan(b))).all())


def ev_time(fn, flush, reps=5, warm=2):
    for _ in range(warm):
        fn()
    ts = []
    for i in range(reps):
        flush.fill_(float(i))
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    P = int(sys.argv[1]) if len(sys.argv) > 1 else 300000
    W = int(sys.argv[2]) if len(sys.argv) > 2 else 800
    H = int(sys.argv[3]) if len(sys.argv) > 3 else 800
    dev = torch.device("cuda:0")
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)
    raw = scene.make_scene(P, seed=0, regime="trained")
    cam = scene.orbit_camera(0, 8, W, H).to(dev)
    params = gstep.GaussianParams(raw, dev, light=scene.make_light(1))
    lut = shade.make_brdf_lut().to(dev)
    rays = scene.canonical_rays(cam, dev)
    gt = torch.rand(3, H, W, device=dev)
    gi = dict(radius=0.8, bias=0.01, thick=0.05, delta=0.0625, step=16, start=8)
    gstep.training_step(params, cam, params.light(), lut, rays, gt, torch.zeros(3, device=dev), gi)
    torch.cuda.synchronize()
    ws = params.last_workspace
    m = {k: ws.map(k).clone() for k in ("normal_view", "depth_pos", "ssr_normal", "linear_rgb", "albedo", "rough_remap",
                                        "metal_used", "F0")}
    fx, fy = W / (2 * cam.tanfovx), H / (2 * cam.tanfovy)
    args = (W, H, fx, fy, gi["radius"], gi["bias"], gi["thick"], gi["delta"], gi["step"], gi["start"])
    occ = torch.empty(1, H, W, device=dev)
    col = torch.empty(3, H, W, device=dev)
    abd = torch.empty(3, H, W, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    scr = torch.empty(int(L.gigs_gi_scratch_bytes(W, H)), dtype=torch.uint8, device=dev)

    def ssao():
        _lib.check(L.gigs_ssao(*args, m["normal_view"].data_ptr(), m["depth_pos"].data_ptr(), occ.data_ptr(), scr.data_ptr(), scr.numel(), st), "ssao")

    def ssr():
        _lib.check(L.gigs_ssr(*args, m["ssr_normal"].data_ptr(), m["depth_pos"].data_ptr(), m["linear_rgb"].data_ptr(),
                              m["albedo"].data_ptr(), m["rough_remap"].data_ptr(), m["metal_used"].data_ptr(),
                              m["F0"].data_ptr(), col.data_ptr(), abd.data_ptr(), scr.data_ptr(), scr.numel(), st), "ssr")

    out = {"P": P, "W": W, "H": H, "gi": gi, "variants": {}}
    if os.environ.get("GI_QUICK"):          # profiling runs: one SSAO + one SSR launch of the chosen variant
        _lib.check(L.gigs_gi_tune(int(os.environ["GI_QUICK"]), int(os.environ.get("GI_BLOCK", "1"))), "tune")
        ssao(); ssr(); torch.cuda.synchronize()
        print(json.dumps({"quick": int(os.environ["GI_QUICK"])}))
        return
    cnt = torch.zeros(2, dtype=torch.int64, device=dev)
    for nm, nrm in (("ssao", "normal_view"), ("ssr", "ssr_normal")):
        _lib.check(L.gigs_gi_count_probes(*args, m[nrm].data_ptr(), m["depth_pos"].data_ptr(), None, 0, cnt.data_ptr(), st), "count")
        torch.cuda.synchronize()
        out["probes_" + nm] = int(cnt[0].item())
    if os.environ.get("GI_HIZ"):
        # experiment: how many probes survive a conservative (min, max) block test of the depth plane?
        z = m["depth_pos"][2]
        out["hiz_kept_fraction"] = {}
        for B in (2, 4, 8, 16):
            hb, wb = (H + B - 1) // B, (W + B - 1) // B
            zp = torch.full((hb * B, wb * B), float("nan"), device=dev)
            zp[:H, :W] = z
            blk = zp.view(hb, B, wb, B).permute(0, 2, 1, 3).reshape(hb, wb, B * B)
            mn = torch.where(torch.isnan(blk), torch.full_like(blk, float("inf")), blk).amin(-1)
            mx = torch.where(torch.isnan(blk), torch.full_like(blk, float("-inf")), blk).amax(-1)
            for prec in ("f32", "f16"):
                if prec == "f16":  # outward rounding to half
                    mnh = mn.half(); mnh = torch.where(mnh.float() > mn, torch.nextafter(mnh, torch.full_like(mnh, -float("inf"))), mnh)
                    mxh = mx.half(); mxh = torch.where(mxh.float() < mx, torch.nextafter(mxh, torch.full_like(mxh, float("inf"))), mxh)
                    mm = torch.stack([mnh.float(), mxh.float()], -1).contiguous()
                else:
                    mm = torch.stack([mn, mx], -1).contiguous()
                _lib.check(L.gigs_gi_count_probes(*args, m["normal_view"].data_ptr(), m["depth_pos"].data_ptr(),
                                                  mm.data_ptr(), B, cnt.data_ptr(), st), "count")
                torch.cuda.synchronize()
                out["hiz_kept_fraction"][f"B{B}_{prec}"] = int(cnt[1].item()) / max(1, int(cnt[0].item()))
    out["probes_upper_bound"] = W * H * 512 * (gi["step"] - gi["start"])
    base = None
    outs = {}
    for v in (0, 1, 2, 11, 12):           # pairs per step; +10 = with the block (min, max) test
        _lib.check(L.gigs_gi_tune(v % 10, v // 10), "tune")
        ssao(); ssr(); torch.cuda.synchronize()
        cur = (occ.clone(), col.clone(), abd.clone())
        outs[v] = cur
        rec = {"ssao_ms": ev_time(ssao, flush), "ssr_ms": ev_time(ssr, flush)}
        if base is None:
            base = cur
        else:
            rec["bit_identical_to_variant0"] = [same(a, b) for a, b in zip(cur, base)]
            rec["max_abs_diff"] = [float((a - b).abs().nan_to_num(0).max()) for a, b in zip(cur, base)]
        out["variants"][str(v)] = rec
    _lib.check(L.gigs_gi_tune(1, 1), "tune")
    try:
        import refshim
        if refshim.available():
            gl = (gi["radius"], gi["bias"], gi["thick"], gi["delta"], gi["step"], gi["start"])
            o_r = refshim.ssao(W, H, fx, fy, *gl, m["normal_view"], m["depth_pos"])
            c_r, a_r = refshim.ssr(W, H, fx, fy, *gl, m["ssr_normal"], m["depth_pos"], m["linear_rgb"], m["albedo"],
                                   m["rough_remap"], m["metal_used"], m["F0"])
            out["vs_reference_kernels"] = {"bit_identical": [same(base[0], o_r), same(base[1], c_r), same(base[2], a_r)],
                                           "max_abs_diff": [float((a - b).abs().nan_to_num(0).max())
                                                            for a, b in zip(base, (o_r, c_r, a_r))]}
            ndiff = lambda a, b: int((~((a == b) | (torch.isnan(a) & torch.isnan(b)))).sum())
            out["differing_values_vs_reference"] = {str(v): [ndiff(x, y) for x, y in zip(outs[v], (o_r, c_r, a_r))]
                                                    for v in outs}
            bad = (~((outs[0][0] == o_r) | (torch.isnan(o_r) & torch.isnan(outs[0][0])))).nonzero()[:6]
            out["first_ssao_mismatches_v0"] = [[int(i) for i in r] + [float(outs[0][0][tuple(r)]), float(o_r[tuple(r)]),
                                                                      float(outs[1][0][tuple(r)])] for r in bad]
            out["ref_ssao_ms"] = ev_time(lambda: refshim.ssao(W, H, fx, fy, *gl, m["normal_view"], m["depth_pos"]), flush, 3, 1)
            out["ref_ssr_ms"] = ev_time(lambda: refshim.ssr(W, H, fx, fy, *gl, m["ssr_normal"], m["depth_pos"],
                                                            m["linear_rgb"], m["albedo"], m["rough_remap"],
                                                            m["metal_used"], m["F0"]), flush, 3, 1)
    except OSError as e:
        out["vs_reference_kernels"] = str(e)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
