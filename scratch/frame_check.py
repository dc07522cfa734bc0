import os, sys, time, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gi-gs_b200"))
from gigs import scene, shade, step as gstep, renderer
dev = torch.device("cuda:0")

def run(P, W, H, start, base_res=256, metallic=True):
    raw = scene.make_scene(P, seed=3, regime="trained")
    lut = shade.make_brdf_lut().to(dev)
    cam = scene.orbit_camera(1, 8, W, H).to(dev)
    rays = scene.canonical_rays(cam, dev)
    gt = torch.rand(3, H, W, device=dev, generator=torch.Generator(device=dev).manual_seed(5))
    bg = torch.zeros(3, device=dev)
    gi = dict(radius=0.8, bias=0.01, thick=0.05, delta=0.0625, step=16, start=start)
    out = {}
    for fused in (False, True):
        params = gstep.GaussianParams(raw, dev, light=scene.make_light(0, base_res=base_res))
        params.zero_grad()
        loss = gstep.training_step(params, cam, params.light(), lut, rays, gt, bg, gi, fused=fused, metallic=metallic)
        torch.cuda.synchronize()
        out[fused] = (loss.item(), params.flat_grad.clone(), params)
    l0, g0, p0 = out[False]; l1, g1, p1 = out[True]
    print(f"P={P} {W}x{H} start={start} metallic={metallic}: loss unfused {l0:.8f} fused {l1:.8f} rel {abs(l0-l1)/abs(l0):.2e}")
    o = 0
    names = list(p0.leaves.keys()) + ["light_diffuse"] + [f"light_spec{i}" for i in range(len(p0.light_leaves)-1)]
    for nm, t in zip(names, list(p0.leaves.values()) + p0.light_leaves):
        n = t.numel()
        a, b = g0[o:o+n], g1[o:o+n]
        o += n
        na = a.norm().item()
        if na == 0 and b.norm().item() == 0: continue
        print(f"   grad {nm:14s} |ref| {na:.4e} rel_l2 {((a-b).norm()/max(na,1e-30)).item():.3e} max_abs {(a-b).abs().max().item():.3e}")
    # maps
    ws = p1.last_workspace
    g = p0.activated()
    with torch.no_grad():
        res = renderer.pbr_forward(cam, g, p0.light(), lut, rays, bg, gi=gi, metallic=metallic)
    for nm, ref in (("render_rgb", res["render_rgb"]), ("render_direct", res["render_direct"]), ("albedo", res["albedo_map"]),
                    ("shade_normal", res["normal_map"]), ("ssr_normal", res["out_normal_view"]), ("occlusion", res["occlusion_map"]),
                    ("depth_pos", res["depth_pos"])):
        d = (ws.map(nm) - ref).abs()
        print(f"   map {nm:14s} max_abs {d.max().item():.3e} n>1e-4 {(d>1e-4).sum().item()} / {d.numel()}")

run(20000, 400, 300, 8, base_res=64)
run(20000, 400, 300, 8, base_res=64, metallic=False)
run(5000, 333, 257, 64, base_res=32)
run(300000, 800, 800, 64)
# timing
raw = scene.make_scene(300000, seed=0, regime="trained")
params = gstep.GaussianParams(raw, dev, light=scene.make_light(0))
lut = shade.make_brdf_lut().to(dev); cam = scene.orbit_camera(0, 8, 800, 800).to(dev)
rays = scene.canonical_rays(cam, dev); gt = torch.rand(3, 800, 800, device=dev); bg = torch.zeros(3, device=dev)
gi = dict(radius=0.8, bias=0.01, thick=0.05, delta=0.0625, step=16, start=64)
for fused in (False, True):
    for i in range(5):
        params.zero_grad(); gstep.training_step(params, cam, params.light(), lut, rays, gt, bg, gi, fused=fused)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(20):
        params.zero_grad(); gstep.training_step(params, cam, params.light(), lut, rays, gt, bg, gi, fused=fused)
    t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"fused={fused}: cpu issue {(t1-t0)*50:.3f} ms/step wall {(t2-t0)*50:.3f} ms/step")
