import os, sys, time, torch, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gi-gs_b200"))
from gigs import scene, shade, step as gstep, _lib
L = _lib.load()
dev = torch.device("cuda:0")
raw = scene.make_scene(300000, seed=0, regime="trained")
lut = shade.make_brdf_lut().to(dev); cam = scene.orbit_camera(0, 8, 800, 800).to(dev)
rays = scene.canonical_rays(cam, dev); gt = torch.rand(3, 800, 800, device=dev); bg = torch.zeros(3, device=dev)
gi = dict(radius=0.8, bias=0.01, thick=0.05, delta=0.0625, step=16, start=64)
NAMES = ["preprocess", "emit_keys", "radix_sort", "tile_ranges", "blend_forward", "blend_backward", "gaussian_backward", "geometry_chain", "ssao", "ssr", "shade_forward", "shade_backward", "median3x3", "median3x3_backward", "bilateral3x3", "depth_to_normal", "ssr_backward", "dist2", "deferred_shade", "deferred_loss", "deferred_backward", "param_grad", "radix_sort_pass", "depth_sort"]
for with_light in ('all', 'none', 'diffuse_only', 'spec_only', 'spec_fine_only', 'spec_coarse_only'):
    params = gstep.GaussianParams(raw, dev, light=scene.make_light(0))
    light = params.light()
    sp = list(light.specular); df = light.diffuse
    if with_light in ('none', 'diffuse_only'): sp = [t.detach() for t in sp]
    if with_light in ('none', 'spec_only', 'spec_fine_only', 'spec_coarse_only'): df = df.detach()
    if with_light == 'spec_fine_only': sp = [t if i < 3 else t.detach() for i, t in enumerate(sp)]
    if with_light == 'spec_coarse_only': sp = [t if i >= 3 else t.detach() for i, t in enumerate(sp)]
    light = shade.Light(specular=sp, diffuse=df)
    for i in range(5):
        gstep.training_step(params, cam, light, lut, rays, gt, bg, gi)
    torch.cuda.synchronize()
    L.gigs_profile_enable(1)
    for i in range(10):
        gstep.training_step(params, cam, light, lut, rays, gt, bg, gi)
    torch.cuda.synchronize()
    st = (C.c_int32 * 4096)(); ms = (C.c_float * 4096)()
    n = L.gigs_profile_read(st, ms, 4096)
    L.gigs_profile_enable(0)
    agg = {}
    for j in range(n): agg[NAMES[st[j]]] = agg.get(NAMES[st[j]], 0) + ms[j] / 10
    print("light grads", with_light, {k: round(v, 4) for k, v in agg.items() if k in ("deferred_backward", "deferred_shade", "blend_backward")})
