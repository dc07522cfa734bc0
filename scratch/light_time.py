"""Times PrefilteredLight.build / backward at base_res 256 (CUDA events, L2 flushed between iterations)."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gi-gs_b200"))
import torch
from gigs import light as GL, _lib
import ctypes as C
L = _lib.load()
dev = torch.device("cuda:0")
base = torch.rand(6, 256, 256, 3, device=dev) * 0.5 + 0.25
fl = GL.PrefilteredLight(base)
fc = GL.PrefilteredLight(base, stored_operators=False)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
gb = torch.zeros_like(base)
def run(fn, n=10):
    ts = []
    for _ in range(3):
        fn()
    for _ in range(n):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]
res = dict(build_ms=run(fl.build), backward_ms=run(lambda: fl.backward(gb, accumulate=False)),
           compute_build_ms=run(fc.build), compute_backward_ms=run(lambda: fc.backward(gb, accumulate=False)),
           weights_MB=fl.layout.weights_bytes / 1e6, n_weights=list(fl.layout.n_weights), n_runs=list(fl.layout.n_runs), lanes_log2=list(fl.layout.lanes_log2))
print(json.dumps(res)); sys.exit(0)
# per-level timings through the single-op entry points
for lvl in range(fl.layout.n_levels):
    r = fl.layout.res[lvl]; rough = fl.layout.roughness[lvl]
    x = torch.rand(6, r, r, 3, device=dev).requires_grad_(True)
    c, b = GL.specular_bounds(r, float(rough), 0.99, dev) if False else (fl.layout.cutoff[lvl], None)
    key_r = 1.0 if lvl == fl.layout.n_levels - 1 else (lvl / (fl.layout.n_levels - 2)) * 0.42 + 0.08
    y = GL.specular_cubemap(x, key_r)
    g = torch.randn_like(y)
    res[f"spec{r}_fwd_ms"] = run(lambda: GL.specular_cubemap(x.detach(), key_r))
    y = GL.specular_cubemap(x, key_r)
    res[f"spec{r}_bwd_ms"] = run(lambda: torch.autograd.grad(y, x, g, retain_graph=True))
x = torch.rand(6, 16, 16, 3, device=dev)
res["diffuse_fwd_ms"] = run(lambda: GL.diffuse_cubemap(x))
print(json.dumps(res))
