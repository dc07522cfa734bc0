"""Generate tests/golden/cubemap_ref.npz by running the REFERENCE's own cubemap-filter kernels
(pbr/renderutils/c_src/cubemap.cu through oracle/_ref/libgigs_ref_cubemap.so) on a B200:

    gpurun -- 'python tests/make_golden_cubemap.py'   # writes gpurun_out/golden/cubemap_ref.npz; copy to tests/golden/

Inputs are seeded (CPU generator) and small; stored: the inputs' seeds, cutoffs, bounds, forward textures, wsum and
backward gradients of each case. Pins the CPU oracle (tests/test_oracle_golden.py::test_cubemap_*) and is re-checked
against our kernels on the GPU (tests/test_gpu_cubemap.py covers the live comparison)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

CASES = [  # (name, res, roughness)
    ("r16_rough1", 16, 1.0), ("r16_rough05", 16, 0.5), ("r32_rough029", 32, 0.29), ("r64_rough008", 64, 0.08)]


def cube(res, seed):
    return torch.rand(6, res, res, 3, generator=torch.Generator().manual_seed(seed)) * 0.5 + 0.25


def grad(res, seed):
    return torch.randn(6, res, res, 3, generator=torch.Generator().manual_seed(seed))


def main():
    import gigs_oracle as O
    import refshim
    ref = refshim.RefCubemap()
    dev = torch.device("cuda:0")
    out = {}
    for name, res, rough in CASES:
        c = O.ndf_cutoff(rough, 0.99)
        x, g = cube(res, 100 + res).to(dev), grad(res, 200 + res).to(dev)
        o, w, b = ref.specular_cubemap(x, rough, 0.99, c)
        gi = ref.specular_cubemap_grad(x, b, w, g, rough, c)
        out[f"{name}.cutoff"] = np.float64(c)
        out[f"{name}.bounds"] = b.cpu().numpy().astype(np.int16)
        out[f"{name}.out"] = o.cpu().numpy()
        out[f"{name}.wsum"] = w.cpu().numpy()
        out[f"{name}.grad_in"] = gi.cpu().numpy()
    x, g = cube(16, 116).to(dev), grad(16, 216).to(dev)
    out["diffuse16.out"] = ref.diffuse_cubemap_fwd(x).cpu().numpy()
    out["diffuse16.grad_in"] = ref.diffuse_cubemap_bwd(x, g).cpu().numpy()
    outdir = os.path.join(ROOT, "gpurun_out", "golden")
    os.makedirs(outdir, exist_ok=True)
    path = os.path.join(outdir, "cubemap_ref.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
