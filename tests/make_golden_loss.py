"""Writes tests/golden/loss_ref.npz from the reference's OWN utils/loss_utils.py (ssim, l1_loss; imported from
/root/reference, CPU float32): loss values and autograd gradients of the first-stage image loss
(1 - lambda) * L1 + lambda * (1 - SSIM) (train.py:320-322) on seeded images, including a non-multiple-of-16 size, a
single-channel image and an image smaller than the window. Run in the build container: python tests/make_golden_loss.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def images(Cn, H, W, seed):
    g = torch.Generator().manual_seed(seed)
    gt = torch.rand(Cn, H, W, generator=g)
    # a "render" correlated with the ground truth (SSIM far from 0) with a few exact matches (L1's sign(0) = 0)
    img = (gt + 0.15 * torch.randn(Cn, H, W, generator=g)).clamp(0, 1)
    img[:, ::7, ::5] = gt[:, ::7, ::5]
    return img, gt


CASES = [(3, 37, 53, 1), (3, 64, 48, 2), (1, 16, 16, 3), (3, 7, 9, 4), (5, 33, 20, 5)]
NCASES = [(41, 29, 6, 0.6), (16, 48, 7, 1.0), (9, 5, 8, 0.3)]      # (H, W, seed, fraction of pixels in the mask)


def normal_inputs(H, W, seed, frac):
    """View-space normal map, normal-from-depth map, its mask and a ground-truth image for the geometry terms."""
    g = torch.Generator().manual_seed(seed)
    import torch.nn.functional as F
    nd = F.normalize(torch.randn(3, H, W, generator=g), dim=0)
    nm = F.normalize(nd + 0.3 * torch.randn(3, H, W, generator=g), dim=0)
    mask = torch.rand(H, W, generator=g) < frac
    nd = nd * mask                                   # normal_from_depth is zero outside its mask
    nm[:, ::4, ::3] = nd[:, ::4, ::3]                # exact matches: sign(0) = 0
    gt = torch.rand(3, H, W, generator=g)
    return nm, nd, mask, gt


def reference_function(name):
    """The reference's train.py cannot be imported here (kornia / nvdiffrast are absent): compile the one function
    from its source file instead."""
    import ast
    import torch.nn.functional as F
    src = open("/root/reference/train.py").read()
    node = next(n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == name)
    ns = {"torch": torch, "F": F}
    exec(compile(ast.Module(body=[node], type_ignores=[]), "/root/reference/train.py", "exec"), ns)
    return ns[name]


def main():
    sys.path.insert(0, "/root/reference")
    from utils.loss_utils import l1_loss, ssim
    out = {"cases": np.array(CASES)}
    for i, (Cn, H, W, seed) in enumerate(CASES):
        img, gt = images(Cn, H, W, seed)
        x = img.clone().requires_grad_(True)
        s = ssim(x, gt)
        l1 = l1_loss(x, gt)
        loss = (1.0 - 0.2) * l1 + 0.2 * (1.0 - s)
        loss.backward()
        xs = img.clone().requires_grad_(True)
        ssim(xs, gt).backward()
        out[f"loss{i}"] = np.array([float(loss), float(l1), float(s)], dtype=np.float64)
        out[f"grad{i}"] = x.grad.numpy()
        out[f"grad_ssim{i}"] = xs.grad.numpy()
    # geometry terms of the first stage (train.py:323-328) from the reference's own get_tv_loss
    import torch.nn.functional as F
    get_tv_loss = reference_function("get_tv_loss")
    for i, (H, W, seed, frac) in enumerate(NCASES):
        nm, nd, mask, gt = normal_inputs(H, W, seed, frac)
        x = nm.clone().requires_grad_(True)
        nl = F.l1_loss(x[:, mask], nd[:, mask])
        tv = get_tv_loss(gt, x, pad=1, step=1)
        (1.0 * nl + 1.0 * tv).backward()
        out[f"nloss{i}"] = np.array([float(nl + tv), float(nl), float(tv)], dtype=np.float64)
        out[f"ngrad{i}"] = x.grad.numpy()
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "loss_ref.npz"), **out)
    print("wrote loss_ref.npz")


if __name__ == "__main__":
    main()
