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
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "loss_ref.npz"), **out)
    print("wrote loss_ref.npz")


if __name__ == "__main__":
    main()
