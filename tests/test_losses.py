"""First-stage image loss: (1 - lambda) * L1 + lambda * (1 - SSIM), train.py:320-322.
CPU: the oracle restatement (dense 11x11 window, no conv2d) against values AND autograd gradients of the reference's own
utils/loss_utils.py (tests/golden/loss_ref.npz). GPU: gigs_image_loss through the C-ABI against the same goldens, the
oracle at another size, and size-independent properties at 800x800."""
import os

import numpy as np
import pytest
import torch

import gigs_oracle as O
from make_golden_loss import CASES, images

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "loss_ref.npz")
# float32 sums of 121 products in a different order than the reference's conv2d: values to 2e-6; a gradient element is
# a sum of 121 terms of size ~lambda/N/sigma^2, compared at the scale of the largest element
VAL_TOL, GRAD_TOL = 2e-6, 2e-5


def _grad_close(a, b, tol=GRAD_TOL):
    scale = float(b.abs().max())
    return float((a - b).abs().max()) <= tol * scale


def test_oracle_matches_reference_loss_utils():
    z = np.load(GOLD)
    for i, (Cn, H, W, seed) in enumerate(CASES):
        img, gt = images(Cn, H, W, seed)
        x = img.clone().requires_grad_(True)
        loss = O.l1_ssim_loss(x, gt, 0.2)
        loss.backward()
        want = z[f"loss{i}"]
        assert float(loss) == pytest.approx(want[0], abs=VAL_TOL)
        assert float(O.ssim(img, gt)) == pytest.approx(want[2], abs=VAL_TOL)
        assert _grad_close(x.grad, torch.from_numpy(z[f"grad{i}"])), i


def test_ssim_window_is_the_reference_window():
    w = O.ssim_window()
    assert w.shape == (11, 11) and float(w.sum()) == pytest.approx(1.0, abs=1e-6)
    assert float(w[5, 5]) == pytest.approx(0.070766, abs=1e-5)     # (1/sum)^2 at the centre for sigma 1.5


# ------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
def test_image_loss_matches_reference_goldens():
    from gigs import losses
    z = np.load(GOLD)
    for i, (Cn, H, W, seed) in enumerate(CASES):
        img, gt = images(Cn, H, W, seed)
        x = img.cuda().requires_grad_(True)
        loss, terms = losses.l1_ssim_loss(x, gt.cuda(), 0.2, return_terms=True)
        loss.backward()
        want = z[f"loss{i}"]
        assert float(loss) == pytest.approx(want[0], abs=VAL_TOL), i
        assert float(terms[1]) == pytest.approx(want[1], abs=VAL_TOL) and float(terms[2]) == pytest.approx(want[2], abs=VAL_TOL)
        assert _grad_close(x.grad.cpu(), torch.from_numpy(z[f"grad{i}"])), i
        xs = img.cuda().requires_grad_(True)
        s = losses.ssim(xs, gt.cuda())
        (3.0 * s).backward()                                      # upstream gradient != 1
        assert float(s) == pytest.approx(want[2], abs=VAL_TOL)
        assert _grad_close(xs.grad.cpu() / 3.0, torch.from_numpy(z[f"grad_ssim{i}"])), i


@pytest.mark.gpu
def test_image_loss_matches_oracle_and_accumulates():
    import ctypes as C
    from gigs import _lib, losses
    img, gt = images(3, 75, 131, 9)
    x = img.clone().requires_grad_(True)
    want = O.l1_ssim_loss(x, gt, 0.35)
    want.backward()
    xg = img.cuda().requires_grad_(True)
    got = losses.l1_ssim_loss(xg, gt.cuda(), 0.35, loss_scale=0.5)
    got.backward()
    assert float(got) == pytest.approx(0.5 * float(want), abs=VAL_TOL)
    assert _grad_close(xg.grad.cpu(), 0.5 * x.grad)
    # raw C-ABI: accumulate into an existing loss / gradient, forward-only call without gradient maps, argument errors
    L = _lib.load()
    a, b = img.cuda().contiguous(), gt.cuda().contiguous()
    need = C.c_uint64(0)
    assert L.gigs_image_loss(3, 131, 75, None, None, 0.35, 1.0, None, C.byref(need), None, 0, None, 0, None, None) == 0
    scratch = torch.empty(need.value, dtype=torch.uint8, device="cuda")
    out = torch.full((3,), 2.0, device="cuda")
    g = torch.full_like(a, 1e-5)            # of the gradient's own size, so that subtracting it back is exact enough
    st = torch.cuda.current_stream().cuda_stream
    assert L.gigs_image_loss(3, 131, 75, a.data_ptr(), b.data_ptr(), 0.35, 1.0, scratch.data_ptr(), C.byref(need),
                             out.data_ptr(), 1, g.data_ptr(), 1, None, st) == 0
    assert float(out[0]) == pytest.approx(2.0 + float(want), abs=VAL_TOL)
    assert _grad_close(g.cpu() - 1e-5, x.grad, 5e-5)
    out2 = torch.zeros(3, device="cuda")
    assert L.gigs_image_loss(3, 131, 75, a.data_ptr(), b.data_ptr(), 0.35, 1.0, scratch.data_ptr(), C.byref(need),
                             out2.data_ptr(), 0, None, 0, None, st) == 0
    assert float(out2[0]) == pytest.approx(float(want), abs=VAL_TOL)
    small = C.c_uint64(16)
    assert L.gigs_image_loss(3, 131, 75, a.data_ptr(), b.data_ptr(), 0.35, 1.0, scratch.data_ptr(), C.byref(small),
                             out2.data_ptr(), 0, None, 0, None, st) < 0
    assert L.gigs_image_loss(0, 131, 75, None, None, 0.35, 1.0, None, C.byref(need), None, 0, None, 0, None, None) < 0
    with pytest.raises(RuntimeError, match="CUDA"):
        losses.ssim(img, gt)


@pytest.mark.gpu
def test_image_loss_properties_at_full_size():
    from gigs import losses
    g = torch.Generator().manual_seed(1)
    gt = torch.rand(3, 800, 800, generator=g).cuda()
    img = (gt + 0.1 * torch.randn(3, 800, 800, generator=g).cuda()).clamp(0, 1).requires_grad_(True)
    assert float(losses.ssim(gt, gt)) == pytest.approx(1.0, abs=1e-6)             # identity
    s_ab, s_ba = float(losses.ssim(img.detach(), gt)), float(losses.ssim(gt, img.detach()))
    assert s_ab == pytest.approx(s_ba, abs=1e-6)                                   # symmetry
    a = losses.l1_ssim_loss(img, gt, 0.2)
    b = losses.l1_ssim_loss(img, gt, 0.2)
    assert float(a) == float(b)                                                    # deterministic reduction
    # directional derivative: (loss(img + h d) - loss(img - h d)) / 2h ~ <grad, d>  (SSIM only: smooth)
    d = torch.randn(3, 800, 800, generator=g).cuda()
    h = 1e-2
    lp = float(losses.l1_ssim_loss((img.detach() + h * d), gt, 1.0))
    lm = float(losses.l1_ssim_loss((img.detach() - h * d), gt, 1.0))
    x2 = img.detach().clone().requires_grad_(True)
    losses.l1_ssim_loss(x2, gt, 1.0).backward()
    lin = float((x2.grad * d).sum())
    assert (lp - lm) / (2 * h) == pytest.approx(lin, rel=2e-2)


# ---- geometry terms of the first stage: normal L1 inside the mask + edge-aware TV of the normal map ----------------
from make_golden_loss import NCASES, normal_inputs


def test_oracle_normal_loss_matches_reference_functions():
    z = np.load(GOLD)
    for i, (H, W, seed, frac) in enumerate(NCASES):
        nm, nd, mask, gt = normal_inputs(H, W, seed, frac)
        x = nm.clone().requires_grad_(True)
        loss = O.normal_loss(x, nd, mask, gt)
        loss.backward()
        assert float(loss) == pytest.approx(z[f"nloss{i}"][0], rel=2e-6)
        assert _grad_close(x.grad, torch.from_numpy(z[f"ngrad{i}"]), 2e-6), i


@pytest.mark.gpu
def test_normal_loss_matches_reference_goldens_and_oracle():
    from gigs import losses
    z = np.load(GOLD)
    for i, (H, W, seed, frac) in enumerate(NCASES):
        nm, nd, mask, gt = normal_inputs(H, W, seed, frac)
        x = nm.cuda().requires_grad_(True)
        loss, terms = losses.normal_loss(x, nd.cuda(), mask.cuda(), gt.cuda(), return_terms=True)
        (2.0 * loss).backward()
        want = z[f"nloss{i}"]
        assert float(loss) == pytest.approx(want[0], rel=5e-6), i
        assert float(terms[1]) == pytest.approx(want[1], rel=5e-6) and float(terms[2]) == pytest.approx(want[2], rel=5e-6)
        assert _grad_close(x.grad.cpu() / 2.0, torch.from_numpy(z[f"ngrad{i}"]), 5e-6), i
    # other weights, no mask, full size: against the oracle
    nm, nd, mask, gt = normal_inputs(800, 800, 12, 0.7)
    x = nm.clone().requires_grad_(True)
    want = O.normal_loss(x, nd, torch.ones_like(mask), gt, normal_weight=0.5, tv_weight=2.0)
    want.backward()
    xg = nm.cuda().requires_grad_(True)
    got = losses.normal_loss(xg, nd.cuda(), None, gt.cuda(), normal_weight=0.5, tv_weight=2.0)
    got.backward()
    assert float(got) == pytest.approx(float(want), rel=5e-6)
    assert _grad_close(xg.grad.cpu(), x.grad, 5e-6)
    # an empty mask is a mean over nothing: NaN, as in the reference
    e = losses.normal_loss(nm.cuda(), nd.cuda(), torch.zeros_like(mask).cuda(), gt.cuda())
    assert torch.isnan(e)


@pytest.mark.gpu
@pytest.mark.parametrize("Cn,H,W", [(3, 5, 300), (1, 1, 17), (3, 17, 1), (2, 31, 33)])
def test_image_loss_odd_shapes_match_oracle(Cn, H, W):
    """Images thinner than the 11-tap window, single rows / columns, sizes that are not multiples of the 16-pixel tile."""
    from gigs import losses
    img, gt = images(Cn, H, W, 40 + H)
    x = img.clone().requires_grad_(True)
    want = O.l1_ssim_loss(x, gt, 0.2)
    want.backward()
    xg = img.cuda().requires_grad_(True)
    got = losses.l1_ssim_loss(xg, gt.cuda(), 0.2)
    got.backward()
    assert float(got) == pytest.approx(float(want), abs=VAL_TOL)
    assert _grad_close(xg.grad.cpu(), x.grad)
