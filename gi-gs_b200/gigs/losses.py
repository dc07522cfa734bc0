"""Loss-side image ops of the training loop behind the names of /root/reference/utils/loss_utils.py.

`ssim(img1, img2)` and `l1_ssim_loss(image, gt, lambda_dssim)` run the fused kernels of csrc/loss.cu through
gigs_image_loss (one forward + one backward launch instead of five grouped 11x11 convolutions, ~15 elementwise kernels
and their autograd replay). No framework fallback: CUDA tensors only.
"""
import ctypes as C

import torch

from . import _lib

_scratch = {}


def _scratch_for(dev, nbytes):
    t = _scratch.get(dev)
    if t is None or t.numel() < nbytes:
        t = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        _scratch[dev] = t
    return t


def _check(image, gt):
    if not (image.is_cuda and gt.is_cuda):
        raise RuntimeError("gigs.losses: image and gt must be CUDA tensors (no CPU fallback)")
    if image.dim() == 4 and image.shape[0] == 1:
        image, gt = image[0], gt[0]
    if image.dim() != 3 or image.shape != gt.shape:
        raise RuntimeError("gigs.losses: image and gt must both be [C,H,W]")
    return image.float().contiguous(), gt.float().contiguous()


class _ImageLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, image, gt, lambda_dssim, loss_scale):
        L = _lib.load()
        Cn, H, W = image.shape
        need = C.c_uint64(0)
        _lib.check(L.gigs_image_loss(Cn, W, H, None, None, lambda_dssim, loss_scale, None, C.byref(need), None, 0, None,
                                     0, None, None), "gigs_image_loss(size)")
        want_grad = image.requires_grad
        # the derivative maps must outlive the call until backward: own buffer when a gradient is wanted
        scratch = (torch.empty(need.value, dtype=torch.uint8, device=image.device) if want_grad
                   else _scratch_for(image.device, need.value))
        out = torch.empty(3, dtype=torch.float32, device=image.device)
        grad = torch.empty_like(image) if want_grad else None
        with torch.cuda.device(image.device):
            _lib.check(L.gigs_image_loss(Cn, W, H, image.data_ptr(), gt.data_ptr(), lambda_dssim, loss_scale,
                                         scratch.data_ptr(), C.byref(need), out.data_ptr(), 0,
                                         grad.data_ptr() if want_grad else None, 0, None,
                                         torch.cuda.current_stream().cuda_stream), "gigs_image_loss")
        ctx.grad = grad
        ctx.mark_non_differentiable(out)
        return out[0], out

    @staticmethod
    def backward(ctx, g_loss, _g_terms):
        if ctx.grad is None:
            return None, None, None, None
        return ctx.grad * g_loss, None, None, None


def l1_ssim_loss(image: torch.Tensor, gt: torch.Tensor, lambda_dssim: float = 0.2, loss_scale: float = 1.0,
                 return_terms: bool = False):
    """train.py:320-322: (1 - lambda_dssim) * l1_loss(image, gt) + lambda_dssim * (1 - ssim(image, gt)); differentiable
    in `image`. With return_terms also the (detached) [loss, l1, ssim] vector."""
    image, gt = _check(image, gt)
    loss, terms = _ImageLoss.apply(image, gt.detach(), float(lambda_dssim), float(loss_scale))
    return (loss, terms) if return_terms else loss


def ssim(img1: torch.Tensor, img2: torch.Tensor, window_size: int = 11, size_average: bool = True) -> torch.Tensor:
    """utils/loss_utils.py:54-100 for its only call pattern (window 11, size_average=True), differentiable in img1."""
    if window_size != 11 or not size_average:
        raise NotImplementedError("gigs.losses.ssim: the reference only ever calls ssim(img1, img2)")
    # lambda = 1, scale = -1: loss = -(1 - ssim)  =>  ssim = loss + 1
    return l1_ssim_loss(img1, img2, lambda_dssim=1.0, loss_scale=-1.0) + 1.0


def l1_loss(network_output: torch.Tensor, gt: torch.Tensor) -> torch.Tensor:
    """utils/loss_utils.py:19-20."""
    return torch.abs(network_output - gt).mean()
