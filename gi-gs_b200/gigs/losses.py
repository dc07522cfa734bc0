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


class _NormalLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, normal_map, normal_from_depth, mask, gt, normal_weight, tv_weight, loss_scale):
        L = _lib.load()
        _, H, W = normal_map.shape
        need = C.c_uint64(0)
        _lib.check(L.gigs_normal_loss(W, H, None, None, None, None, normal_weight, tv_weight, loss_scale, None,
                                      C.byref(need), None, 0, None, 0, None, None), "gigs_normal_loss(size)")
        scratch = _scratch_for(normal_map.device, need.value)
        out = torch.empty(3, dtype=torch.float32, device=normal_map.device)
        want_grad = normal_map.requires_grad
        grad = torch.empty_like(normal_map) if want_grad else None
        with torch.cuda.device(normal_map.device):
            _lib.check(L.gigs_normal_loss(W, H, normal_map.data_ptr(), normal_from_depth.data_ptr(),
                                          mask.data_ptr() if mask is not None else None, gt.data_ptr(), normal_weight,
                                          tv_weight, loss_scale, scratch.data_ptr(), C.byref(need), out.data_ptr(), 0,
                                          grad.data_ptr() if want_grad else None, 0, None,
                                          torch.cuda.current_stream().cuda_stream), "gigs_normal_loss")
        ctx.grad = grad
        ctx.mark_non_differentiable(out)
        return out[0], out

    @staticmethod
    def backward(ctx, g_loss, _g_terms):
        return (ctx.grad * g_loss if ctx.grad is not None else None), None, None, None, None, None, None


def normal_loss(normal_map: torch.Tensor, normal_from_depth: torch.Tensor, mask, gt_image: torch.Tensor,
                normal_weight: float = 1.0, tv_weight: float = 1.0, loss_scale: float = 1.0, return_terms: bool = False):
    """train.py:323-328: normal_weight * F.l1_loss(normal_map[:, mask], normal_map_from_depth[:, mask]) + tv_weight *
    get_tv_loss(gt_image, normal_map, pad=1, step=1); differentiable in normal_map. mask: bool/uint8 [H,W] or None."""
    normal_map, normal_from_depth = _check(normal_map, normal_from_depth.detach())
    gt_image = gt_image.detach().float().contiguous()
    if mask is not None:
        mask = mask.reshape(normal_map.shape[1:]).to(torch.uint8).contiguous()
    loss, terms = _NormalLoss.apply(normal_map, normal_from_depth, mask, gt_image, float(normal_weight),
                                    float(tv_weight), float(loss_scale))
    return (loss, terms) if return_terms else loss
