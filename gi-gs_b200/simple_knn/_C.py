"""Drop-in for `simple_knn._C` (reference submodules/simple-knn/ext.cpp: `distCUDA2`)."""
import ctypes as C
import os
import sys

import torch

_PKG_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _PKG_ROOT not in sys.path:
    sys.path.insert(0, _PKG_ROOT)
from gigs import _lib  # noqa: E402

_L = _lib.load()


def distCUDA2(points: torch.Tensor) -> torch.Tensor:
    """Mean squared distance to the 3 nearest other points, [P,3] -> [P] (spatial.cu:14-24)."""
    if not points.is_cuda:
        raise RuntimeError("points must be a CUDA tensor")
    P = points.size(0)
    pts = points.contiguous().float()
    means = torch.zeros((P,), dtype=torch.float32, device=points.device)
    need = C.c_uint64(0)
    _lib.check(_L.gigs_dist2(P, None, None, None, C.byref(need), None), "gigs_dist2(size)")
    scratch = torch.empty((need.value,), dtype=torch.uint8, device=points.device)
    with torch.cuda.device(points.device):
        _lib.check(_L.gigs_dist2(P, _lib.ptr(pts), means.data_ptr(), scratch.data_ptr(), C.byref(need),
                                 torch.cuda.current_stream().cuda_stream), "gigs_dist2")
    return means
