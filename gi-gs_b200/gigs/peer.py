"""Host side of gigs_peer_allreduce (csrc/peer_reduce.cu): the gradient buffer of the view-sharded training step lives
in a symmetric allocation that every rank of the node maps into its own address space
(torch.distributed._symmetric_memory: CUDA IPC / fabric handles exchanged through the process group's store), so that
the exchange step is one kernel of ours over NVLink loads and stores instead of NCCL collectives.
torch.distributed stays the plumbing (rendezvous, the fallback collective when peer mapping is unavailable)."""
import ctypes as C
import os
from typing import List, Optional, Tuple

import torch

from . import _lib


class PeerBuffer:
    """A float32 buffer of `numel` elements per rank, peer-mapped, plus the flag block of the barrier protocol."""

    N_CTAS = 0          # 0: the kernel scales its grid with the bytes of the call

    def __init__(self, numel: int, device, group=None):
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm
        group = group or dist.group.WORLD
        self.buf = symm.empty(numel, dtype=torch.float32, device=device)
        self.buf.zero_()
        self._h = symm.rendezvous(self.buf, group)
        self.world, self.rank = int(self._h.world_size), int(self._h.rank)
        if self.world > 16:
            raise RuntimeError("gigs.peer: at most 16 ranks per node")
        self.flags = symm.empty(64, dtype=torch.int32, device=device)
        self.flags.zero_()
        self._hf = symm.rendezvous(self.flags, group)
        self._bufs = (C.c_uint64 * self.world)(*[int(p) for p in self._h.buffer_ptrs])
        self._flags = (C.c_uint64 * self.world)(*[int(p) for p in self._hf.buffer_ptrs])
        # NVLS: the allocator also maps the buffer through the NVSwitch multicast object when the fabric has one
        self.multicast_ptr = 0
        # (measured on 2 B200s the switch path is the slower one — the traffic is the same and multimem loads issue
        # more slowly than plain peer loads; it pays from 4 ranks on, where it divides the link traffic by ~N/2)
        want = os.environ.get("GIGS_PEER_NVLS", "auto")
        if want == "1" or (want == "auto" and self.world >= 4):
            try:
                self.multicast_ptr = int(getattr(self._h, "multicast_ptr", 0) or 0)
            except Exception:
                self.multicast_ptr = 0
        self.epoch = 0
        torch.cuda.synchronize(device)
        dist.barrier(group)          # every rank's buffers are zeroed before anybody's first call can touch them

    def all_reduce(self, spans: Optional[List[Tuple[int, int]]] = None) -> None:
        """Sum the given [begin, end) float spans (default: the whole buffer) over all ranks, in place, on the current
        stream. Same spans, same call sequence on every rank."""
        L = _lib.load()
        spans = spans if spans is not None else [(0, self.buf.numel())]
        spans = [(int(a), int(b)) for a, b in spans if b > a]
        if not spans:
            return
        for i in range(0, len(spans), 16):
            part = spans[i:i + 16]
            lo = (C.c_uint64 * len(part))(*[a for a, _ in part])
            hi = (C.c_uint64 * len(part))(*[b for _, b in part])
            self.epoch += 1
            with torch.cuda.device(self.buf.device):
                _lib.check(L.gigs_peer_allreduce(self.world, self.rank, self._bufs, self._flags, self.multicast_ptr,
                                                 self.epoch, len(part), lo, hi, self.N_CTAS,
                                                 torch.cuda.current_stream().cuda_stream),
                           "gigs_peer_allreduce")

    def error_epoch(self) -> int:
        """0, or the epoch of the call in which a peer did not arrive within the (about one minute) timeout. That call's
        kernel traps, so in practice the process learns of it as a CUDA error on its next call; the word is for a
        post-mortem (synchronises the device)."""
        return int(self.flags[2 * self.world + 1].item())


def available() -> bool:
    """Peer mapping needs a process group with more than one rank on CUDA devices and the symmetric-memory allocator."""
    if os.environ.get("GIGS_PEER_AR", "1") == "0":
        return False
    try:
        import torch.distributed as dist
        import torch.distributed._symmetric_memory  # noqa: F401
        return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1 and torch.cuda.is_available()
    except Exception:
        return False
