"""Host side of the gigs_b200 hot path (PyTorch is plumbing: device memory, streams, torch.distributed)."""
