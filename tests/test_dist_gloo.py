"""CPU, world_size 2, gloo: the view-sharded step's collective logic (BASELINE C4 / SURVEY §8e):
sum over ranks of per-rank view gradients, all-reduced in ONE collective over the flat buffer,
equals the single-process gradient of the K-view mean loss."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gigs import scene, step as gstep


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _view_loss(params, k, K):
    """A stand-in differentiable 'view' (no GPU here): a view-dependent function of every parameter group."""
    g = params.activated()
    w = 1.0 + 0.1 * k
    loss = (g["means3D"] * w).pow(2).mean() + (g["opacity"] * w).mean() + (g["scales"] * g["albedo"]).mean() * w
    loss = loss + g["shs"].mul(w).sin().mean() + g["rotations"][:, 0].mean() + g["normal"].abs().mean() * w
    loss = loss + (g["roughness"] * g["metallic"]).mean() * w
    for t in params.light_leaves:
        loss = loss + (t * w).mean()
    return loss / K


def _worker(rank, world, port, K, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    raw = scene.make_scene(50, seed=1)
    params = gstep.GaussianParams(raw, "cpu", light=scene.make_light(0, base_res=32))
    params.zero_grad()
    for k in gstep.shard_views(K, rank, world):
        _view_loss(params, k, K).backward()
    dist.all_reduce(params.flat_grad, op=dist.ReduceOp.SUM)
    if rank == 0:
        torch.save(params.flat_grad.clone(), out)
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_allreduce_equals_single_rank_sum(tmp_path):
    K, world = 8, 2
    out = str(tmp_path / "g.pt")
    mp.spawn(_worker, args=(world, _free_port(), K, out), nprocs=world, join=True)
    sharded = torch.load(out)
    raw = scene.make_scene(50, seed=1)
    params = gstep.GaussianParams(raw, "cpu", light=scene.make_light(0, base_res=32))
    params.zero_grad()
    for k in range(K):
        _view_loss(params, k, K).backward()
    ref = params.flat_grad
    assert sharded.shape == ref.shape
    assert float((sharded - ref).abs().max()) <= 1e-6 * float(ref.abs().max()) + 1e-9


def _worker_spans(rank, world, port, out):
    """The fused path's contract: only the material + light spans are written, zero_grad / all_reduce_grads with
    fused_only=True touch exactly those spans (what bench.py and multi_view_step do on N GPUs)."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    raw = scene.make_scene(40, seed=2)
    params = gstep.GaussianParams(raw, "cpu", light=scene.make_light(0, base_res=32))
    params.zero_grad()                       # unknown history -> full clear, tracking starts
    keys = ["albedo", "roughness", "metallic"] + [f"light{i}" for i in range(len(params.light_leaves))]
    params.mark_dirty(keys)
    with torch.no_grad():
        for k in ("albedo", "roughness", "metallic"):
            params.leaves[k].grad.add_(float(rank + 1))
        for t in params.light_leaves:
            t.grad.add_(0.5 * (rank + 1))
    params.all_reduce_grads(fused_only=True)
    if rank == 0:
        torch.save(dict(flat=params.flat_grad.clone(), spans=params._merged_dirty()), out)
    params.zero_grad(fused_only=True)
    assert float(params.flat_grad.abs().sum()) == 0.0
    dist.barrier()
    dist.destroy_process_group()


def test_span_limited_allreduce_and_zero(tmp_path):
    world = 2
    out = str(tmp_path / "s.pt")
    mp.spawn(_worker_spans, args=(world, _free_port(), out), nprocs=world, join=True)
    r = torch.load(out)
    raw = scene.make_scene(40, seed=2)
    params = gstep.GaussianParams(raw, "cpu", light=scene.make_light(0, base_res=32))
    assert len(r["spans"]) == 2              # [albedo|roughness|metallic] and the light textures: two collectives
    want = torch.zeros_like(params.flat_grad)
    for k in ("albedo", "roughness", "metallic"):
        lo, hi = params._span[k]
        want[lo:hi] = 3.0                    # (1) + (2)
    for i in range(len(params.light_leaves)):
        lo, hi = params._span[f"light{i}"]
        want[lo:hi] = 1.5
    assert torch.equal(r["flat"], want)
