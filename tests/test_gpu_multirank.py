"""GPU, 2 ranks, NCCL: the view-sharded K-view step (BASELINE C4) with the span-limited, overlapped all-reduce equals
the single-process K-view step. Skipped on boxes with one GPU (the driver's -m gpu tier); run with gpurun --gpus 2."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu

from gigs import scene, shade, step as gstep

GI = dict(radius=0.8, bias=0.01, thick=0.05, delta=0.0625, step=16, start=64)
P, W, H, K = 20000, 160, 128, 4


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _inputs(dev, base_light=False, peer=False):
    raw = scene.make_scene(P, seed=3)
    if base_light:   # the light as its trainable base cubemap: build_mips per step, its backward before the all-reduce
        base = torch.rand(6, 64, 64, 3, generator=torch.Generator().manual_seed(9)) * 0.5 + 0.25
        params = gstep.GaussianParams(raw, dev, light_base=base, peer=peer)
    else:
        params = gstep.GaussianParams(raw, dev, light=scene.make_light(0, base_res=64), peer=peer)
    lut = shade.make_brdf_lut(64, 64).to(dev)
    cams = [scene.orbit_camera(k, 8, W, H).to(dev) for k in range(K)]
    gen = torch.Generator().manual_seed(0)
    gts = [torch.rand(3, H, W, generator=gen).to(dev) for _ in range(K)]
    return params, lut, cams, gts, scene.canonical_rays(cams[0], dev), torch.zeros(3, device=dev)


def _kw(base_light):
    return dict(brdf_tv_weight=1.0, env_tv_weight=0.01) if base_light else {}


def _worker(rank, world, port, out, base_light, peer=False):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    params, lut, cams, gts, rays, bg = _inputs(dev, base_light, peer)
    if peer:
        assert params._peer is not None, "peer-mapped gradient buffer not available on this box"
    for _ in range(2):   # twice: the second step exercises zero_grad(fused_only) + a reused comm stream
        total = gstep.multi_view_step(params, cams, params.light(), lut, lambda c: rays, gts, bg, GI, rank=rank,
                                      world=world, **_kw(base_light))
    torch.cuda.synchronize()
    if peer:
        assert params._peer.error_epoch() == 0
        # the in-place two-shot sums are bit-identical on every rank
        mine = params.flat_grad.clone()
        other = mine.clone()
        dist.broadcast(other, src=0)
        assert torch.equal(mine, other)
    if rank == 0:
        torch.save(dict(grad=params.flat_grad.cpu(), total=float(total)), out)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("base_light,peer", [(False, False), (True, False), (False, True), (True, True)])
def test_two_rank_multi_view_step_matches_single_process(tmp_path, base_light, peer):
    """peer=True: the exchange is gigs_peer_allreduce (one kernel over NVLink peer memory) instead of NCCL."""
    out = str(tmp_path / "r.pt")
    mp.spawn(_worker, args=(2, _free_port(), out, base_light, peer), nprocs=2, join=True)
    r = torch.load(out)
    dev = torch.device("cuda", 0)
    params, lut, cams, gts, rays, bg = _inputs(dev, base_light)
    total = gstep.multi_view_step(params, cams, params.light(), lut, lambda c: rays, gts, bg, GI, **_kw(base_light))
    ref = params.flat_grad.cpu()
    assert abs(r["total"] - float(total)) <= 1e-6 * abs(float(total))
    rel = float((r["grad"] - ref).norm() / ref.norm())
    assert rel <= 1e-4, rel


def _peer_worker(rank, world, port):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from gigs import peer
    n = 1_000_003                      # odd size: unaligned spans take the scalar path
    pb = peer.PeerBuffer(n, dev)
    g = torch.Generator().manual_seed(100 + rank)
    for it, spans in enumerate(([(0, n)], [(0, 4096), (5000, 5003), (65536, 900000)], [(3, 1001)], [(0, n)])):
        x = torch.randn(n, generator=g).to(dev)
        pb.buf.copy_(x)
        want = x.clone()
        dist.all_reduce(want)          # NCCL, for comparison
        pb.all_reduce(spans)
        torch.cuda.synchronize()
        assert pb.error_epoch() == 0
        for lo, hi in spans:
            assert torch.allclose(pb.buf[lo:hi], want[lo:hi], rtol=1e-6, atol=1e-6), (it, lo, hi)
        mask = torch.ones(n, dtype=torch.bool, device=dev)
        for lo, hi in spans:
            mask[lo:hi] = False
        assert torch.equal(pb.buf[mask], x[mask])          # outside the spans nothing moved
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_peer_allreduce_matches_nccl_on_arbitrary_spans():
    mp.spawn(_peer_worker, args=(2, _free_port()), nprocs=2, join=True)


def _stage1_worker(rank, world, port, out, peer):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from gigs import densify
    params, lut, cams, gts, rays, bg = _inputs(dev, False, peer)
    st = densify.DensifyState(P, dev)
    total = gstep.multi_view_first_stage_step(params, cams, gts, bg, GI, rank=rank, world=world, stats=st)
    gstep.reduce_densify_stats(st, world)
    torch.cuda.synchronize()
    if rank == 0:
        torch.save(dict(grad=params.flat_grad.cpu(), total=float(total), accum=st.xyz_gradient_accum.cpu(),
                        denom=st.denom.cpu(), radii=st.max_radii2D.cpu()), out)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("peer", [False, True])
def test_two_rank_first_stage_step_matches_single_process(tmp_path, peer):
    """First stage: every group has gradients, the whole 268 B/Gaussian buffer is exchanged (peer kernel or NCCL), and the
    per-view densification statistics are reduced over the ranks."""
    from gigs import densify
    out = str(tmp_path / "s1.pt")
    mp.spawn(_stage1_worker, args=(2, _free_port(), out, peer), nprocs=2, join=True)
    r = torch.load(out)
    dev = torch.device("cuda", 0)
    params, lut, cams, gts, rays, bg = _inputs(dev)
    st = densify.DensifyState(P, dev)
    total = gstep.multi_view_first_stage_step(params, cams, gts, bg, GI, stats=st)
    ref = params.flat_grad.cpu()
    assert abs(r["total"] - float(total)) <= 1e-6 * abs(float(total))
    for k in gstep.PARAM_KEYS:
        lo, hi = params._span[k]
        if float(ref[lo:hi].norm()) > 0:
            assert float((r["grad"][lo:hi] - ref[lo:hi]).norm() / ref[lo:hi].norm()) <= 1e-4, k
    assert torch.allclose(r["accum"], st.xyz_gradient_accum.cpu(), rtol=1e-4, atol=1e-9)
    assert torch.equal(r["denom"], st.denom.cpu()) and torch.equal(r["radii"], st.max_radii2D.cpu())


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs in one process")
def test_second_device_in_the_same_process():
    """cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device attribute: the frame on cuda:1 after cuda:0 in ONE
    process (round 1 kept one process-wide flag per kernel and would have launched with the 48 KB default there)."""
    import sys, os
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gi-gs_b200"))
    from gigs import scene, shade, step as gstep
    P, W, H = 5000, 160, 128
    raw = scene.make_scene(P, seed=1, regime="trained")
    gi = dict(radius=0.8, bias=0.01, thick=0.05, delta=0.0625, step=16, start=8)
    out = []
    for d in (0, 1, 0):
        dev = torch.device("cuda", d)
        torch.cuda.set_device(dev)
        params = gstep.GaussianParams(raw, dev, light=scene.make_light(0, base_res=64))
        cam = scene.orbit_camera(1, 8, W, H).to(dev)
        lut = shade.make_brdf_lut(64, 64).to(dev)
        rays = scene.canonical_rays(cam, dev)
        gt = torch.rand(3, H, W, generator=torch.Generator().manual_seed(3)).to(dev)
        loss = gstep.training_step(params, cam, params.light(), lut, rays, gt, torch.zeros(3, device=dev), gi)
        torch.cuda.synchronize(dev)
        out.append((float(loss), params.last_workspace.map("render_rgb").cpu(), params.flat_grad.cpu()))
    torch.cuda.set_device(0)
    assert out[0][0] == out[1][0] == out[2][0]
    assert torch.equal(out[0][1], out[1][1]) and torch.equal(out[0][1], out[2][1])
    assert float((out[0][2] - out[1][2]).norm()) <= 1e-4 * float(out[0][2].norm())
