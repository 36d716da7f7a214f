"""Host-side logic of the multi-GPU modes (supnerf_b200/parallel.py), on CPU with the gloo backend and world_size 2:
shard arithmetic, the global-denominator partial loss, the single flat all_reduce of pose/latent gradients, and the
ray all_gather.  The CPU oracle stands in for the CUDA render (the kernels cannot run here); the N>1 GPU path itself is
exercised by tests/test_gpu_multi.py and bench.py --gpus N on the B200 box."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT  # noqa: F401  (sys.path)
from oracle import oracle
from supnerf_b200 import parallel

IM, S = 16, 16


def test_ray_shards_tile_aligned_and_cover():
    for n, s in ((262144, 128), (4096, 64), (1000, 64), (1024, 37), (6, 64), (0, 64)):
        for world in (1, 2, 3, 4, 8):
            spans = [parallel.ray_shard(n, s, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for (a0, b0), (a1, b1) in zip(spans, spans[1:]):
                assert b0 == a1 and a0 <= b0
            for a, b in spans[:-1]:
                assert (a * s) % 128 == 0 and (b * s) % 128 == 0 or b == n
    with pytest.raises(ValueError):
        parallel.ray_shard(10, 64, 2, 2)


def test_interleaved_ray_shards_partition_and_balance():
    """parallel.ray_shard_indices: every ray in exactly one shard, whole 128-ray tiles (=> whole decoder tiles for any S), tile t
    on rank t % G, shard sizes within one tile of each other."""
    for n in (262144, 4096, 1000, 130, 6, 0):
        for world in (1, 2, 3, 4, 8):
            ids = [parallel.ray_shard_indices(n, 64, r, world) for r in range(world)]
            assert sorted(torch.cat(ids).tolist()) == list(range(n))
            for r, i in enumerate(ids):
                assert torch.equal(i, i.sort().values)
                assert torch.all((i // 128) % world == r)
            sizes = [int(i.numel()) for i in ids]
            assert max(sizes) - min(sizes) <= 128
    with pytest.raises(ValueError):
        parallel.ray_shard_indices(10, 64, 2, 2)
    # tile = 1 ("strided" layout): ray i on rank i % G
    for n in (1000, 7, 0):
        for world in (1, 2, 8):
            ids = [parallel.ray_shard_indices(n, 64, r, world, tile=1) for r in range(world)]
            assert sorted(torch.cat(ids).tolist()) == list(range(n))
            assert all(torch.all(i % world == r) for r, i in enumerate(ids))


def test_object_shards_partition():
    for n in (0, 1, 16, 32, 33):
        for world in (1, 2, 8):
            got = sorted(sum((parallel.object_shard(n, r, world) for r in range(world)), []))
            assert got == list(range(n))
    assert parallel.object_shard(32, 3, 8) == [3, 11, 19, 27]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _full_reference():
    """float64 throughout: the test is about the sharding logic, not fp32 summation order (the pose gradient is ill-conditioned)."""
    obj = oracle.synthetic_object(5, im_sz=IM)
    obj = {k: (v.double() if torch.is_tensor(v) and v.is_floating_point() else v) for k, v in obj.items()}
    sd = {k: v.double() for k, v in oracle.init_codenerf_state(shape_blocks=3, texture_blocks=1, seed=5).items()}
    shp, tex = [t.double() for t in oracle.synthetic_latents(5, 1)]
    jit = torch.rand(IM * IM, S, generator=torch.Generator().manual_seed(5)).double()
    return obj, sd, shp, tex, jit


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.set_num_threads(1)
        obj, sd, shp0, tex0, jit = _full_reference()
        n = IM * IM
        a, b = parallel.ray_shard(n, S, rank, world)
        ids = torch.arange(a, b)
        cam = obj["cam_pose"].clone().requires_grad_()
        shp, tex = shp0.clone().requires_grad_(), tex0.clone().requires_grad_()
        rgb, dep, acc, _ = oracle.render_rays_box(sd, obj["K"], cam, obj["wlh"], obj["roi"], IM, S, shp, tex, jit[ids], ray_ids=ids)
        tgt_all, occ_all = obj["img"].reshape(-1, 3), obj["mask_occ"].reshape(-1, 1)
        # partial loss over the GLOBAL denominator (what parallel.refine_loss_sharded evaluates with the CUDA loss kernel)
        den = occ_all.abs().sum() + 1e-9
        occ = occ_all[a:b]
        part = ((rgb - tgt_all[a:b]) ** 2 * occ.abs()).sum() / den + 0.1 * (torch.exp(-occ * (0.5 - acc.unsqueeze(-1))) * occ.abs()).sum() / den
        part.backward()
        loss = parallel.allreduce_grads([cam, shp, tex], part)
        full_rgb = parallel.gather_rays(rgb.detach(), n, S)
        # interleaved layout: scatter this rank's tiles of a known per-ray value, gather, expect the identity
        mine = parallel.ray_shard_indices(n, S, rank, world)
        marks = torch.stack([mine.double(), mine.double() * 2 + 1], -1)
        back = parallel.gather_rays(marks, n, S, layout="interleaved")
        assert torch.equal(back[:, 0], torch.arange(n).double()) and torch.equal(back[:, 1], torch.arange(n).double() * 2 + 1)
        q.put((rank, float(loss), cam.grad.clone(), shp.grad.clone(), tex.grad.clone(), full_rgb))
    finally:
        dist.destroy_process_group()


def test_ray_sharded_allreduce_matches_single_process():
    world = 2

    def run_world():
        ctx = mp.get_context("spawn")
        q = ctx.Queue()
        port = _free_port()
        procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
        for p in procs:
            p.start()
        try:
            out = sorted([q.get(timeout=600) for _ in range(world)], key=lambda t: t[0])
            for p in procs:
                p.join(timeout=60)
                if p.exitcode != 0:
                    raise RuntimeError("worker exit code %r" % p.exitcode)
            return out
        finally:
            for p in procs:
                if p.is_alive():
                    p.kill()

    res, last = None, None
    for _attempt in range(3):   # the rendezvous port is picked by bind-and-release: retry if another process grabbed it meanwhile
        try:
            res = run_world()
            break
        except Exception as exc:   # noqa: BLE001  (numeric checks below are NOT retried)
            last = exc
    assert res is not None, last
    # single-process truth
    obj, sd, shp0, tex0, jit = _full_reference()
    cam = obj["cam_pose"].clone().requires_grad_()
    shp, tex = shp0.clone().requires_grad_(), tex0.clone().requires_grad_()
    rgb, dep, acc, _ = oracle.render_rays_box(sd, obj["K"], cam, obj["wlh"], obj["roi"], IM, S, shp, tex, jit)
    loss = oracle.refine_losses(rgb, acc, obj["img"].reshape(-1, 3), obj["mask_occ"].reshape(-1, 1))[0]
    loss.backward()

    def rel(x, y):
        return ((x - y).abs().max() / y.abs().max()).item()
    for rank, l, g_cam, g_shp, g_tex, full_rgb in res:
        assert abs(l - loss.item()) <= 1e-6 * abs(loss.item())          # the flat all-reduce buffer is fp32
        assert rel(g_cam, cam.grad) < 1e-5 and rel(g_shp, shp.grad) < 1e-5 and rel(g_tex, tex.grad) < 1e-5
        assert rel(full_rgb, rgb.detach()) < 1e-9
    # every rank ends with identical bits (so identical optimiser steps)
    for t0, t1 in zip(res[0][2:5], res[1][2:5]):
        assert torch.equal(t0, t1)


def _dp_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        model = torch.nn.Sequential(torch.nn.Linear(7, 5), torch.nn.ReLU(), torch.nn.Linear(5, 3))
        x = torch.arange(4 * 7, dtype=torch.float32).reshape(4, 7) / 10.0
        mine = parallel.object_shard(4, rank, world)            # objects r, r+G, ...
        model(x[mine]).pow(2).sum().backward()
        n = parallel.allreduce_weight_grads(model)
        flat_grads = [p.grad.clone() for p in model.parameters()]
        # the same exchange through the bucketed, hook-driven path (config C5's overlap machinery; no side stream on the CPU)
        model.zero_grad(set_to_none=True)
        params = list(model.parameters())
        buckets = parallel.BucketedGradAllReduce([params[2:], params[:2]])   # the last layer's gradients are ready first
        model(x[mine]).pow(2).sum().backward()
        n2 = buckets.finish()
        bucket_grads = [p.grad.clone() for p in model.parameters()]
        buckets.remove()
        q.put((rank, n, flat_grads, n2, bucket_grads))
    finally:
        dist.destroy_process_group()


def test_data_parallel_weight_gradient_allreduce():
    """C5 host logic: objects split over 2 ranks, one flat all-reduce of the weight gradients, averaged."""
    world = 2
    ctx = mp.get_context("spawn")
    res, last = None, None
    for _attempt in range(3):
        q = ctx.Queue()
        port = _free_port()
        procs = [ctx.Process(target=_dp_worker, args=(r, world, port, q)) for r in range(world)]
        for p in procs:
            p.start()
        try:
            res = sorted([q.get(timeout=300) for _ in range(world)], key=lambda t: t[0])
            for p in procs:
                p.join(timeout=60)
            break
        except Exception as exc:   # noqa: BLE001
            last = exc
        finally:
            for p in procs:
                if p.is_alive():
                    p.kill()
    assert res is not None, last
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(7, 5), torch.nn.ReLU(), torch.nn.Linear(5, 3))
    x = torch.arange(4 * 7, dtype=torch.float32).reshape(4, 7) / 10.0
    model(x).pow(2).sum().backward()
    for rank, n, grads, n2, bucket_grads in res:
        assert n == n2 == sum(p.numel() for p in model.parameters())
        for g, gb, p in zip(grads, bucket_grads, model.parameters()):
            assert torch.allclose(g, p.grad / world, rtol=1e-5, atol=1e-7)
            assert torch.equal(gb, g)      # bucketed exchange == flat exchange, bit for bit (same sums, same scale)

