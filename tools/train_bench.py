#!/usr/bin/env python
"""Config C5 decoder half: one training step's render per GPU -- B = 8 objects x 1024 rays x 64 samples through
the decoder (trainer_nerf_nuscenes.py:40-60 ParallelModel.forward: model(xyz, viewdir, codes) -> volume_rendering_batch ->
the two losses) forward + backward INCLUDING every weight gradient; fp32 back end vs bf16 training mode.
  python tools/train_bench.py [--objects 8] [--rays 1024] [--steps 10]
Data-parallel (config C5's "data-parallel weight allreduce on 8 x B200"; weak scaling, every rank its own 8 objects):
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/train_bench.py --prec bf16 --optimizer
after the backward ONE all_reduce over the flat buffer of all weight gradients (parallel.allreduce_weight_grads), then (with
--optimizer) a fused torch AdamW step on the weights and the codes (so the next forward re-packs the bf16 weight images)."""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import supnerf_b200 as snb  # noqa: E402
from supnerf_b200 import synthetic  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--objects", type=int, default=8)
ap.add_argument("--rays", type=int, default=1024)
ap.add_argument("--samples", type=int, default=64)
ap.add_argument("--steps", type=int, default=10)
ap.add_argument("--prec", default="fp32,bf16")
ap.add_argument("--optimizer", action="store_true", help="also take an AdamW step on weights + codes every step")
a = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=dev)
B, n, S = a.objects, a.rays, a.samples
g = torch.Generator().manual_seed(5 + rank)
xyz = ((torch.rand(B, n, S, 3, generator=g) - 0.5) * 1.2).to(dev)
vd = torch.nn.functional.normalize(torch.randn(B, n, 1, 3, generator=g), dim=-1).repeat(1, 1, S, 1).to(dev)
z = (torch.rand(B, S, generator=g).sort(-1).values * 4 + 8).to(dev)
tgt, occ = torch.rand(B, n, 3, generator=g).to(dev), torch.randint(-1, 2, (B, n, 1), generator=g).float().to(dev)
shp0, tex0 = synthetic.synthetic_latents(5, B)
sd = synthetic.init_codenerf_state(shape_blocks=3, texture_blocks=1, seed=5)
out = {}
for prec in a.prec.split(","):
    m = snb.SUPNeRF(3, 1, 3, 3, 256)
    m.load_state_dict(sd)
    m = m.to(dev)
    m.precision = prec
    shp, tex = shp0.to(dev).requires_grad_(), tex0.to(dev).requires_grad_()
    opt = torch.optim.AdamW([{"params": list(m.parameters()), "lr": 1e-4}, {"params": [shp, tex], "lr": 1e-3}], fused=True) if a.optimizer else None

    def step():
        m.zero_grad(set_to_none=True)
        shp.grad = tex.grad = None
        sig, rgbs = m(xyz.reshape(-1, S, 3), vd.reshape(-1, S, 3), shp, tex)
        rgb, dep, acc = snb.utils.volume_rendering_batch(sig.view(B, n, S, 1), rgbs.view(B, n, S, 3), z)
        den = occ.abs().sum() + 1e-9
        loss = ((rgb - tgt) ** 2 * occ.abs()).sum() / den + 0.1 * (torch.exp(-occ * (0.5 - acc.unsqueeze(-1))) * occ.abs()).sum() / den
        loss.backward()
        if world > 1:
            snb.parallel.allreduce_weight_grads(m)   # one flat all_reduce(sum) + 1/G; the codes are rank-local
        if opt is not None:
            opt.step()
        return loss

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        loss = step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    out[prec] = {"ms_per_step": round(ms, 3), "rays_per_s": round(world * B * n / (ms / 1e3), 1), "loss": float(loss.detach()),
                 "tflops_fwd_dgrad_wgrad_per_gpu": round(3 * 2 * 449664 * B * n * S / (ms / 1e3) / 1e12, 1)}
if rank == 0:
    print(json.dumps({"config": "C5 decoder step: %d objects x %d rays x %d samples per GPU, fwd + bwd with ALL weight gradients%s%s, %d GPU(s)"
                                % (B, n, S, ", one all_reduce of the flat weight-gradient buffer" if world > 1 else "",
                                   ", fused AdamW step on weights + codes" if a.optimizer else "", world), "n_gpus": world, **out}))
if world > 1:
    dist.destroy_process_group()
