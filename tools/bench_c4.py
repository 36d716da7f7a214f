#!/usr/bin/env python
"""Config C4: ONE high-resolution object, 512x512 rays x 128 samples, ray-sharded over the launched ranks with
one NCCL all-reduce of the flat pose/latent gradient buffer per step.  Strong scaling: total work fixed.
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/bench_c4.py [--steps K]"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import supnerf_b200 as snb  # noqa: E402
from supnerf_b200 import synthetic  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--im", type=int, default=512)
ap.add_argument("--samples", type=int, default=128)
ap.add_argument("--layout", default="interleaved", choices=["contiguous", "interleaved"],
                help="ray shards: one contiguous tile per rank, or every G-th 128-ray tile (balanced under miss-ray compaction)")
a = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
obj = synthetic.synthetic_object(4, im_sz=a.im)
sd = synthetic.init_codenerf_state(shape_blocks=3, texture_blocks=1, seed=4)
m = snb.SUPNeRF(3, 1, 3, 3, 256)
m.load_state_dict(sd)
m = m.to(dev)
m.precision = "bf16"
m.requires_grad_(False)
R = snb.renderer.NeRFRenderer(n_samples=a.samples)
shp0, tex0 = synthetic.synthetic_latents(4, 1)
cam = obj["cam_pose"].to(dev).requires_grad_()
shp, tex = shp0.to(dev).requires_grad_(), tex0.to(dev).requires_grad_()
img, mask, K = obj["img"].to(dev), obj["mask_occ"].to(dev), obj["K"].to(dev)


def step():
    cam.grad = shp.grad = tex.grad = None
    rgb, dep, acc, tgt, occ, occ_all = snb.parallel.render_rays_sharded(R, m, dev, img, mask, cam, obj["wlh"], K, obj["roi"], shp, tex,
                                                                       im_sz=a.im, rank=rank, world=world, layout=a.layout)
    part = snb.parallel.refine_loss_sharded(rgb, acc, tgt, occ, occ_all)
    part.backward()
    return snb.parallel.allreduce_grads([cam, shp, tex], part)


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
ms = e0.elapsed_time(e1)
if world > 1:
    t = torch.tensor([ms], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
if rank == 0:
    n = a.im * a.im
    print(json.dumps({"config": "C4: %dx%d rays x %d samples, ray-sharded x%d (%s tiles), one all-reduce of 525 floats per step" % (a.im, a.im, a.samples, world, a.layout),
                      "n_gpus": world, "ms_per_step": round(ms / a.steps, 3), "rays_per_s": round(n * a.steps / (ms / 1e3), 1),
                      "scaling": "strong", "loss": float(loss)}))
if world > 1:
    dist.destroy_process_group()
