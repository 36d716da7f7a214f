#!/usr/bin/env python
"""Short, self-contained run of the bench workload for ncu: 1 warm-up + 1 measured step of `--objects` objects
(128x128 rays x 64 samples each, AutoRF-mix 3/1/256, fwd+bwd to pose + latents).  Prints the step time (CUDA events)."""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import supnerf_b200 as snb  # noqa: E402
from oracle import oracle  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--objects", type=int, default=2)
ap.add_argument("--precision", default="bf16")
ap.add_argument("--im-sz", type=int, default=128)
a = ap.parse_args()
dev = torch.device("cuda", 0)
sd = oracle.init_codenerf_state(shape_blocks=3, texture_blocks=1, seed=0)
model = snb.AutoRFMix(3, 1, 256)
model.load_state_dict(sd)
model = model.to(dev)
model.precision = a.precision
model.requires_grad_(False)
R = snb.renderer.NeRFRenderer(n_samples=64)
objs = bench.make_objects(100, a.objects, a.im_sz)
d = [dict(K=o["K"].to(dev), cam=o["cam_pose"].to(dev).requires_grad_(), wlh=o["wlh"], roi=o["roi"], tgt=o["img"].reshape(-1, 3).to(dev),
          occ=o["mask_occ"].reshape(-1, 1).to(dev), shp=o["shapecode"].to(dev).requires_grad_(), tex=o["texturecode"].to(dev).requires_grad_())
     for o in objs]


def step():
    for x in d:
        x["cam"].grad = x["shp"].grad = x["tex"].grad = None
        ro, vd = snb.utils.get_rays(x["K"], x["cam"], x["roi"], uv_steps=[a.im_sz, a.im_sz])
        xyz, vdr, zv, _ = R.prepare_sampled_rays(ro, vd, x["wlh"])
        sig, rgbs = model(xyz, vdr, x["shp"], x["tex"])
        rgb, dep, acc = R.volume_render(sig.squeeze(-1), rgbs, zv)
        bench.refine_loss(rgb, acc, x["tgt"], x["occ"]).backward()


step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
step()
e1.record()
torch.cuda.synchronize()
print("step_ms", e0.elapsed_time(e1), "objects", a.objects, "rays", a.objects * a.im_sz ** 2)
