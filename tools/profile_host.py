#!/usr/bin/env python
"""Host-side cost of one C2 step (16 objects through NeRFRenderer.render_rays + refine_loss + backward): wall time to ISSUE the
step (no synchronisation) vs GPU time, and a cProfile of the issuing code."""
import cProfile
import os
import pstats
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import supnerf_b200 as snb  # noqa: E402
from supnerf_b200 import synthetic  # noqa: E402

dev = torch.device("cuda", 0)
sd = synthetic.init_codenerf_state(shape_blocks=3, texture_blocks=1, seed=0)
model = snb.AutoRFMix(3, 1, 256)
model.load_state_dict(sd)
model = model.to(dev)
model.precision = "bf16"
model.requires_grad_(False)
R = snb.renderer.NeRFRenderer(n_samples=64)
objs = []
for i in range(16):
    o = synthetic.synthetic_object(100 + i, im_sz=128)
    s, t = synthetic.synthetic_latents(100 + i, 1)
    objs.append(dict(K=o["K"].to(dev), cam=o["cam_pose"].to(dev).requires_grad_(), wlh=o["wlh"], roi=o["roi"], img=o["img"].to(dev),
                     mask=o["mask_occ"].to(dev), shp=s.to(dev).requires_grad_(), tex=t.to(dev).requires_grad_()))


def step():
    for d in objs:
        d["cam"].grad = d["shp"].grad = d["tex"].grad = None
        rgb, dep, acc, tgt, occ = R.render_rays(model, dev, d["img"], d["mask"], d["cam"], d["wlh"], d["K"], d["roi"], d["shp"], d["tex"], im_sz=128)
        loss = snb.losses.refine_loss(rgb, acc, tgt, occ, 0.1)[0]
        loss.backward()


for _ in range(3):
    step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5):
    step()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print("issue %.2f ms/step, issue+drain %.2f ms/step" % ((t1 - t0) / 5 * 1e3, (t2 - t0) / 5 * 1e3))
pr = cProfile.Profile()
pr.enable()
for _ in range(5):
    step()
pr.disable()
torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
