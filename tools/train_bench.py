#!/usr/bin/env python
"""Config C5 decoder half: one training step's render per GPU -- B = 8 objects x 1024 rays x 64 samples through
the decoder (trainer_nerf_nuscenes.py:40-60 ParallelModel.forward: model(xyz, viewdir, codes) -> volume_rendering_batch ->
the two losses) forward + backward INCLUDING every weight gradient; fp32 back end vs bf16 training mode.
  python tools/train_bench.py [--objects 8] [--rays 1024] [--steps 10]"""
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
a = ap.parse_args()
dev = torch.device("cuda", 0)
B, n, S = a.objects, a.rays, a.samples
g = torch.Generator().manual_seed(5)
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

    def step():
        m.zero_grad(set_to_none=True)
        shp.grad = tex.grad = None
        sig, rgbs = m(xyz.reshape(-1, S, 3), vd.reshape(-1, S, 3), shp, tex)
        rgb, dep, acc = snb.utils.volume_rendering_batch(sig.view(B, n, S, 1), rgbs.view(B, n, S, 3), z)
        den = occ.abs().sum() + 1e-9
        loss = ((rgb - tgt) ** 2 * occ.abs()).sum() / den + 0.1 * (torch.exp(-occ * (0.5 - acc.unsqueeze(-1))) * occ.abs()).sum() / den
        loss.backward()
        return loss

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        loss = step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    out[prec] = {"ms_per_step": round(ms, 3), "rays_per_s": round(B * n / (ms / 1e3), 1), "loss": float(loss.detach()),
                 "tflops_fwd_dgrad_wgrad": round(3 * 2 * 449664 * B * n * S / (ms / 1e3) / 1e12, 1)}
print(json.dumps({"config": "C5 decoder step: %d objects x %d rays x %d samples, fwd + bwd with ALL weight gradients, 1 GPU" % (B, n, S), **out}))
