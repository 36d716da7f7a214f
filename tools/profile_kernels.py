#!/usr/bin/env python
"""Short run of the three hot kernels for `ncu --set full`: the tcgen05 decoder forward / backward on `--rays` rays x 64
samples (AutoRF-mix 3/1/256, one object) and the compositing forward / backward on `--crays` rays.  One warm-up, one
measured pass each; prints CUDA-event times."""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import supnerf_b200 as snb  # noqa: E402
from supnerf_b200 import ops  # noqa: E402
from oracle import oracle  # noqa: E402  (synthetic weights / latents only)

ap = argparse.ArgumentParser()
ap.add_argument("--rays", type=int, default=16384)
ap.add_argument("--crays", type=int, default=1 << 20)
ap.add_argument("--samples", type=int, default=64)
a = ap.parse_args()
dev = "cuda"
S = a.samples
sd = oracle.init_codenerf_state(shape_blocks=3, texture_blocks=1, seed=0)
m = snb.AutoRFMix(3, 1, 256)
m.load_state_dict(sd)
m = m.to(dev)
m.precision = "bf16"
m.requires_grad_(False)
g = torch.Generator().manual_seed(0)
xyz = ((torch.rand(a.rays, S, 3, generator=g) - 0.5) * 1.6).to(dev).requires_grad_()
vd = torch.nn.functional.normalize(torch.randn(a.rays, 1, 3, generator=g), dim=-1).repeat(1, S, 1).to(dev).requires_grad_()
shp, tex = [t.to(dev).requires_grad_() for t in oracle.synthetic_latents(0, 1)]


def ev():
    return torch.cuda.Event(enable_timing=True)


def decoder_pass():
    e = [ev() for _ in range(4)]
    e[0].record()
    sig, rgb = m(xyz, vd, shp, tex)
    e[1].record()
    gs, gr = torch.ones_like(sig), torch.ones_like(rgb)
    e[2].record()
    torch.autograd.grad([sig, rgb], [xyz, vd, shp, tex], [gs, gr])
    e[3].record()
    torch.cuda.synchronize()
    return e[0].elapsed_time(e[1]), e[2].elapsed_time(e[3])


n = a.crays
sigma = (torch.rand(n, S, generator=g) * 4 - 1).to(dev).requires_grad_()
rgbs = torch.rand(n, S, 3, generator=g).to(dev).requires_grad_()
z = (torch.rand(n, S, generator=g).sort(-1).values + 0.5).to(dev).requires_grad_()
go = (torch.rand(n, 3, device=dev), torch.rand(n, device=dev), torch.rand(n, device=dev))


def composite_pass():
    e = [ev() for _ in range(4)]
    e[0].record()
    out = ops.composite(sigma, rgbs, z, True)
    e[1].record()
    e[2].record()
    torch.autograd.grad(out, [sigma, rgbs, z], go)
    e[3].record()
    torch.cuda.synchronize()
    return e[0].elapsed_time(e[1]), e[2].elapsed_time(e[3])


decoder_pass()
for _ in range(3):
    f, b = decoder_pass()
    print(f"  pass: fwd {f:.3f} ms  bwd {b:.3f} ms")
fl = 2 * 449664 * a.rays * S / 1e12
print(f"decoder rays={a.rays} S={S}: fwd {f:.3f} ms ({fl / f * 1e3:.0f} TF/s)  bwd {b:.3f} ms ({fl / b * 1e3:.0f} TF/s)  [C-ABI call incl. latent kernels]")
composite_pass()
best = [1e9, 1e9]
for _ in range(5):
    f, b = composite_pass()
    best = [min(best[0], f), min(best[1], b)]
bf, bb = n * (20 * S + 20), n * (40 * S + 20)   # fwd reads 20S writes 20; bwd re-reads 20S + 20, writes 20S
print(f"composite rays={n} S={S}: fwd {best[0]:.3f} ms ({bf / best[0] / 1e6:.0f} GB/s)  bwd {best[1]:.3f} ms ({bb / best[1] / 1e6:.0f} GB/s)  "
      f"fwd+bwd {(bf + bb) / (best[0] + best[1]) / 1e6:.0f} GB/s  [python autograd call; kernel + output allocation]")
