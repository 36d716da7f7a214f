#!/usr/bin/env python
"""Per-row error profile of the split-precision decoder's input gradients against the fp64 oracle: separates arithmetic error (every
row off by ~1e-6 of its own size) from ReLU-gate flips (a unit whose pre-activation is within rounding of zero opens in one
implementation and not in the other: single rows off by percents, all others exact).  GPU only."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import supnerf_b200 as S  # noqa: E402
from oracle import oracle  # noqa: E402


def main():
    torch.manual_seed(0)
    sd = oracle.init_codenerf_state(shape_blocks=3, texture_blocks=1, seed=23)
    n = 128 * 64
    xyz = torch.rand(n, 1, 3) - 0.5
    vd = torch.nn.functional.normalize(torch.randn(n, 1, 3), dim=-1)
    shp, tex = oracle.synthetic_latents(5, 1)
    up_s, up_c = torch.randn(n, 1, 1), torch.randn(n, 1, 3)
    out = {}
    for dt in (torch.float32, torch.float64):
        sdd = {k: v.to(dt) for k, v in sd.items()}
        ins = [t.detach().clone().to(dt).requires_grad_() for t in (xyz, vd, shp, tex)]
        sig, rgb = oracle.codenerf_decoder(sdd, *ins)
        ((sig * up_s.to(dt)).sum() + (rgb * up_c.to(dt)).sum()).backward()
        out[dt] = ins[0].grad.reshape(n, 3).double()
    m = S.CodeNeRF(shape_blocks=3, texture_blocks=1)
    m.load_state_dict(sd)
    m = m.to("cuda")
    m.precision = "fp32"
    m.requires_grad_(False)
    for tc in (True, False):
        S.ops.FP32_TENSOR_CORES = tc
        gin = [t.detach().clone().cuda().requires_grad_() for t in (xyz, vd, shp, tex)]
        sig, rgb = m(*gin)
        ((sig * up_s.cuda()).sum() + (rgb * up_c.cuda()).sum()).backward()
        g = gin[0].grad.reshape(n, 3).double().cpu()
        truth = out[torch.float64]
        row_err = (g - truth).abs().amax(1) / truth.abs().amax(1).clamp_min(1e-30)
        srt = torch.sort(row_err, descending=True)[0]
        print("tensor cores" if tc else "FFMA        ", "rows %d  median %.2e  p99 %.2e  worst five %s   (fp32 oracle: median %.2e worst %.2e)" % (
            n, row_err.median(), srt[n // 100], ["%.1e" % v for v in srt[:5]],
            ((out[torch.float32] - truth).abs().amax(1) / truth.abs().amax(1).clamp_min(1e-30)).median(),
            ((out[torch.float32] - truth).abs().amax(1) / truth.abs().amax(1).clamp_min(1e-30)).max()))


if __name__ == "__main__":
    main()
