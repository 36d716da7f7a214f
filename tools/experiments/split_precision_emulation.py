#!/usr/bin/env python
"""CPU emulation of split-precision tensor-core arithmetic for the decoder's fp32 (1e-5 parity) mode: how far from the fp32 oracle /
fp64 truth does the CodeNeRF decoder land when every GEMM operand is split into 16-bit parts and only the leading cross products are
issued?  Products of 16-bit parts are exact in fp32; accumulation is emulated in fp32 (torch fp32 matmul of the up-cast parts).
    fp16 x 2 parts, 3 MMAs (hh, hl, lh)        <- candidate: a third of the bf16 rate
    bf16 x 3 parts, 6 MMAs (i + j <= 2)        <- a sixth
    bf16 x 2 parts, 3 MMAs
Forward outputs and input / latent gradients, error metric = tests/conftest.py:rel_err.
    python tools/experiments/split_precision_emulation.py [--rz [--corr-first]]
--rz: the accumulator is truncated (round toward zero) after every K = 16 instruction, as the tensor core does; --corr-first: the
kernels' product order (all correction products of a layer before its leading products).  Gradients in --rz mode are dominated by
single ReLU gates that open in one arithmetic and not in the other (see fp32_tc_row_errors.py for the per-row picture)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle  # noqa: E402


RZ_ACCUMULATE = "--rz" in sys.argv
CORR_FIRST = "--corr-first" in sys.argv


def rz32(x64):
    """float64 -> the fp32 value next toward zero (as float64)."""
    f = x64.float()
    over = f.double().abs() > x64.abs()
    stepped = torch.nextafter(f, torch.zeros_like(f))
    return torch.where(over, stepped, f).double()


def split(x, dtype, parts):
    out, r = [], x.clone()
    for _ in range(parts):
        p = r.to(dtype).to(x.dtype)
        out.append(p)
        r = r - p
    return out


def make_mm(mode):
    if mode in ("fp32", "fp64"):
        return lambda a, b: a @ b
    dtype, parts, keep = {"fp16x2": (torch.float16, 2, 1), "bf16x3": (torch.bfloat16, 3, 2), "bf16x2": (torch.bfloat16, 2, 1)}[mode]

    def mm(a, b):
        # power-of-two scaling keeps the low parts out of the fp16 subnormal range (exact, undone after the sum)
        sa = 2.0 ** torch.floor(torch.log2(a.abs().amax(dim=1, keepdim=True).clamp_min(1e-30))) if dtype == torch.float16 else 1.0   # per row
        sb = 2.0 ** torch.floor(torch.log2(b.abs().max().clamp_min(1e-30))) if dtype == torch.float16 else 1.0
        pa, pb = split(a / sa, dtype, parts), split(b / sb, dtype, parts)
        if RZ_ACCUMULATE:
            # the tensor core's accumulator: every K = 16 instruction adds its (exact) block sum and TRUNCATES to fp32 (round toward zero)
            acc = torch.zeros(a.shape[0], b.shape[1], dtype=torch.float64)
            terms = ((1, 0), (0, 0), (0, 1)) if parts == 2 else [(i, j) for i in range(parts) for j in range(parts) if i + j <= keep]
            if CORR_FIRST:   # every K block's correction products first (the accumulator is still ~2^-11 of its final size), then the leading ones
                order = [(k0, t) for t in terms if t != (0, 0) for k0 in range(0, a.shape[1], 16)] + [(k0, (0, 0)) for k0 in range(0, a.shape[1], 16)]
            else:            # K block by K block
                order = [(k0, t) for k0 in range(0, a.shape[1], 16) for t in terms]
            for k0, (i, j) in order:
                acc = rz32(acc + pa[i][:, k0:k0 + 16].double() @ pb[j][k0:k0 + 16].double())
            return acc.float() * (sa * sb)
        acc = torch.zeros(a.shape[0], b.shape[1], dtype=torch.float32)
        for i in reversed(range(parts)):           # small terms first
            for j in reversed(range(parts)):
                if i + j <= keep:
                    acc = acc + pa[i] @ pb[j]
        return acc * (sa * sb)
    return mm


class Lin(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, mm):
        ctx.save_for_backward(x, w)
        ctx.mm = mm
        return mm(x, w.t().contiguous())

    @staticmethod
    def backward(ctx, g):
        x, w = ctx.saved_tensors
        return ctx.mm(g, w), None, None


def decoder(sd, xyz, vd, sl, tl, mode):
    mm = make_mm(mode)
    dt = torch.float64 if mode == "fp64" else torch.float32
    c = lambda t: t.to(dt)   # noqa: E731
    lin = lambda name, x: Lin.apply(x, c(sd[name + ".weight"]), mm) + c(sd[name + ".bias"])   # noqa: E731
    exact = lambda name, x: torch.nn.functional.linear(x, c(sd[name + ".weight"]), c(sd[name + ".bias"]))   # noqa: E731  (SIMT parts)
    bs, bt = oracle.decoder_blocks(sd)
    x, v = oracle.positional_encoding(c(xyz), 10), oracle.positional_encoding(c(vd), 4)
    y = torch.relu(lin("encoding_xyz.0", x))
    for j in range(1, bs + 1):
        z = torch.relu(exact(f"shape_latent_layer_{j}.0", c(sl)))
        y = torch.relu(lin(f"shape_layer_{j}.0", y + z))
    y = lin("encoding_shape", y)
    sig = torch.nn.functional.softplus(exact("sigma.0", y))
    y = torch.relu(lin("encoding_viewdir.0", torch.cat([y, v], -1)))
    for j in range(1, bt + 1):
        z = torch.relu(exact(f"texture_latent_layer_{j}.0", c(tl)))
        y = torch.relu(lin(f"texture_layer_{j}.0", y + z))
    h = torch.relu(lin("rgb.0", y))
    return sig, exact("rgb.2", h)


def rel(a, b):
    return ((a.double() - b.double()).abs().max() / b.double().abs().max()).item()


def run(n=4096, modes=("fp32", "fp16x2", "bf16x3", "bf16x2"), rz=False, corr_first=False, seed=0):
    """-> {mode: {tensor: (error vs the fp32 run, error vs the fp64 run)}} for the shipped 3 / 1 / 256 decoder on n random samples."""
    global RZ_ACCUMULATE, CORR_FIRST
    RZ_ACCUMULATE, CORR_FIRST = rz, corr_first
    torch.manual_seed(seed)
    sd = oracle.init_codenerf_state(shape_blocks=3, texture_blocks=1, seed=7)
    xyz0 = (torch.rand(n, 3) - 0.5)
    vd0 = torch.nn.functional.normalize(torch.randn(n, 3), dim=-1)
    sl0, tl0 = oracle.synthetic_latents(5, 1)
    gs, gr = torch.randn(n, 1), torch.randn(n, 3)
    res = {}
    for mode in ("fp64", "fp32") + tuple(m for m in modes if m != "fp32"):
        xyz, vd, sl, tl = [t.clone().requires_grad_() for t in (xyz0, vd0, sl0, tl0)]
        sig, rgb = decoder(sd, xyz, vd, sl.expand(n, -1), tl.expand(n, -1), mode)
        ((sig.double() * gs.double()).sum() + (rgb.double() * gr.double()).sum()).backward()
        res[mode] = dict(sigma=sig.detach(), rgb=rgb.detach(), g_xyz=xyz.grad, g_vd=vd.grad, g_shape=sl.grad, g_tex=tl.grad)
    return {mode: {k: (rel(res[mode][k], res["fp32"][k]), rel(res[mode][k], res["fp64"][k])) for k in res[mode]} for mode in modes}


def main():
    out = run(rz="--rz" in sys.argv, corr_first="--corr-first" in sys.argv)
    print("%-8s %-8s %12s %12s" % ("mode", "tensor", "vs fp32", "vs fp64"))
    for mode, d in out.items():
        for k, (e32, e64) in d.items():
            print("%-8s %-8s %12.3g %12.3g" % (mode, k, e32, e64))


if __name__ == "__main__":
    main()
