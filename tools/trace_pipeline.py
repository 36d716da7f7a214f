#!/usr/bin/env python
"""Per-step pipeline timeline of CTA 0 of the two-tile tcgen05 decoder kernels (snb_tc_set_trace): for one steady-state
tile pair prints, per step and slot, when the MMA warp saw the operand ready, finished issuing, when the epilogue saw the
accumulator and when it published (cycles relative to the pair's first stamp)."""
import ctypes, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import supnerf_b200 as snb
from oracle import oracle
dev = "cuda"
which = sys.argv[1] if len(sys.argv) > 1 else "fwd"
sd = oracle.init_codenerf_state(shape_blocks=3, texture_blocks=1, seed=0)
m = snb.AutoRFMix(3, 1, 256); m.load_state_dict(sd); m = m.to(dev); m.precision = "bf16"; m.requires_grad_(False)
N, S = 16384, 64
g = torch.Generator().manual_seed(0)
xyz = ((torch.rand(N, S, 3, generator=g) - 0.5) * 1.6).to(dev).requires_grad_()
vd = torch.nn.functional.normalize(torch.randn(N, 1, 3, generator=g), dim=-1).repeat(1, S, 1).to(dev).requires_grad_()
shp, tex = [t.to(dev).requires_grad_() for t in oracle.synthetic_latents(0, 1)]
lib = snb._lib.load()
# forward under the (default) cta_group::2 kernels: encoding_viewdir is one step -> 8 steps; otherwise 9.  Override: argv[2].
n_pairs = 28
n_steps = int(sys.argv[2]) if len(sys.argv) > 2 else (8 if which == "fwd" and os.environ.get("SNB_TC_CG2", "1") != "0" else 9)
buf = torch.zeros(n_pairs * n_steps * 2 * 4 + 64, dtype=torch.int64, device=dev)
sig, rgb = m(xyz, vd, shp, tex)   # warm-up
torch.cuda.synchronize()
if which == "fwd":
    lib.snb_tc_set_trace(m._handle(torch.device(dev)).h, ctypes.c_void_p(buf.data_ptr()))
    sig, rgb = m(xyz, vd, shp, tex)
    torch.cuda.synchronize()
    lib.snb_tc_set_trace(m._handle(torch.device(dev)).h, None)
else:
    gs, gr = torch.ones_like(sig), torch.ones_like(rgb)
    lib.snb_tc_set_trace(m._handle(torch.device(dev)).h, ctypes.c_void_p(buf.data_ptr()))
    torch.autograd.grad([sig, rgb], [xyz, vd, shp, tex], [gs, gr])
    torch.cuda.synchronize()
    lib.snb_tc_set_trace(m._handle(torch.device(dev)).h, None)
t = buf[: n_pairs * n_steps * 8].cpu().reshape(n_pairs, n_steps, 2, 4)
pair = 10
t0 = int(t[pair][t[pair] > 0].min())
print(f"{which}: pair {pair} of CTA 0, cycles since the pair's first stamp (0 = not stamped)")
print("step slot | ready_seen  issued | acc_seen published | mma+commit(acc_seen-ready_seen)  epilogue(published-acc_seen)")
for si in range(n_steps):
    for sl in range(2):
        a = [int(v) - t0 if int(v) > 0 else 0 for v in t[pair, si, sl]]
        print(f"{si:4d} {sl:4d} | {a[0]:10d} {a[1]:7d} | {a[2]:8d} {a[3]:9d} | {a[2]-a[0]:8d} {a[3]-a[2]:8d}")
tot = int(t[pair + 1][t[pair + 1] > 0].min()) - t0
print("pair period (cycles):", tot)
