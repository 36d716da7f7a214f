#!/usr/bin/env python
"""Times the decoder C-ABI calls alone (CUDA events, current stream): fwd and bwd at 1 object x 16384 rays x 64 samples."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import supnerf_b200 as snb
from oracle import oracle
dev = "cuda"
sd = oracle.init_codenerf_state(shape_blocks=3, texture_blocks=1, seed=0)
m = snb.AutoRFMix(3, 1, 256); m.load_state_dict(sd); m = m.to(dev); m.precision = sys.argv[1] if len(sys.argv) > 1 else "bf16"; m.requires_grad_(False)
N, S = 16384, 64
g = torch.Generator().manual_seed(0)
xyz = ((torch.rand(N, S, 3, generator=g) - 0.5) * 1.6).to(dev).requires_grad_()
vd = torch.nn.functional.normalize(torch.randn(N, 1, 3, generator=g), dim=-1).repeat(1, S, 1).to(dev).requires_grad_()
shp, tex = [t.to(dev).requires_grad_() for t in oracle.synthetic_latents(0, 1)]
def run(n):
    tf, tb = [], []
    for _ in range(n):
        e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        e[0].record(); sig, rgb = m(xyz, vd, shp, tex); e[1].record()
        gs, gr = torch.ones_like(sig), torch.ones_like(rgb)
        e[2].record(); torch.autograd.grad([sig, rgb], [xyz, vd, shp, tex], [gs, gr]); e[3].record()
        torch.cuda.synchronize()
        tf.append(e[0].elapsed_time(e[1])); tb.append(e[2].elapsed_time(e[3]))
    return min(tf), min(tb)
run(3)
import ctypes
lib = snb._lib.load()
hnd = m._handle(xyz.device).h
lib.snb_kernel_timing_enable(hnd, 1)
f, b = run(10)
buf = (ctypes.c_float * 64)()
kf = [buf[i] for i in range(lib.snb_kernel_timing_read(hnd, 0, buf, 64))]
kb = [buf[i] for i in range(lib.snb_kernel_timing_read(hnd, 1, buf, 64))]
lib.snb_kernel_timing_enable(hnd, 0)
fl = 2 * 449664 * N * S / 1e12
if kf: print(f"   kernel only: fwd {min(kf):.3f} ms ({2*449664*N*S/1e12/min(kf)*1e3:.0f} TF/s)  bwd {min(kb):.3f} ms ({2*449664*N*S/1e12/min(kb)*1e3:.0f} TF/s)")
print(f"SNB_TC_EXP={os.environ.get('SNB_TC_EXP','0')} fwd {f:.3f} ms ({fl/f*1e3:.0f} TF/s)  bwd {b:.3f} ms ({fl/b*1e3:.0f} TF/s)")
