import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    return {k: z[k] for k in z.files}


def T(a, dtype=None, device=None):
    t = torch.from_numpy(np.asarray(a))
    if dtype is not None:
        t = t.to(dtype)
    if device is not None:
        t = t.to(device)
    return t


def rel_err(a, b):
    """per-tensor max|a-b| / max|b|  (the metric SURVEY §8(d) fixes for every float comparison)."""
    a = torch.as_tensor(a).detach().double().cpu()
    b = torch.as_tensor(b).detach().double().cpu()
    den = b.abs().max().item()
    return (a - b).abs().max().item() / (den if den > 0 else 1.0)


def close_vs_truth(got, ref32, truth64, tol=1e-5, slack=4.0):
    """Parity criterion for ill-conditioned fp32 reductions (pose / bias gradients: sums with heavy cancellation,
    where the fp32 reference itself is only accurate to ~1e-4 of its own value).  `truth64` is the fp64 oracle;
    `ref32` the fp32 reference (golden or fp32 oracle).  Pass iff the kernel is within `tol` of the truth, or no
    further from it than `slack` x the fp32 reference's own distance.  Returns (ok, err_got, err_ref)."""
    e_got = rel_err(got, truth64)
    e_ref = rel_err(ref32, truth64)
    return e_got <= max(tol, slack * e_ref), e_got, e_ref
