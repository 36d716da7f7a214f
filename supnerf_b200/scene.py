"""Multi-object scene compositor: the per-ray merge of several objects' samples by depth that
``scripts/demo.py`` (``vis_scene``, lines 560-569) writes inline with torch.sort / searchsorted / scatter_ on the CPU,
as one CUDA kernel plus the package's compositing kernel.  Forward only, like the reference (it runs under no_grad)."""
import torch

from . import _lib, ops
from ._lib import check, f32c, on_device, ptr, require_cuda, stream_ptr


def merge_objects(z_vals, sigmas, rgbs, return_args=False):
    """z_vals, sigmas (R, K), rgbs (R, K, 3) with K = n_objects * n_samples (<= 1024), every object's samples along each
    ray -> (z_sort, sigmas_sort, rgbs_sort[, z_args]) exactly as demo.py:560-567 builds them: depths sorted ascending,
    ``z_args = searchsorted(z_sort, z_vals)``, values scattered to ``z_args`` (ties: the last sample in index order wins,
    the remaining slots of the tie group are zero)."""
    lib = _lib.load()
    require_cuda(z_vals, sigmas, rgbs)
    z, s, c = f32c(z_vals), f32c(sigmas), f32c(rgbs)
    r, k = z.shape
    if tuple(s.shape) != (r, k) or tuple(c.shape) != (r, k, 3):
        raise ValueError("merge_objects: expected z_vals (R,K), sigmas (R,K), rgbs (R,K,3)")
    z_sort, s_sort, c_sort = torch.empty_like(z), torch.empty_like(s), torch.empty_like(c)
    args = torch.empty(r, k, dtype=torch.int64, device=z.device) if return_args else None
    with on_device(z.device):
        check(lib.snb_merge_sort_samples(ptr(z), ptr(s), ptr(c), r, k, ptr(z_sort), ptr(s_sort), ptr(c_sort), ptr(args), stream_ptr()),
              "snb_merge_sort_samples")
    return (z_sort, s_sort, c_sort, args) if return_args else (z_sort, s_sort, c_sort)


def render_merged(z_vals, sigmas, rgbs, white_bkgd=True):
    """demo.py:560-569: merge the objects' samples by depth, then ``volume_rendering3(..., white_bkgd=True)``.
    -> rgb (R,3), depth (R,), acc_trans (R,)."""
    z_sort, s_sort, c_sort = merge_objects(z_vals, sigmas, rgbs)
    return ops.composite(s_sort, c_sort, z_sort, white_bkgd, True)
