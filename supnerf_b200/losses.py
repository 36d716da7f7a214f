"""The refine-iteration losses, one CUDA launch per direction.

The reference has no function for them: the same five lines are written inline in every optimiser / trainer
(optimizer_nuscenes.py:729-736, optimizer_kitti.py, optimizer_waymo.py, trainer_unified_nuscenes.py:316-332):

    loss_rgb = torch.sum((rgb_rays - rgb_tgt) ** 2 * torch.abs(occ_pixels)) / (torch.sum(torch.abs(occ_pixels)) + 1e-9)
    loss_occ = torch.sum(torch.exp(-occ_pixels * (0.5 - acc_trans_rays.unsqueeze(-1))) * torch.abs(occ_pixels)) / (torch.sum(torch.abs(occ_pixels)) + 1e-9)
    loss = loss_rgb + self.hpams['loss_occ_coef'] * loss_occ

``refine_loss`` computes exactly that with the kernels of csrc/loss.cu; no CPU fallback."""
import torch

from . import _lib
from ._lib import check, f32c, on_device, ptr, require_cuda, stream_ptr


_SCRATCH_BYTES = None


class _RefineLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, rgb, acc, tgt, occ, coef, den):
        lib = _lib.load()
        require_cuda(rgb, acc, tgt, occ, den)
        rgb, acc, tgt, occ = f32c(rgb), f32c(acc), f32c(tgt), f32c(occ)
        den = f32c(den).reshape(1) if den is not None else None
        n = acc.numel()
        dev = rgb.device
        out = torch.empty(3, device=dev, dtype=torch.float32)
        global _SCRATCH_BYTES
        if _SCRATCH_BYTES is None:
            _SCRATCH_BYTES = lib.snb_refine_loss_scratch_bytes()
        scratch = torch.empty(_SCRATCH_BYTES, dtype=torch.uint8, device=dev)
        with on_device(dev):
            check(lib.snb_refine_loss_fwd(ptr(rgb), ptr(acc), ptr(tgt), ptr(occ), n, float(coef), ptr(den), ptr(out), ptr(scratch),
                                          stream_ptr()), "snb_refine_loss_fwd")
        ctx.save_for_backward(rgb, acc, tgt, occ, scratch)
        ctx.coef = float(coef)
        loss, parts = out[0], out[1:]
        ctx.mark_non_differentiable(parts, out)
        ctx.set_materialize_grads(False)
        return loss, parts, out

    @staticmethod
    def backward(ctx, g_loss, _g_parts, _g_vec):
        lib = _lib.load()
        rgb, acc, tgt, occ, scratch = ctx.saved_tensors
        if g_loss is None:
            return None, None, None, None, None, None
        n = acc.numel()
        g_rgb = torch.empty_like(rgb)
        g_acc = torch.empty_like(acc)
        g_loss = f32c(g_loss)
        with on_device(rgb.device):
            check(lib.snb_refine_loss_bwd(ptr(rgb), ptr(acc), ptr(tgt), ptr(occ), n, ctx.coef, ptr(scratch), ptr(g_loss), ptr(g_rgb),
                                          ptr(g_acc), stream_ptr()), "snb_refine_loss_bwd")
        return g_rgb, g_acc, None, None, None, None


def refine_loss(rgb_rays, acc_trans_rays, rgb_tgt, occ_pixels, loss_occ_coef=0.1, den=None):
    """-> (loss, loss_rgb, loss_occ), 0-dim CUDA tensors; ``loss`` is differentiable to rgb_rays and acc_trans_rays.
    rgb_rays, rgb_tgt (N,3); acc_trans_rays (N,); occ_pixels (N,1).  ``den``: optional caller-computed denominator
    (a 1-element tensor; the ray-sharded mode passes the global ``sum|occ| + 1e-9``)."""
    n = acc_trans_rays.numel()
    if tuple(rgb_rays.shape) != (n, 3) or tuple(rgb_tgt.shape) != (n, 3) or occ_pixels.numel() != n:
        raise ValueError("refine_loss: expected rgb (N,3), acc (N,), tgt (N,3), occ (N,1)")
    loss, parts, _ = _RefineLoss.apply(rgb_rays, acc_trans_rays.reshape(-1), rgb_tgt, occ_pixels.reshape(-1), loss_occ_coef, den)
    return loss, parts[0], parts[1]


def refine_loss_vec(rgb_rays, acc_trans_rays, rgb_tgt, occ_pixels, loss_occ_coef=0.1, den=None):
    """-> (loss, vec): `loss` as refine_loss; `vec` (3,) = [loss, loss_rgb, loss_occ] as the kernel wrote them (one tensor, no
    stacking kernels: what a graph-captured loop keeps to read the losses after a replay)."""
    loss, _, vec = _RefineLoss.apply(rgb_rays, acc_trans_rays.reshape(-1), rgb_tgt, occ_pixels.reshape(-1), loss_occ_coef, den)
    return loss, vec


class _RefineLossBatch(torch.autograd.Function):
    """The refine losses of B objects (each over its own denominator) in one launch per direction (snb_refine_loss_batch_*)."""

    @staticmethod
    def forward(ctx, rgb, acc, tgt, occ, coef):
        lib = _lib.load()
        require_cuda(rgb, acc, tgt, occ)
        rgb, acc, tgt, occ = f32c(rgb), f32c(acc), f32c(tgt), f32c(occ)
        b, n = acc.shape
        dev = rgb.device
        out = torch.empty(b, 3, device=dev, dtype=torch.float32)
        scratch = torch.empty(lib.snb_refine_loss_batch_scratch_bytes(b), dtype=torch.uint8, device=dev)
        with on_device(dev):
            check(lib.snb_refine_loss_batch_fwd(ptr(rgb), ptr(acc), ptr(tgt), ptr(occ), b, n, float(coef), ptr(out), ptr(scratch), stream_ptr()),
                  "snb_refine_loss_batch_fwd")
        ctx.save_for_backward(rgb, acc, tgt, occ, scratch)
        ctx.coef = float(coef)
        loss = out[:, 0]
        ctx.mark_non_differentiable(out)
        ctx.set_materialize_grads(False)
        return loss, out

    @staticmethod
    def backward(ctx, g_loss, _g_out):
        lib = _lib.load()
        rgb, acc, tgt, occ, scratch = ctx.saved_tensors
        if g_loss is None:
            return None, None, None, None, None
        b, n = acc.shape
        g_rgb, g_acc = torch.empty_like(rgb), torch.empty_like(acc)
        g_loss = f32c(g_loss)
        if not g_loss.is_contiguous() or g_loss.stride() != (1,):
            g_loss = g_loss.contiguous()
        with on_device(rgb.device):
            check(lib.snb_refine_loss_batch_bwd(ptr(rgb), ptr(acc), ptr(tgt), ptr(occ), b, n, ctx.coef, ptr(scratch), ptr(g_loss), ptr(g_rgb),
                                                ptr(g_acc), stream_ptr()), "snb_refine_loss_batch_bwd")
        return g_rgb, g_acc, None, None, None


def refine_loss_batch(rgb_rays, acc_trans_rays, rgb_tgt, occ_pixels, loss_occ_coef=0.1):
    """The refine losses (optimizer_nuscenes.py:729-736) of B objects at once: rgb_rays, rgb_tgt (B,N,3); acc_trans_rays (B,N);
    occ_pixels (B,N,1) or (B,N).  -> (loss (B,), parts (B,3) = [loss, loss_rgb, loss_occ] per object); ``loss`` is differentiable."""
    b, n = acc_trans_rays.shape
    if tuple(rgb_rays.shape) != (b, n, 3) or tuple(rgb_tgt.shape) != (b, n, 3) or occ_pixels.numel() != b * n:
        raise ValueError("refine_loss_batch: expected rgb (B,N,3), acc (B,N), tgt (B,N,3), occ (B,N,1)")
    return _RefineLossBatch.apply(rgb_rays, acc_trans_rays, rgb_tgt, occ_pixels.reshape(b, n), loss_occ_coef)
