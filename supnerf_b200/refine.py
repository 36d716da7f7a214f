"""Refine-iteration glue around the render hot path: the test-time optimisation loop of
``optimizer_nuscenes.py:684-769`` (identical in optimizer_kitti.py / optimizer_waymo.py) for ONE object --

    rot_vec, trans_vec -> cam2opt                         (:684-699; the axis-angle map is pytorch3d's, one step upstream of the path)
    render_rays_v2(model, ..., cam2opt, ...)              (:716-726 -> utils.py:435-502)
    loss_rgb + loss_occ_coef * loss_occ ; backward        (:729-737)
    AdamW step on shapecode, texturecode, rot_vec, trans_vec   (:757-769, groups/lrs of :1762-1769)

-- with every host synchronisation removed so that one iteration can be captured in a CUDA graph and replayed: the
reference reads ``cam_pose[:, -1].tolist()`` on the host every iteration to build the shared sample vector
(utils.py:468-469 -> sample_from_rays :154-167); here near/far are evaluated on the device with the same arithmetic
(float64 norm, float32 ``torch.linspace`` formula), and the per-iteration jitter -- drawn on the CPU generator by the
same ``torch.rand(n_samples)`` calls, in the same order -- is uploaded once as a table.  ``ms per refine iteration``
then is kernel time (~25 launches replayed from one graph: 0.21 ms at 32x32 rays on a B200) instead of ~2.4 ms of
Python / launch latency.

The render, loss and their backward are the package's fused kernels (ops.render_shell, losses.refine_loss); with
``fused=True`` (default) the pose map + sample vector and the AdamW update are single launches too (csrc/refine.cu:
snb_refine_pose_fwd/bwd, snb_adamw_step); ``fused=False`` keeps them as torch ops (torch.optim.AdamW(capturable=True)) --
the tests check both against the loop written with the reference-shaped API.  CUDA only."""
import ctypes
import math

import numpy as np
import torch

from . import _lib, losses, models, ops
from . import utils as U
from ._lib import check, f32c, on_device, ptr, require_cuda, stream_ptr


class _PoseAndSamples(torch.autograd.Function):
    """rot_vec, trans_vec -> cam2opt (3,4) and the shared sample vector z (S) in ONE launch (csrc/refine.cu); backward in one."""

    @staticmethod
    def forward(ctx, rot_vec, trans_vec, jitter, opt_cam_pose, obj_diag, n_samples):
        lib = _lib.load()
        require_cuda(rot_vec, trans_vec, jitter)
        rot_vec, trans_vec, jitter = f32c(rot_vec), f32c(trans_vec), f32c(jitter)
        dev = rot_vec.device
        cam = torch.empty(3, 4, device=dev, dtype=torch.float32)
        z = torch.empty(int(n_samples), device=dev, dtype=torch.float32)
        with on_device(dev):
            check(lib.snb_refine_pose_fwd(ptr(rot_vec), ptr(trans_vec), int(bool(opt_cam_pose)), float(obj_diag), int(n_samples), ptr(jitter),
                                          ptr(cam), ptr(z), stream_ptr()), "snb_refine_pose_fwd")
        ctx.save_for_backward(rot_vec, trans_vec)
        ctx.opt_cam_pose = int(bool(opt_cam_pose))
        ctx.mark_non_differentiable(z)
        ctx.set_materialize_grads(False)
        return cam, z

    @staticmethod
    def backward(ctx, g_cam, _g_z):
        lib = _lib.load()
        rot_vec, trans_vec = ctx.saved_tensors
        if g_cam is None:
            return None, None, None, None, None, None
        g_cam = f32c(g_cam)
        g_rot, g_trans = torch.empty_like(rot_vec), torch.empty_like(trans_vec)
        with on_device(rot_vec.device):
            check(lib.snb_refine_pose_bwd(ptr(rot_vec), ptr(trans_vec), ctx.opt_cam_pose, ptr(g_cam), ptr(g_rot), ptr(g_trans), stream_ptr()),
                  "snb_refine_pose_bwd")
        return g_rot, g_trans, None, None, None, None


class FusedAdamW:
    """torch.optim.AdamW's update for a handful of small tensors in ONE launch (snb_adamw_step); state lives on the device."""

    def __init__(self, groups, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2):
        self.params = [g["params"] for g in groups]
        self.lrs = [float(g["lr"]) for g in groups]
        self.betas, self.eps, self.weight_decay = betas, eps, weight_decay
        dev = self.params[0].device
        self.m = [torch.zeros_like(p) for p in self.params]
        self.v = [torch.zeros_like(p) for p in self.params]
        self.step_t = torch.zeros((), device=dev, dtype=torch.float32)
        n = len(self.params)
        self._sizes = (ctypes.c_int32 * n)(*[p.numel() for p in self.params])
        self._lrs = (ctypes.c_float * n)(*self.lrs)

    def zero_grad(self, set_to_none=True):
        """set_to_none (default): the next backward ASSIGNS the gradients instead of accumulating into zero-filled ones: no fill
        and no add kernel per parameter (inside a CUDA graph the assigned tensors keep their addresses across replays)."""
        for p in self.params:
            if p.grad is not None:
                if set_to_none:
                    p.grad = None
                else:
                    p.grad.zero_()

    def reset(self):
        for t in self.m + self.v:
            t.zero_()
        self.step_t.zero_()

    def step(self):
        lib = _lib.load()
        n = len(self.params)
        arr = lambda ts: (ctypes.c_void_p * n)(*[t.data_ptr() for t in ts])  # noqa: E731
        with on_device(self.params[0].device):
            check(lib.snb_adamw_step(n, arr(self.params), arr([p.grad for p in self.params]), arr(self.m), arr(self.v), self._sizes, self._lrs,
                                     self.betas[0], self.betas[1], self.eps, self.weight_decay, ptr(self.step_t), stream_ptr()), "snb_adamw_step")


def axis_angle_to_matrix(v):
    """Rodrigues' formula, R = I + sin(t)/t [v]x + (1-cos(t))/t^2 [v]x^2 -- what pytorch3d.transforms.axis_angle_to_matrix
    evaluates (via quaternions); pytorch3d itself is not part of the reference tree (parity unpinned there)."""
    t = torch.sqrt((v * v).sum() + 1e-20)
    zero = torch.zeros((), device=v.device, dtype=v.dtype)
    kx = torch.stack([torch.stack([zero, -v[2], v[1]]), torch.stack([v[2], zero, -v[0]]), torch.stack([-v[1], v[0], zero])])
    return torch.eye(3, device=v.device, dtype=v.dtype) + torch.sin(t) / t * kx + (1 - torch.cos(t)) / (t * t) * (kx @ kx)


def shell_samples_on_device(cam_pose, obj_diag, n_samples, jitter):
    """utils.sample_from_rays' shared vector (utils.py:154-167 with near/far of :468-469) without leaving the device.
    near/far: float64 like the reference's python floats; linspace: torch's float32 formula (step = (end-start)/(n-1);
    first half start + i*step, second half end - (n-1-i)*step); jitter (S,) = the torch.rand(n_samples) draw."""
    n = torch.linalg.vector_norm(cam_pose[:, -1].detach().double())
    half = float(obj_diag) / 2          # obj_diag is np.float32 in the reference: exactly representable
    near, far = n - half, n + half
    dist = (far - near) / (2 * n_samples)
    start, end = (near + dist).float(), (far - dist).float()
    step = (end - start) / (n_samples - 1) if n_samples > 1 else torch.zeros_like(start)
    i = torch.arange(n_samples, device=cam_pose.device, dtype=torch.float32)
    lo = start + step * i
    hi = end - step * (n_samples - 1 - i)
    z = torch.where(i < n_samples // 2, lo, hi)
    return z + jitter * ((far - near) / (2 * n_samples)).float()


class ObjectRefiner:
    """One object's refine loop.  ``step()`` runs one iteration eagerly; ``capture()`` records it into a CUDA graph and
    ``run(n)`` replays it.  State (codes, pose parameters, AdamW moments) lives in this object's tensors."""

    def __init__(self, model, device, img, mask_occ, K, roi, obj_diag, shapecode, texturecode, rot_vec, trans_vec, n_samples=64,
                 im_sz=32, lr_shape=0.02, lr_texture=0.02, lr_pose=0.01, loss_occ_coef=0.1, shapenet_obj_cood=True,
                 opt_cam_pose=False, max_iters=100, fused=True, lidar_xy=None):
        if not isinstance(model, models._DecoderBase):
            raise TypeError("ObjectRefiner needs a supnerf_b200 decoder")
        self.model, self.device = model, torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("supnerf_b200 has no CPU path")
        dev = self.device
        self.n_samples, self.im_sz, self.coef = int(n_samples), int(im_sz), float(loss_occ_coef)
        self.obj_diag, self.swap, self.opt_cam_pose = np.float32(obj_diag), bool(shapenet_obj_cood), bool(opt_cam_pose)
        self.K = K.to(dev, torch.float32)
        self.px, self.py = U._pixel_grid_on(dev, roi, [im_sz, im_sz])
        img, mask_occ = U._resize_targets(img, mask_occ, im_sz)          # utils.py:448-453
        self.rgb_tgt = img.reshape(-1, 3).to(dev).contiguous()
        self.occ = mask_occ.reshape(-1, 1).to(dev).contiguous()
        self.shapecode = shapecode.detach().to(dev).clone().requires_grad_()
        self.texturecode = texturecode.detach().to(dev).clone().requires_grad_()
        self.rot_vec = rot_vec.detach().to(dev).reshape(3).clone().requires_grad_()
        self.trans_vec = trans_vec.detach().to(dev).reshape(3).clone().requires_grad_()
        # fused = True: pose map + sample vector and the AdamW update as single launches (csrc/refine.cu); False: torch ops
        self.fused = bool(fused)
        if self.fused:
            self.opt = FusedAdamW([{"params": self.shapecode, "lr": lr_shape}, {"params": self.texturecode, "lr": lr_texture},
                                   {"params": self.rot_vec, "lr": lr_pose}, {"params": self.trans_vec, "lr": lr_pose}])
        else:
            self.opt = torch.optim.AdamW([{"params": [self.shapecode], "lr": lr_shape}, {"params": [self.texturecode], "lr": lr_texture},
                                          {"params": [self.rot_vec], "lr": lr_pose}, {"params": [self.trans_vec], "lr": lr_pose}],
                                         capturable=True)
        # Per-iteration evaluation the reference interleaves with the optimisation (optimizer_nuscenes.py:740-769): the PSNR loss over
        # the object mask only (loss_rgb2) and a no-grad depth render of the pixels that carry a lidar return (render_rays_specified).
        # lidar_xy = (x_vec, y_vec): integer pixel offsets inside the crop (np.where of the depth map, :759); None = not evaluated.
        self.lidar = None
        if lidar_xy is not None:
            x_vec, y_vec = np.asarray(lidar_xy[0]), np.asarray(lidar_xy[1])
            self.n_lidar = int(x_vec.shape[0])
            if self.n_lidar > 0:
                lx = torch.from_numpy(x_vec + int(roi[0])).reshape(-1).to(dev, torch.float32)
                ly = torch.from_numpy(y_vec + int(roi[1])).reshape(-1).to(dev, torch.float32)
                pad = (-self.n_lidar) % 2            # the bf16 decoder works on 128-row tiles: an even ray count at 64 samples
                if pad:
                    lx, ly = torch.cat([lx, lx[-1:]]), torch.cat([ly, ly[-1:]])
                self.lidar = (lx.contiguous(), ly.contiguous())
        self.occ_pos = self.occ.clamp_min(0.0)       # mask_rgb: occ_pixels with the negative (background) entries zeroed (:741-742)
        self.loss_rgb2 = torch.zeros((), device=dev)
        self.depth_pred = torch.zeros(0, device=dev)
        # the reference draws torch.rand(n_samples) on the CPU generator once per render call (utils.py:164): per iteration one draw
        # for the optimised render and, when the lidar pixels are evaluated, a second one for render_rays_specified -- same calls, same order
        per_it = 2 if self.lidar is not None else 1
        draws = torch.stack([torch.rand(self.n_samples) for _ in range(max_iters * per_it)]).to(dev)
        self.jitter = draws[0::per_it].contiguous()
        self.jitter_lidar = draws[1::per_it].contiguous() if self.lidar is not None else None
        self.it = torch.zeros((), dtype=torch.long, device=dev)
        self.loss = torch.zeros(3, device=dev)
        self.graph = None

    def cam2opt(self):
        rot = axis_angle_to_matrix(self.rot_vec)
        t = self.trans_vec.unsqueeze(-1)
        if not self.opt_cam_pose:                     # optimizer_nuscenes.py:695-697
            rot = rot.transpose(-2, -1)
            t = -rot @ t
        return torch.cat((rot, t), dim=-1)

    def step(self):
        """One iteration (no host synchronisation anywhere)."""
        if self.shapecode.grad is not None:
            self.opt.zero_grad() if self.fused else self.opt.zero_grad(set_to_none=False)
        jit = self.jitter.index_select(0, self.it.reshape(1)).reshape(-1)
        if self.fused:
            cam, z = _PoseAndSamples.apply(self.rot_vec, self.trans_vec, jit, self.opt_cam_pose, self.obj_diag, self.n_samples)
        else:
            cam = self.cam2opt()
            z = shell_samples_on_device(cam, self.obj_diag, self.n_samples, jit)
        prec = self.model.precision or models.get_default_precision()
        rgb, dep, acc = ops.render_shell(self.model._handle(self.device), prec, self.n_samples, float(self.obj_diag), self.swap,
                                         self.px, self.py, self.K, cam, z, self.shapecode, self.texturecode, self.model._weights())
        loss, vec = losses.refine_loss_vec(rgb, acc, self.rgb_tgt, self.occ, self.coef)
        loss.backward()
        with torch.no_grad():       # the per-iteration evaluation (optimizer_nuscenes.py:740-769), before the parameter update as there
            self.loss_rgb2 = losses.refine_loss(rgb.detach(), acc.detach(), self.rgb_tgt, self.occ_pos, 0.0)[1]
            if self.lidar is not None:
                jit2 = self.jitter_lidar.index_select(0, self.it.reshape(1)).reshape(-1)
                if self.fused:
                    cam2, z2 = _PoseAndSamples.apply(self.rot_vec.detach(), self.trans_vec.detach(), jit2, self.opt_cam_pose, self.obj_diag, self.n_samples)
                else:
                    cam2 = cam.detach()
                    z2 = shell_samples_on_device(cam2, self.obj_diag, self.n_samples, jit2)
                _, dep2, _ = ops.render_shell(self.model._handle(self.device), prec, self.n_samples, float(self.obj_diag), self.swap,
                                              self.lidar[0], self.lidar[1], self.K, cam2, z2, self.shapecode.detach(), self.texturecode.detach(),
                                              self.model._weights())
                self.depth_pred = dep2[:self.n_lidar]
        self.opt.step()
        self.loss = vec   # [loss, loss_rgb, loss_occ] as the loss kernel wrote them (graph mode: a static tensor of the graph's pool)
        self.it += 1
        return self.loss

    def capture(self, warmup=3):
        """Warm up on a side stream (allocator, weight packing, AdamW state), then capture one iteration."""
        s = torch.cuda.Stream(device=self.device)
        s.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(s):
            snapshot = [t.detach().clone() for t in (self.shapecode, self.texturecode, self.rot_vec, self.trans_vec)]
            for _ in range(warmup):
                self.step()
            # undo the warm-up updates: parameters, AdamW moments and step counters, iteration counter
            with torch.no_grad():
                for t, v in zip((self.shapecode, self.texturecode, self.rot_vec, self.trans_vec), snapshot):
                    t.copy_(v)
                if self.fused:
                    self.opt.reset()
                else:
                    for st in self.opt.state.values():
                        for v in st.values():
                            if torch.is_tensor(v):
                                v.zero_()
                self.it.zero_()
        torch.cuda.current_stream(self.device).wait_stream(s)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.step()
        # the capture itself does not execute: state is still the pre-capture one
        return self

    def run(self, iters):
        """Run `iters` iterations (graph replays if captured).  Returns the last [loss, loss_rgb, loss_occ] (device tensor)."""
        if int(iters) > self.jitter.shape[0]:
            raise ValueError("iters exceeds the pre-drawn jitter table (max_iters)")
        for _ in range(int(iters)):
            if self.graph is not None:
                self.graph.replay()
            else:
                self.step()
        return self.loss


def run_objects(refiners, iters, n_streams=4):
    """Refine several INDEPENDENT objects side by side (object-parallel: own pose, codes, rays, loss and AdamW state):
    iteration k of every object is issued round-robin over `n_streams` CUDA streams, so one object's small kernels (pose map,
    sampler, compositing, loss, AdamW) run under another object's decoder kernel instead of leaving the GPU idle between two
    decoder launches.  No host synchronisation; the caller's current stream waits for all of them at the end.
    Returns the objects' last [loss, loss_rgb, loss_occ] tensors."""
    if not refiners:
        return []
    dev = refiners[0].device
    main = torch.cuda.current_stream(dev)
    streams = [torch.cuda.Stream(device=dev) for _ in range(max(1, min(int(n_streams), len(refiners))))]
    for s in streams:
        s.wait_stream(main)
    for _ in range(int(iters)):
        for i, r in enumerate(refiners):
            with torch.cuda.stream(streams[i % len(streams)]):
                r.run(1)
    for s in streams:
        main.wait_stream(s)
    return [r.loss for r in refiners]


class ObjectGroup:
    """Several INDEPENDENT objects (config C3: a GPU's 4 of the 32 objects) refined side by side from ONE CUDA graph per iteration:
    the capture forks the objects' iterations onto one stream each and joins them, so a replay is a single graph launch whose
    branches the device schedules concurrently (one object's ~30 small kernels under another's decoder kernel) -- instead of one
    graph launch per object per iteration issued round-robin from the host (run_objects)."""

    def __init__(self, refiners):
        if not refiners:
            raise ValueError("ObjectGroup needs at least one refiner")
        self.refiners = list(refiners)
        self.device = self.refiners[0].device
        self.graph = None

    def capture(self, warmup=3):
        dev = self.device
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):   # warm-up (allocator, weight packing, AdamW state), then undo it -- as ObjectRefiner.capture does
            for r in self.refiners:
                snapshot = [t.detach().clone() for t in (r.shapecode, r.texturecode, r.rot_vec, r.trans_vec)]
                for _ in range(warmup):
                    r.step()
                with torch.no_grad():
                    for t, v in zip((r.shapecode, r.texturecode, r.rot_vec, r.trans_vec), snapshot):
                        t.copy_(v)
                    if r.fused:
                        r.opt.reset()
                    else:
                        for st in r.opt.state.values():
                            for v in st.values():
                                if torch.is_tensor(v):
                                    v.zero_()
                    r.it.zero_()
        torch.cuda.current_stream(dev).wait_stream(side)
        streams = [torch.cuda.Stream(device=dev) for _ in self.refiners]
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            main = torch.cuda.current_stream(dev)
            for s_, r in zip(streams, self.refiners):
                s_.wait_stream(main)                 # fork
                with torch.cuda.stream(s_):
                    r.step()
            for s_ in streams:
                main.wait_stream(s_)                 # join
        return self

    def run(self, iters):
        """`iters` iterations of every object.  -> the objects' last [loss, loss_rgb, loss_occ] tensors."""
        if any(int(iters) > r.jitter.shape[0] for r in self.refiners):
            raise ValueError("iters exceeds the pre-drawn jitter table (max_iters)")
        for _ in range(int(iters)):
            if self.graph is not None:
                self.graph.replay()
            else:
                for r in self.refiners:
                    r.step()
        return [r.loss for r in self.refiners]


class _PoseAndSamplesBatch(torch.autograd.Function):
    """(B,3) rot_vec, trans_vec -> cam2opt (B,3,4) + every object's shared sample vector z (B,S) [+ a second one, z2, from another jitter
    table] in ONE launch (csrc/refine.cu: snb_refine_pose_batch_fwd); the jitter row is picked on the device by the optimiser's step counter."""

    @staticmethod
    def forward(ctx, rot_vec, trans_vec, jitter, jitter2, step_t, obj_diag, opt_cam_pose, n_samples):
        lib = _lib.load()
        require_cuda(rot_vec, trans_vec, jitter, step_t, obj_diag)
        rot_vec, trans_vec = f32c(rot_vec), f32c(trans_vec)
        b, t, s = jitter.shape
        dev = rot_vec.device
        cam = torch.empty(b, 3, 4, device=dev, dtype=torch.float32)
        z = torch.empty(b, s, device=dev, dtype=torch.float32)
        z2 = torch.empty(b, s, device=dev, dtype=torch.float32) if jitter2 is not None else None
        with on_device(dev):
            check(lib.snb_refine_pose_batch_fwd(ptr(rot_vec), ptr(trans_vec), int(b), int(bool(opt_cam_pose)), ptr(obj_diag), int(n_samples),
                                                ptr(jitter), ptr(jitter2), ptr(step_t), int(t), ptr(cam), ptr(z), ptr(z2), stream_ptr()),
                  "snb_refine_pose_batch_fwd")
        ctx.save_for_backward(rot_vec, trans_vec)
        ctx.opt_cam_pose = int(bool(opt_cam_pose))
        ctx.set_materialize_grads(False)
        if z2 is None:
            ctx.mark_non_differentiable(z)
            return cam, z
        ctx.mark_non_differentiable(z, z2)
        return cam, z, z2

    @staticmethod
    def backward(ctx, g_cam, *_):
        lib = _lib.load()
        rot_vec, trans_vec = ctx.saved_tensors
        if g_cam is None:
            return (None,) * 8
        g_cam = f32c(g_cam)
        g_rot, g_trans = torch.empty_like(rot_vec), torch.empty_like(trans_vec)
        with on_device(rot_vec.device):
            check(lib.snb_refine_pose_batch_bwd(ptr(rot_vec), ptr(trans_vec), int(rot_vec.shape[0]), ctx.opt_cam_pose, ptr(g_cam), ptr(g_rot),
                                                ptr(g_trans), stream_ptr()), "snb_refine_pose_batch_bwd")
        return (g_rot, g_trans) + (None,) * 6


class BatchRefiner:
    """The refine loops of B INDEPENDENT objects (one GPU's share of config C3) as ONE launch set per iteration: every stage of
    ObjectRefiner.step -- pose map + sample vectors, the utils.py-stack render, the losses, their backward, the per-iteration
    evaluation (loss_rgb2, lidar-pixel depths) and the AdamW update -- runs once over all objects (ops.render_shell_batch,
    losses.refine_loss_batch, snb_refine_pose_batch_*), ~25 launches per iteration whatever B is instead of ~30 per object.  The objects
    stay independent: per-object pose, codes, rays, targets, jitter draws, losses (the summed loss has block-diagonal gradients) and
    elementwise AdamW state with one shared step count.

    Built from ObjectRefiners (which hold the per-object inputs and the jitter tables drawn in the reference's order); they must agree on
    n_samples, crop size, learning rates, loss coefficient, axis convention and decoder.  ``write_back()`` copies the optimised
    state into them."""

    def __init__(self, refiners):
        if not refiners:
            raise ValueError("BatchRefiner needs at least one refiner")
        r0 = refiners[0]
        same = lambda f: all(f(r) == f(r0) for r in refiners)   # noqa: E731
        if not (same(lambda r: (r.n_samples, r.im_sz, r.coef, r.swap, r.opt_cam_pose, r.fused, tuple(r.opt.lrs) if r.fused else None,
                                r.jitter.shape[0])) and all(r.model is r0.model and r.device == r0.device for r in refiners)):
            raise ValueError("BatchRefiner: the objects must share the decoder, device, n_samples, crop size, loss and optimiser settings")
        if not r0.fused:
            raise ValueError("BatchRefiner builds on the fused refiners (fused=True)")
        self.refiners = list(refiners)
        self.model, self.device = r0.model, r0.device
        dev = self.device
        self.n_samples, self.coef, self.swap, self.opt_cam_pose = r0.n_samples, r0.coef, r0.swap, r0.opt_cam_pose
        st = lambda f: torch.stack([f(r) for r in refiners]).contiguous()   # noqa: E731
        self.px, self.py = st(lambda r: r.px.reshape(-1)), st(lambda r: r.py.reshape(-1))
        self.K = st(lambda r: r.K)
        self.rgb_tgt, self.occ = st(lambda r: r.rgb_tgt), st(lambda r: r.occ.reshape(-1))
        self.occ_pos = self.occ.clamp_min(0.0)
        self.obj_diag = torch.tensor([float(r.obj_diag) for r in refiners], device=dev, dtype=torch.float32)
        self.shapecode = st(lambda r: r.shapecode.detach().reshape(-1)).requires_grad_()
        self.texturecode = st(lambda r: r.texturecode.detach().reshape(-1)).requires_grad_()
        self.rot_vec = st(lambda r: r.rot_vec.detach()).requires_grad_()
        self.trans_vec = st(lambda r: r.trans_vec.detach()).requires_grad_()
        lrs = r0.opt.lrs
        self.opt = FusedAdamW([{"params": self.shapecode, "lr": lrs[0]}, {"params": self.texturecode, "lr": lrs[1]},
                               {"params": self.rot_vec, "lr": lrs[2]}, {"params": self.trans_vec, "lr": lrs[3]}],
                              betas=r0.opt.betas, eps=r0.opt.eps, weight_decay=r0.opt.weight_decay)
        self.jitter = st(lambda r: r.jitter)                                  # (B, T, S)
        self.max_iters = int(self.jitter.shape[1])
        # lidar pixels: every object's list padded to one length (a whole number of 256-row decoder super tiles: 4 rays at 64 samples) with copies of
        # its last pixel; the copies' depths are sliced off again.  Objects without lidar pixels render a dummy pixel and report nothing.
        self.n_lidar = [r.n_lidar if r.lidar is not None else 0 for r in refiners]
        self.lidar, self.jitter_lidar = None, None
        if any(self.n_lidar):
            q = 256 // math.gcd(256, self.n_samples)          # rays per 256 decoder rows (4 at 64 samples)
            width = -(-max(self.n_lidar) // q) * q
            lx, ly, jl = [], [], []
            for r, n in zip(refiners, self.n_lidar):
                if n:
                    x, y = r.lidar[0][:n], r.lidar[1][:n]
                    jl.append(r.jitter_lidar)
                else:
                    x, y = r.px.reshape(-1)[:1], r.py.reshape(-1)[:1]
                    jl.append(torch.zeros_like(r.jitter))
                lx.append(torch.cat([x, x[-1:].expand(width - x.numel())]))
                ly.append(torch.cat([y, y[-1:].expand(width - y.numel())]))
            self.lidar = (torch.stack(lx).contiguous(), torch.stack(ly).contiguous())
            self.jitter_lidar = torch.stack(jl).contiguous()
        b = len(refiners)
        self._ones = torch.ones(b, device=dev)
        self.loss = torch.zeros(b, 3, device=dev)
        self.loss_rgb2 = torch.zeros(b, device=dev)
        self.depth_pred = None
        self.it = 0              # host-side count of issued iterations (the device-side row index is the optimiser's step counter)
        self.graph = None

    def _params(self):
        return (self.shapecode, self.texturecode, self.rot_vec, self.trans_vec)

    def step(self):
        """One iteration of every object (no host synchronisation anywhere)."""
        if self.shapecode.grad is not None:
            self.opt.zero_grad()
        out = _PoseAndSamplesBatch.apply(self.rot_vec, self.trans_vec, self.jitter, self.jitter_lidar, self.opt.step_t, self.obj_diag,
                                         self.opt_cam_pose, self.n_samples)
        cam, z = out[0], out[1]
        prec = self.model.precision or models.get_default_precision()
        h, w = self.model._handle(self.device), self.model._weights()
        rgb, _dep, acc = ops.render_shell_batch(h, prec, self.n_samples, self.swap, self.px, self.py, self.K, cam, z, self.obj_diag,
                                                self.shapecode, self.texturecode, w)
        loss, parts = losses.refine_loss_batch(rgb, acc, self.rgb_tgt, self.occ, self.coef)
        torch.autograd.backward(loss, self._ones)
        with torch.no_grad():       # the per-iteration evaluation (optimizer_nuscenes.py:740-769), before the parameter update as there
            self.loss_rgb2 = losses.refine_loss_batch(rgb.detach(), acc.detach(), self.rgb_tgt, self.occ_pos, 0.0)[1][:, 1]
            if self.lidar is not None:
                _, dep2, _ = ops.render_shell_batch(h, prec, self.n_samples, self.swap, self.lidar[0], self.lidar[1], self.K, cam.detach(),
                                                    out[2], self.obj_diag, self.shapecode.detach(), self.texturecode.detach(), w)
                self.depth_pred = dep2           # (B, width): object b's lidar depths are depth_pred[b, :n_lidar[b]]
        self.opt.step()
        self.loss = parts           # (B,3) = [loss, loss_rgb, loss_occ] per object
        return self.loss

    def capture(self, warmup=3):
        """Warm up on a side stream (allocator, weight packing), undo the warm-up updates, then capture one iteration."""
        s = torch.cuda.Stream(device=self.device)
        s.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(s):
            snapshot = [t.detach().clone() for t in self._params()]
            for _ in range(warmup):
                self.step()
            with torch.no_grad():
                for t, v in zip(self._params(), snapshot):
                    t.copy_(v)
                self.opt.reset()
        torch.cuda.current_stream(self.device).wait_stream(s)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.step()
        return self

    def run(self, iters):
        """`iters` iterations of every object (graph replays if captured).  -> (B,3) last [loss, loss_rgb, loss_occ] per object."""
        if self.it + int(iters) > self.max_iters:
            raise ValueError("iters exceeds the pre-drawn jitter tables (max_iters)")
        for _ in range(int(iters)):
            if self.graph is not None:
                self.graph.replay()
            else:
                self.step()
        self.it += int(iters)
        return self.loss

    def lidar_depths(self):
        """Per object: the depths at its lidar pixels from the last iteration's evaluation (None for an object without lidar pixels)."""
        if self.depth_pred is None:
            return [None] * len(self.refiners)
        return [self.depth_pred[b, :n] if n else None for b, n in enumerate(self.n_lidar)]

    def write_back(self):
        """Copy the optimised codes and pose parameters into the ObjectRefiners this batch was built from."""
        with torch.no_grad():
            for b, r in enumerate(self.refiners):
                r.shapecode.copy_(self.shapecode[b].reshape(r.shapecode.shape))
                r.texturecode.copy_(self.texturecode[b].reshape(r.texturecode.shape))
                r.rot_vec.copy_(self.rot_vec[b])
                r.trans_vec.copy_(self.trans_vec[b])
                r.loss = self.loss[b]
                r.loss_rgb2 = self.loss_rgb2[b]
        return self.refiners
