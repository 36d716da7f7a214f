"""Drop-in for the reference's ``src/renderer.py``: ``NeRFRenderer`` (same constructor, methods, argument
orders and return tuples), ``volume_rendering3`` and ``render_rays_v3``.  Ray generation, the ray/AABB
slab test, stratified sampling, the decoder and the compositing all run in the sm_100a kernels of
libsupnerf_b200.so and stay differentiable to the pose, the latents and the weights."""
import random

import numpy as np
import torch

from . import models, ops
from . import utils as U
from .utils import get_rays, get_rays_specified, ray_box_intersection, ray_box_intersection_tensor  # noqa: F401

_Z_STEPS = {}


def _z_steps_on(device, n_samples):
    """torch.linspace(0, 1 - 1/S, S) (renderer.py:36-38) on `device`, cached: a pure function of (S, device)."""
    key = (str(device), int(n_samples))
    z = _Z_STEPS.get(key)
    if z is None:
        step = 1.0 / n_samples
        z = torch.linspace(0, 1 - step, n_samples, device=device)
        _Z_STEPS[key] = z
    return z


class RayBatch:
    """Device-resident constants of a batch of object crops (NeRFRenderer.make_batch): px, py (B,N) pixel coordinates, K (B,3,3),
    box (B,4) = {diag/2, l/diag, w/diag, h/diag}, rgb_tgt (B,N,3), occ_pixels (B,N,1)."""
    __slots__ = ("px", "py", "K", "box", "rgb_tgt", "occ_pixels")

    def __init__(self, px, py, K, box, rgb_tgt, occ_pixels):
        self.px, self.py, self.K, self.box, self.rgb_tgt, self.occ_pixels = px, py, K, box, rgb_tgt, occ_pixels


FUSED_RENDER = True  # render_rays / render_rays_specified go through ops.render_box (one autograd node); False = staged ops


class NeRFRenderer(torch.nn.Module):
    def __init__(self, n_samples=64, noise_std=0.0, white_bkgd=True):
        """renderer.py:16-25 (noise_std is stored and never used, as in the reference)."""
        super().__init__()
        self.n_samples = n_samples
        self.noise_std = noise_std
        self.white_bkgd = white_bkgd

    def sample_from_ray(self, rays):
        """renderer.py:27-41."""
        return U.sample_from_rays_v2(rays, self.n_samples)

    def volume_render(self, sigmas, rgbs, z_vals):
        """renderer.py:43-65: sigmas (N,S), rgbs (N,S,3), z_vals (N,S)."""
        return U._composite_any(sigmas, rgbs, z_vals, self.white_bkgd, True)

    def volume_render_batch(self, sigmas, rgbs, z_vals):
        """renderer.py:67-89: sigmas (B,n,S[,1]), rgbs (B,n,S,3), z_vals (B,n,S).  With white_bkgd the reference sums
        the pixel alpha over the wrong dimension (renderer.py:86, only shape-valid when n == S); that branch is
        reproduced literally on top of the kernel's black-background result."""
        rgb, dep, acc = U._composite_any(sigmas, rgbs, z_vals, False, True)
        dep = dep.unsqueeze(-1)  # renderer.py:82 keeps a trailing 1
        if self.white_bkgd:
            s = sigmas.squeeze(-1) if sigmas.dim() == rgbs.dim() else sigmas
            deltas = torch.cat([z_vals[..., 1:] - z_vals[..., :-1], torch.ones_like(z_vals[..., :1]) * 1e10], -1)
            alphas = 1 - torch.exp(-torch.relu(s) * deltas)
            trans = torch.cat([torch.ones_like(alphas[..., :1]), 1 - alphas + 1e-10], -1)
            weights = alphas * torch.cumprod(trans, -1)[..., :-1]
            pix_alpha = weights.sum(dim=1)
            rgb = rgb + 1 - pix_alpha.unsqueeze(-1)
        return rgb, dep, acc

    def prepare_sampled_rays(self, rays_o, viewdir, obj_sz):
        """renderer.py:91-115 -> xyz (N,S,3), viewdir (N,S,3), z_vals (N,S), intersect (N,) bool."""
        diag, half = ops.box_constants(obj_sz)
        dev = U._device_of(rays_o, viewdir)
        n = rays_o.shape[0]
        step = 1.0 / self.n_samples
        z_steps = torch.linspace(0, 1 - step, self.n_samples, device=dev)
        jitter = torch.rand_like(z_steps.unsqueeze(0).repeat(n, 1))  # same call/shape as renderer.py:39-40
        xyz, vd, z_vals, hit = ops.sample_box(rays_o.to(dev), viewdir.to(dev), z_steps, jitter, diag / 2, half)
        if not rays_o.is_cuda:
            xyz, vd, z_vals, hit = xyz.to(rays_o.device), vd.to(rays_o.device), z_vals.to(rays_o.device), hit.to(rays_o.device)
        return xyz, vd, z_vals, hit

    def _render_fused(self, model, device, px, py, K, cam_pose, obj_sz, shapecode, texturecode, jitter=None):
        """get_rays -> prepare_sampled_rays -> model -> volume_render (renderer.py:125,153-165) as ONE autograd node /
        two C-ABI calls (ops.render_box).  Same RNG consumption as the staged path: one torch.rand_like of (N,S)."""
        device = torch.device(device)
        diag, half = ops.box_constants(obj_sz)
        n = px.numel()
        z_steps = _z_steps_on(device, self.n_samples)
        if jitter is None:
            jitter = torch.rand_like(torch.empty(n, self.n_samples, device=device))  # renderer.py:39-40
        prec = model.precision or models.get_default_precision()
        rgb, dep, acc, _hit = ops.render_box(model._handle(device), prec, self.n_samples, self.white_bkgd, diag / 2, half,
                                             px, py, K.to(device, non_blocking=True), cam_pose.to(device, non_blocking=True),
                                             z_steps, jitter, shapecode.to(device, non_blocking=True),
                                             texturecode.to(device, non_blocking=True), model._weights())
        return rgb, dep, acc

    def make_batch(self, device, imgs, masks_occ, obj_szs, Ks, rois, im_sz=64):
        """The per-batch constants of ``render_rays_batch``: pixel grids, targets, occupancy masks and box constants of B object
        crops, on the device (a refine-style loop re-renders the same crops every iteration: build once, render many times).
        imgs (B,h,w,3) / masks_occ (B,h,w,1) stacked tensors or lists of per-object crops, resampled to im_sz x im_sz like
        renderer.py:127-133 (stacked tensors that already have that size take one copy for the whole batch)."""
        device = torch.device(device)
        if device.type != "cuda":
            raise RuntimeError("supnerf_b200 has no CPU path")
        b = len(rois)
        boxes, pxs, pys = [], [], []
        for i in range(b):
            diag, half = ops.box_constants(obj_szs[i])
            boxes.append([float(diag / 2), float(half[0]), float(half[1]), float(half[2])])
            px, py = U._pixel_grid_on(device, rois[i], [im_sz, im_sz])
            pxs.append(px)
            pys.append(py)
        if torch.is_tensor(imgs) and torch.is_tensor(masks_occ) and imgs.dim() == 4 and tuple(imgs.shape[1:3]) == (im_sz, im_sz) \
                and tuple(masks_occ.shape[1:3]) == (im_sz, im_sz):
            # torchvision's Resize to the size the crop already has returns its input: only the mask's int32 round trip remains
            rgb_tgt = imgs.to(device, non_blocking=True).reshape(b, -1, 3)
            occ_pixels = masks_occ.to(device, non_blocking=True).reshape(b, -1, 1).type(torch.int32).type(torch.float32)
        else:
            tgts, occs = [], []
            for i in range(b):
                img, mask = U._resize_targets(imgs[i], masks_occ[i], im_sz)
                tgts.append(img.reshape(-1, 3))
                occs.append(mask.reshape(-1, 1))
            rgb_tgt = torch.stack(tgts).to(device, non_blocking=True)
            occ_pixels = torch.stack(occs).to(device, non_blocking=True)
        Ks = torch.as_tensor(Ks) if not torch.is_tensor(Ks) else Ks
        if Ks.dim() == 2:
            Ks = Ks.unsqueeze(0).expand(b, 3, 3)
        return RayBatch(torch.stack(pxs), torch.stack(pys), Ks.to(device, torch.float32, non_blocking=True).contiguous(),
                        torch.tensor(boxes, dtype=torch.float32).to(device, non_blocking=True), rgb_tgt, occ_pixels)

    def render_batch(self, model, batch, cam_poses, shapecodes, texturecodes, jitter=None, fused_sampler=False):
        """B objects of a prepared ``RayBatch`` in ONE launch set (csrc/render_batch.cu).  cam_poses (B,3,4); shapecodes /
        texturecodes (B,D).  -> rgb (B,N,3), depth (B,N), acc (B,N).  One torch.rand_like of (B,N,S) per call."""
        if not isinstance(model, models._DecoderBase):
            raise RuntimeError("render_batch needs a supnerf_b200 decoder")
        device = batch.px.device
        b, n = batch.px.shape
        if jitter is None:
            jitter = torch.rand_like(torch.empty(b, n, self.n_samples, device=device))
        rgb, dep, acc, _hit = ops.render_box_batch(model._handle(device), self.n_samples, self.white_bkgd, batch.px, batch.py, batch.K,
                                                   cam_poses.to(device, non_blocking=True), batch.box, _z_steps_on(device, self.n_samples),
                                                   jitter, shapecodes.to(device, non_blocking=True),
                                                   texturecodes.to(device, non_blocking=True), model._weights(), fused_sampler=fused_sampler,
                                                   precision=model.precision or models.get_default_precision())
        return rgb, dep, acc

    def render_rays_batch(self, model, device, imgs, masks_occ, cam_poses, obj_szs, Ks, rois, shapecodes, texturecodes, im_sz=64,
                          jitter=None, fused_sampler=False):
        """``render_rays`` (renderer.py:117-167, ``n_rays=None``) of B objects in ONE launch set -- what the reference does with a
        Python loop over the objects of a scene (optimizer_nuscenes.py:716-726; configs[1]: 16 objects per step).
        cam_poses (B,3,4); obj_szs B x (w,l,h); Ks (B,3,3) or one (3,3); rois B x (4,); shapecodes / texturecodes (B,D).
        -> rgb (B,N,3), depth (B,N), acc (B,N), rgb_tgt (B,N,3), occ_pixels (B,N,1).  Frozen weights; the model's precision picks the
        bf16 or the fp32-grade (split-precision) tensor-core decoder (no CPU or per-object fallback)."""
        batch = self.make_batch(device, imgs, masks_occ, obj_szs, Ks, rois, im_sz)
        rgb, dep, acc = self.render_batch(model, batch, cam_poses, shapecodes, texturecodes, jitter, fused_sampler)
        return rgb, dep, acc, batch.rgb_tgt, batch.occ_pixels

    @staticmethod
    def _can_fuse(model, device, kitti2nusc, shapecode):
        return (FUSED_RENDER and not kitti2nusc and isinstance(model, models._DecoderBase) and shapecode.shape[0] == 1
                and torch.device(device).type == "cuda")

    def _decode_and_render(self, model, device, xyz, viewdir, z_vals, shapecode, texturecode, kitti2nusc):
        if kitti2nusc:
            xyz, viewdir = U._kitti2nusc(xyz, viewdir, device)
        sigmas, rgbs = model(xyz.to(device), viewdir.to(device), shapecode, texturecode)
        return self.volume_render(sigmas.squeeze(), rgbs, z_vals.to(device))

    def render_rays(self, model, device, img, mask_occ, cam_pose, obj_sz, K, roi, shapecode, texturecode,
                    kitti2nusc=False, im_sz=64, n_rays=None):
        """renderer.py:117-167."""
        if self._can_fuse(model, device, kitti2nusc, shapecode):
            px, py = U._pixel_grid_on(torch.device(device), roi, [im_sz, im_sz])
            img, mask_occ = U._resize_targets(img, mask_occ, im_sz)
            rgb_tgt = img.reshape(-1, 3).to(device, non_blocking=True)
            occ_pixels = mask_occ.reshape(-1, 1).to(device, non_blocking=True)
            if n_rays is not None:
                n_rays = np.minimum(px.numel(), n_rays)
                random_ray_ids = np.random.permutation(px.numel())[:n_rays]
                px, py = px[random_ray_ids], py[random_ray_ids]
                rgb_tgt = rgb_tgt[random_ray_ids]
                occ_pixels = occ_pixels[random_ray_ids]
            rgb_rays, depth_rays, acc_trans_rays = self._render_fused(model, device, px, py, K, cam_pose, obj_sz, shapecode,
                                                                      texturecode)
            return rgb_rays, depth_rays, acc_trans_rays, rgb_tgt, occ_pixels
        rays_o, viewdir = get_rays(K, cam_pose, roi, uv_steps=[im_sz, im_sz])
        img, mask_occ = U._resize_targets(img, mask_occ, im_sz)
        rgb_tgt = img.reshape(-1, 3).to(device, non_blocking=True)
        occ_pixels = mask_occ.reshape(-1, 1).to(device, non_blocking=True)
        if n_rays is not None:
            n_rays = np.minimum(rays_o.shape[0], n_rays)
            random_ray_ids = np.random.permutation(rays_o.shape[0])[:n_rays]
            rays_o = rays_o[random_ray_ids]
            viewdir = viewdir[random_ray_ids]
            rgb_tgt = rgb_tgt[random_ray_ids]
            occ_pixels = occ_pixels[random_ray_ids]
        xyz, viewdir, z_vals, intersect = self.prepare_sampled_rays(rays_o.to(device), viewdir.to(device), obj_sz)
        rgb_rays, depth_rays, acc_trans_rays = self._decode_and_render(model, device, xyz, viewdir, z_vals, shapecode,
                                                                       texturecode, kitti2nusc)
        return rgb_rays, depth_rays, acc_trans_rays, rgb_tgt, occ_pixels

    def render_rays_specified(self, model, device, img, mask_occ, cam_pose, obj_sz, K, roi, x_vec, y_vec, shapecode,
                              texturecode, kitti2nusc=False):
        """renderer.py:169-201."""
        if self._can_fuse(model, device, kitti2nusc, shapecode):
            dev = torch.device(device)
            px = torch.from_numpy(np.asarray(x_vec + roi[0].numpy())).t().reshape(-1).to(dev, torch.float32)
            py = torch.from_numpy(np.asarray(y_vec + roi[1].numpy())).t().reshape(-1).to(dev, torch.float32)
            rgb_tgt = img[y_vec, x_vec, :].to(device)
            occ_pixels = mask_occ[y_vec, x_vec, :].to(device)
            rgb_rays, depth_rays, acc_trans_rays = self._render_fused(model, device, px, py, K, cam_pose, obj_sz, shapecode,
                                                                      texturecode)
            return rgb_rays, depth_rays, acc_trans_rays, rgb_tgt, occ_pixels
        rays_o, viewdir = get_rays_specified(K, cam_pose, x_vec + roi[0].numpy(), y_vec + roi[1].numpy())
        rgb_tgt = img[y_vec, x_vec, :].to(device)
        occ_pixels = mask_occ[y_vec, x_vec, :].to(device)
        xyz, viewdir, z_vals, intersect = self.prepare_sampled_rays(rays_o.to(device), viewdir.to(device), obj_sz)
        rgb_rays, depth_rays, acc_trans_rays = self._decode_and_render(model, device, xyz, viewdir, z_vals, shapecode,
                                                                       texturecode, kitti2nusc)
        return rgb_rays, depth_rays, acc_trans_rays, rgb_tgt, occ_pixels

    def prepare_pixel_samples(self, img, mask_occ, cam_pose, obj_sz, K, roi, n_rays, im_sz=None):
        """renderer.py:203-236."""
        if im_sz is None:
            rays_o, viewdir = get_rays(K, cam_pose, roi)
        else:
            rays_o, viewdir = get_rays(K, cam_pose, roi, uv_steps=[im_sz, im_sz])
            img, mask_occ = U._resize_targets(img, mask_occ, im_sz)
        n_rays = np.minimum(rays_o.shape[0], n_rays)
        random_ray_ids = np.random.permutation(rays_o.shape[0])[:n_rays]
        rays_o = rays_o[random_ray_ids]
        viewdir = viewdir[random_ray_ids]
        rgb_tgt = img.reshape(-1, 3)[random_ray_ids]
        occ_pixels = mask_occ.reshape(-1, 1)[random_ray_ids]
        xyz, viewdir, z_vals, intersect = self.prepare_sampled_rays(rays_o, viewdir, obj_sz)
        return xyz, viewdir, z_vals, rgb_tgt, occ_pixels

    def render_full_img(self, model, device, cam_pose, obj_sz, K, roi, shapecode, texturecode, out_depth=False,
                        debug_occ=False, kitti2nusc=False):
        """renderer.py:238-294."""
        rays_o, viewdir = get_rays(K, cam_pose, roi)
        xyz, viewdir, z_vals, intersect = self.prepare_sampled_rays(rays_o.to(device), viewdir.to(device), obj_sz)
        if kitti2nusc:
            xyz, viewdir = U._kitti2nusc(xyz, viewdir, device)
        # one decoder launch + one compositing launch for the whole image (the reference's row-block loop, renderer.py:262-275,
        # only bounds its activation memory: every sample is decoded and composited independently of its block)
        sigmas, rgbs = model(xyz.to(device), viewdir.to(device), shapecode, texturecode)
        rgb_rays, depth_rays, acc_trans_rays = self.volume_render(sigmas.squeeze(), rgbs, z_vals.to(device))
        h, w = int(roi[3]) - int(roi[1]), int(roi[2]) - int(roi[0])
        generated_img = rgb_rays.reshape(h, w, 3)
        if debug_occ:
            import cv2
            acc = acc_trans_rays.reshape(h, w)
            cv2.imshow('est_occ', ((torch.ones_like(acc) - acc).cpu().numpy() * 255).astype(np.uint8))
            cv2.waitKey()
        if out_depth:
            return generated_img, depth_rays.reshape(h, w)
        return generated_img

    def render_virtual_imgs(self, model, device, obj_sz, K, shapecode, texturecode, radius=40., tilt=np.pi / 6,
                            pan_num=8, img_sz=128, kitti2nusc=False):
        """renderer.py:296-352 (visualisation helper)."""
        x_min, x_max = K[0, 2] - img_sz / 2, K[0, 2] + img_sz / 2
        y_min, y_max = K[1, 2] - img_sz / 2, K[1, 2] + img_sz / 2
        roi = np.asarray([x_min, y_min, x_max, y_max]).astype(np.int64)
        out = []
        for cam_pose in U.virtual_view_poses(radius, tilt, pan_num):
            img = self.render_full_img(model, device, cam_pose, obj_sz, K, roi, shapecode, texturecode, kitti2nusc=kitti2nusc)
            out.append(U._draw_axes(img, cam_pose, K, img_sz))
        return out


class GraphedBatchStep:
    """One refine-style step of B objects as ONE CUDA-graph launch: the host->device copy of the step's inputs from the caller's
    pinned host buffers (crops, occupancy masks, intrinsics, poses, codes), ``NeRFRenderer.render_batch``, the batched refine losses
    (optimizer_nuscenes.py:729-736), their backward to the poses and codes, and the device->host copy of
    ``[loss, d cam_pose (12), d shapecode (D), d texturecode (D)]`` per object into a pinned result buffer -- what the loop around
    ``render_rays`` does per iteration (optimizer_nuscenes.py:716-737), without ~0.3 ms of per-step Python and launch latency on the
    host.  Shapes, crops' rois and object sizes are fixed at construction (a refine loop re-renders the same crops); the CONTENT of
    the host buffers may change between steps.  ``run()`` enqueues a step; ``result`` (B, 1 + 12 + 2 D, pinned) is valid after a
    synchronisation of the stream (``torch.cuda.current_stream().synchronize()``).  Frozen weights (refine mode)."""

    def __init__(self, renderer, model, device, imgs, masks_occ, cam_poses, obj_szs, Ks, rois, shapecodes, texturecodes, im_sz=64,
                 loss_occ_coef=0.1, jitter=None, warmup=2):
        from . import losses
        device = torch.device(device)
        host = dict(imgs=imgs, masks_occ=masks_occ, cam_poses=cam_poses, Ks=Ks, shapecodes=shapecodes, texturecodes=texturecodes)
        for k, t in host.items():
            if not (torch.is_tensor(t) and t.device.type == "cpu" and t.is_pinned() and t.dtype == torch.float32 and t.is_contiguous()):
                raise ValueError("GraphedBatchStep: %s must be a contiguous pinned float32 host tensor" % k)
        b = cam_poses.shape[0]
        if tuple(imgs.shape) != (b, im_sz, im_sz, 3) or tuple(masks_occ.shape)[:3] != (b, im_sz, im_sz) or tuple(Ks.shape) != (b, 3, 3):
            raise ValueError("GraphedBatchStep: imgs (B,im_sz,im_sz,3), masks_occ (B,im_sz,im_sz[,1]) and Ks (B,3,3) at the render size")
        self.host, self.device, self.renderer, self.model = host, device, renderer, model
        d = shapecodes.shape[1]
        self.batch = renderer.make_batch(device, imgs, masks_occ, obj_szs, Ks, rois, im_sz)      # pixel grids, box constants; the rest is refreshed per step
        self.cam = torch.empty(b, 3, 4, device=device).requires_grad_()
        self.shp = torch.empty(b, d, device=device).requires_grad_()
        self.tex = torch.empty(b, d, device=device).requires_grad_()
        self._mask = torch.empty(b, im_sz * im_sz, 1, device=device)
        self.result = torch.empty(b, 1 + 12 + 2 * d, dtype=torch.float32).pin_memory()
        self._ones = torch.ones(b, device=device)
        self._jitter, self._coef, self._losses = jitter, float(loss_occ_coef), losses
        side = torch.cuda.Stream(device=device)
        side.wait_stream(torch.cuda.current_stream(device))
        with torch.cuda.stream(side):        # warm-up on a side stream (allocator, weight packing), as torch.cuda.graph requires
            for _ in range(max(1, int(warmup))):
                self._step()
        torch.cuda.current_stream(device).wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self._step()

    def _step(self):
        h, bt = self.host, self.batch
        with torch.no_grad():
            self.cam.copy_(h["cam_poses"], non_blocking=True)
            self.shp.copy_(h["shapecodes"], non_blocking=True)
            self.tex.copy_(h["texturecodes"], non_blocking=True)
            bt.K.copy_(h["Ks"], non_blocking=True)
            bt.rgb_tgt.copy_(h["imgs"].reshape(bt.rgb_tgt.shape), non_blocking=True)
            self._mask.copy_(h["masks_occ"].reshape(self._mask.shape), non_blocking=True)
            bt.occ_pixels = self._mask.type(torch.int32).type(torch.float32)          # renderer.py:133
        rgb, _dep, acc = self.renderer.render_batch(self.model, bt, self.cam, self.shp, self.tex, jitter=self._jitter)
        loss, parts = self._losses.refine_loss_batch(rgb, acc, bt.rgb_tgt, bt.occ_pixels, self._coef)
        g_cam, g_shp, g_tex = torch.autograd.grad(loss, [self.cam, self.shp, self.tex], grad_outputs=self._ones)
        b = self.cam.shape[0]
        self.result.copy_(torch.cat([parts[:, :1], g_cam.reshape(b, 12), g_shp, g_tex], 1), non_blocking=True)

    def run(self):
        """Enqueue one step on the current stream (one graph launch).  -> the pinned result buffer (valid after a stream synchronisation)."""
        self.graph.replay()
        return self.result


def volume_rendering3(sigmas, rgbs, z_vals, white_bkgd=False):
    """renderer.py:355-379: sigmas (N,S,1), rgbs (N,S,3), z_vals (N,S)."""
    return U._composite_any(sigmas, rgbs, z_vals, white_bkgd, True)


def render_rays_v3(model, device, img, mask_occ, cam_pose, obj_wlh, K, roi, n_samples, shapecode, texturecode,
                   shapenet_obj_cood, sym_aug, kitti2nusc=False, im_sz=64, n_rays=None, adjust_scale=1.0):
    """renderer.py:382-473.  As in the reference the slab test runs on DETACHED rays (renderer.py:425-432: numpy
    copies), so near/far carry no pose gradient; and, as there, the sampler uses a default NeRFRenderer()
    (renderer.py:394), i.e. 64 strata whatever ``n_samples`` says — viewdir is repeated ``n_samples`` times
    (renderer.py:436), so the two must agree for the shapes to line up."""
    renderer = NeRFRenderer()
    rays_o, viewdir = get_rays(K, cam_pose, roi, uv_steps=[im_sz, im_sz])
    img, mask_occ = U._resize_targets(img, mask_occ, im_sz)
    rgb_tgt = img.reshape(-1, 3).to(device, non_blocking=True)
    occ_pixels = mask_occ.reshape(-1, 1).to(device, non_blocking=True)
    if n_rays is not None:
        n_rays = np.minimum(rays_o.shape[0], n_rays)
        random_ray_ids = np.random.permutation(rays_o.shape[0])[:n_rays]
        rays_o = rays_o[random_ray_ids]
        viewdir = viewdir[random_ray_ids]
        rgb_tgt = rgb_tgt[random_ray_ids]
        occ_pixels = occ_pixels[random_ray_ids]
    obj_diag = np.linalg.norm(obj_wlh).astype(np.float32)
    obj_w, obj_l, obj_h = obj_wlh
    # the reference builds the AABB in float64 here (no .astype(np.float32), renderer.py:419-422) and divides the
    # float32 origins by the float32 half diagonal; the slab test then runs in float64 on the host.  The kernel
    # evaluates it in fp32 with the fp32-rounded box: identical hit masks except for rays within 1 ulp of an edge.
    half = np.asarray([obj_l / obj_diag, obj_w / obj_diag, obj_h / obj_diag]).astype(np.float32)
    # slab test on detached rays + 64-stratum sampler + xyz + z_vals (renderer.py:425-463) as ONE kernel: the box sampler with
    # detached bounds (near / far carry no gradient, exactly as the reference's numpy round trip); same RNG draw (one rand_like (N, 64))
    dev = torch.device(device)
    n = rays_o.shape[0]
    z_steps = _z_steps_on(dev, renderer.n_samples)
    jitter = torch.rand_like(z_steps.unsqueeze(0).repeat(n, 1))
    xyz, viewdir, z_vals, _hit = ops.sample_box(rays_o.to(dev), viewdir.to(dev), z_steps, jitter, obj_diag / 2, half, detach_bounds=True)
    if n_samples != renderer.n_samples:   # the reference repeats viewdir n_samples times against 64 strata: shapes must agree there too
        raise ValueError("render_rays_v3 samples with a default NeRFRenderer() (64 strata): n_samples must be 64")
    xyz = xyz * adjust_scale
    if sym_aug and random.uniform(0, 1) > 0.5:
        xyz = xyz * xyz.new_tensor([1., -1., 1.])
        viewdir = viewdir * viewdir.new_tensor([1., -1., 1.])
    if kitti2nusc:
        xyz, viewdir = U._kitti2nusc(xyz, viewdir, device)
    if shapenet_obj_cood:
        xyz, viewdir = U._swap(xyz), U._swap(viewdir)
    sigmas, rgbs = model(xyz.to(device), viewdir.to(device), shapecode, texturecode)
    rgb_rays, depth_rays, acc_trans_rays = volume_rendering3(sigmas, rgbs, z_vals.to(device))
    return rgb_rays, depth_rays, acc_trans_rays, rgb_tgt, occ_pixels
