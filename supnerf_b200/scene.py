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


def roi_process(roi, H=None, W=None, roi_margin=0, sq_pad=False):
    """utils.py:1392-1415: margin, optional square padding, clip to the image."""
    roi_new = roi.clone()
    roi_new[0:2] -= roi_margin
    roi_new[2:4] += roi_margin
    if sq_pad:
        cx, cy = (roi_new[0] + roi_new[2]) / 2, (roi_new[1] + roi_new[3]) / 2
        sz = torch.maximum(roi_new[2] - roi_new[0], roi_new[3] - roi_new[1])
        roi_new[0], roi_new[2] = cx - sz / 2, cx + sz / 2
        roi_new[1], roi_new[3] = cy - sz / 2, cy + sz / 2
    if H is not None and W is not None:
        roi_new[0:2] = torch.maximum(roi_new[0:2], torch.as_tensor(0))
        roi_new[2] = torch.minimum(roi_new[2], torch.as_tensor(W - 1))
        roi_new[3] = torch.minimum(roi_new[3], torch.as_tensor(H - 1))
    return roi_new


def scene_rays(K, obj_poses, obj_wlh, H, W, manipulation=(0., 0., 0.), rend_aabb=True, device="cuda"):
    """The ray bookkeeping of ``vis_scene`` (scripts/demo.py:437-523): move the objects, project their boxes to per-object rois,
    generate every object's rays over its roi in ITS frame (utils.get_rays), bound them by the object's box (slab test) or its
    bounding shell, and keep the pixels that at least one object covers.
    -> valid_rays (R, Nb, 8) = [o / (diag/2), d, near, far] (-1 where an object does not cover the pixel), valid_indices (H*W,) bool,
    obj_diags (Nb,), rois (Nb, 4) int32.  Ray generation and the slab test run in the package's kernels on `device`."""
    import numpy as np
    from . import pose_estimator as pe
    from . import utils as U
    device = torch.device(device)
    Nb = obj_poses.shape[0]
    obj_poses = obj_poses.clone().float()
    obj_poses[:, :, 3] += torch.tensor(manipulation, dtype=torch.float32).unsqueeze(0)
    K = K.float()
    corners_2d = pe.view_points_batch(pe.corners_of_box_batch(obj_poses, obj_wlh.float()), K.unsqueeze(0).repeat(Nb, 1, 1), normalize=True)
    rois = torch.zeros((Nb, 4), dtype=torch.float32)
    rois[:, 0], rois[:, 1] = corners_2d[:, 0].min(dim=1)[0], corners_2d[:, 1].min(dim=1)[0]
    rois[:, 2], rois[:, 3] = corners_2d[:, 0].max(dim=1)[0], corners_2d[:, 1].max(dim=1)[0]
    rois = rois.type(torch.int32)
    for ii in range(Nb):
        rois[ii] = roi_process(rois[ii], H, W, roi_margin=0, sq_pad=False)
    all_rays = torch.full((H, W, Nb, 8), -1.0, dtype=torch.float32, device=device)
    diags = []
    for i in range(Nb):
        pose = obj_poses[i]
        R_c2o = pose[:3, :3].transpose(0, 1)
        cam_pose = torch.cat([R_c2o, -R_c2o @ pose[:3, 3:]], dim=1).to(device)
        xmin, ymin, xmax, ymax = [int(v) for v in rois[i]]
        if xmax <= xmin or ymax <= ymin:
            diags.append(np.linalg.norm(obj_wlh[i].numpy()).astype(np.float32))
            continue
        rays_o, viewdir = U.get_rays(K.to(device), cam_pose, rois[i])
        diag = np.linalg.norm(obj_wlh[i].numpy()).astype(np.float32)
        diags.append(diag)
        h, w = ymax - ymin, xmax - xmin
        o_n = rays_o / (diag / 2)
        all_rays[ymin:ymax, xmin:xmax, i, :3] = o_n.view(h, w, 3)
        all_rays[ymin:ymax, xmin:xmax, i, 3:6] = viewdir.view(h, w, 3)
        if rend_aabb:
            ow, ol, oh = [float(v) for v in obj_wlh[i]]
            half = torch.tensor([ol / diag, ow / diag, oh / diag], dtype=torch.float32, device=device).reshape(1, 3).repeat(o_n.shape[0], 1)
            tn, tf, hit = ops.ray_box(o_n.contiguous(), viewdir.contiguous(), -half, half)
            minus1 = torch.full_like(tn, -1.0)
            all_rays[ymin:ymax, xmin:xmax, i, 6] = torch.where(hit, tn, minus1).view(h, w)
            all_rays[ymin:ymax, xmin:xmax, i, 7] = torch.where(hit, tf, minus1).view(h, w)
        else:
            n_ = torch.linalg.norm(cam_pose[:, -1])
            all_rays[ymin:ymax, xmin:xmax, i, 6] = (n_ - diag / 2) / (diag / 2)
            all_rays[ymin:ymax, xmin:xmax, i, 7] = (n_ + diag / 2) / (diag / 2)
    valid = (all_rays[..., 7].view(H * W, Nb) - all_rays[..., 6].view(H * W, Nb)).max(-1)[0] > 0
    return all_rays.view(H * W, Nb, 8)[valid], valid, torch.tensor(diags, dtype=torch.float32, device=device), rois


def render_scene(model, device, K, obj_poses, obj_wlh, shapecodes, texturecodes, H, W, n_samples, manipulation=(0., 0., 0.),
                 rend_aabb=True, adjust_scale=1.0, shapenet_obj_cood=True, ray_batch_size=4096, jitter=None):
    """``vis_scene`` (scripts/demo.py:425-579): several reconstructed objects rendered together into one H x W image from a
    (possibly manipulated) camera -- per pixel every object's samples (64-stratum sampler between its box bounds), one decoder call
    over all objects' rows (object-major, the decoder's batched-latent layout), empty space white with zero density, the samples of
    all objects merged by depth (snb_merge_sort_samples) and composited on a white background.  Forward only (the reference runs it
    under no_grad).  `jitter`: optional (R, Nb, S) uniform draws replacing the per-batch torch.rand_like (tests).
    -> canvas (H, W, 3) uint8 like the reference."""
    from . import utils as U
    device = torch.device(device)
    Nb = obj_poses.shape[0]
    valid_rays, valid, diags, _ = scene_rays(K, obj_poses, obj_wlh, H, W, manipulation, rend_aabb, device)
    canvas = torch.ones(H * W, 3, dtype=torch.float32, device=device)
    outs = []
    with torch.no_grad():
        step = 1.0 / n_samples
        z_steps = torch.linspace(0, 1 - step, n_samples, device=device)
        shp, tex = shapecodes.to(device), texturecodes.to(device)
        for r0 in range(0, valid_rays.shape[0], ray_batch_size):
            batch = valid_rays[r0:r0 + ray_batch_size]
            Nr = batch.shape[0]
            rays = batch.reshape(-1, 8).contiguous()
            if jitter is None:
                jit = torch.rand_like(z_steps.unsqueeze(0).repeat(rays.shape[0], 1))
            else:
                jit = jitter[r0:r0 + Nr].reshape(-1, n_samples).to(device)
            z_coarse = ops.stratified_z(rays, z_steps, jit)                                   # sample_from_rays_v2
            empty = z_coarse == -1
            xyz = rays[:, None, :3] + z_coarse[:, :, None] * rays[:, None, 3:6]
            viewdir = rays[:, 3:6].unsqueeze(-2).repeat(1, n_samples, 1)
            half_d = (diags.view(1, -1, 1, 1).repeat(Nr, 1, 1, 1).flatten(0, 1)) / 2
            z_vals = torch.norm((xyz - rays[:, None, :3]) * half_d, p=2, dim=-1)
            z_vals[empty] = -1
            xyz = xyz.view(Nr, Nb, n_samples, 3).permute(1, 0, 2, 3).flatten(0, 1) * adjust_scale
            viewdir = viewdir.view(Nr, Nb, n_samples, 3).permute(1, 0, 2, 3).flatten(0, 1)
            if shapenet_obj_cood:
                xyz, viewdir = U._swap(xyz), U._swap(viewdir)
            n_rows = Nr * n_samples
            pad = (-n_rows) % 128 if (model.precision or "") == "bf16" else 0   # the tcgen05 decoder needs whole 128-row tiles per object
            if pad:
                extra = (pad + n_samples - 1) // n_samples
                xyz = torch.cat([xyz.view(Nb, Nr, n_samples, 3), xyz.view(Nb, Nr, n_samples, 3)[:, -1:].expand(Nb, extra, n_samples, 3)], 1).flatten(0, 1)
                viewdir = torch.cat([viewdir.view(Nb, Nr, n_samples, 3), viewdir.view(Nb, Nr, n_samples, 3)[:, -1:].expand(Nb, extra, n_samples, 3)], 1).flatten(0, 1)
            sigmas, rgbs = model(xyz.contiguous(), viewdir.contiguous(), shp, tex)
            Nr_p = xyz.shape[0] // Nb
            rgbs = rgbs.view(Nb, Nr_p, n_samples, 3)[:, :Nr].permute(1, 0, 2, 3).flatten(0, 1).clone()
            sigmas = sigmas.view(Nb, Nr_p, n_samples)[:, :Nr].permute(1, 0, 2).flatten(0, 1).clone()
            rgbs[empty] = 1
            sigmas[empty] = 0
            rgb, _, _ = render_merged(z_vals.view(-1, Nb * n_samples), sigmas.view(-1, Nb * n_samples), rgbs.view(-1, Nb * n_samples, 3))
            outs.append(rgb)
    if outs:
        canvas[valid] = torch.cat(outs, 0)
    return (canvas.view(H, W, 3).cpu().numpy() * 255).astype("uint8")


def save_opts_w_pose(path, num_obj, optimized_shapecodes, optimized_texturecodes, optimized_poses, psnr_eval, ssim_eval, depth_err_mean,
                     lidar_pts_cnt, R_eval, T_eval):
    """The refine loops' result file ``codes+poses.pth`` (optimizer_nuscenes.py:1463-1476): same keys, torch.save."""
    torch.save({"num_obj": num_obj, "optimized_shapecodes": optimized_shapecodes, "optimized_texturecodes": optimized_texturecodes,
                "optimized_poses": optimized_poses, "psnr_eval": psnr_eval, "ssim_eval": ssim_eval, "depth_err_mean": depth_err_mean,
                "lidar_pts_cnt": lidar_pts_cnt, "R_eval": R_eval, "T_eval": T_eval}, path)


def save_cross_eval(path, psnr_eval_mat_per_ins, depth_eval_mat_per_ins, cnt_lidar_pts_per_ins, code_save_iters):
    """``cross_eval.pth`` (optimizer_nuscenes.py:1405-1409)."""
    torch.save({"psnr_eval_mat_per_ins": psnr_eval_mat_per_ins, "depth_eval_mat_per_ins": depth_eval_mat_per_ins,
                "cnt_lidar_pts_per_ins": cnt_lidar_pts_per_ins, "CODE_SAVE_ITERS_": code_save_iters}, path)


def load_result(path):
    """Either result file written by the reference or by this package (plain torch.load of a dict of tensors / dicts)."""
    return torch.load(path, map_location="cpu", weights_only=False)
